#!/usr/bin/env python3
"""Worker of tests/test_partition_nccl.py (torchrun, one rank per GPU): the eval forward of ONE mesh partitioned over the
ranks (gnn/partition.py: 1-D node partition in k-d order + one NCCL all-to-all of halo rows per SAGE layer) against the same
model's unpartitioned forward on rank 0's GPU.  Writes the relative error and the halo statistics to OUT.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    from deep_fem_uav_wing.gnn import synth
    from deep_fem_uav_wing.gnn.model import GraphSAGEModel
    from deep_fem_uav_wing.gnn.partition import PartitionedMeshInference

    res = {"world": world, "cases": []}
    from deep_fem_uav_wing.gnn import ops

    min_nodes, min_reuse = ops.TC_AGG_MIN_NODES, ops.TC_AGG_MIN_REUSE
    for kind, n, h, layers, dtype, blocks in (("tet-random", 60000, 64, 3, torch.float32, False), ("tri", 50000, 128, 4, torch.float32, False),
                                              ("tet-random", 120000, 128, 3, torch.bfloat16, False), ("tet-random", 120000, 128, 3, torch.bfloat16, True)):
        # blocks: the local aggregation runs the tensor-core block kernel on the rectangular [own | halo] layout (the size policy
        # would leave a 60k-row piece to the gather kernel)
        ops.TC_AGG_MIN_NODES, ops.TC_AGG_MIN_REUSE = (1000, 0.0) if blocks else (min_nodes, min_reuse)
        mesh = synth.tet_lattice_wing(n, seed=5, node_order="random") if kind.startswith("tet") else synth.surface_tri_wing(n, seed=5)
        x, ei = torch.from_numpy(mesh["x"]).to(device), torch.from_numpy(mesh["edge_index"]).to(device)
        torch.manual_seed(11)  # identical weights on every rank
        model = GraphSAGEModel(10, h, 1, layers, dropout=0.0).to(device).eval().set_compute_dtype(dtype)
        model.node_reorder = "never"
        pm = PartitionedMeshInference(model, x, ei)
        assert (pm.plan is not None) == blocks, "local block plan expected exactly in the `blocks` case"
        ops.TC_AGG_MIN_NODES, ops.TC_AGG_MIN_REUSE = min_nodes, min_reuse
        out_own = pm()
        full = pm.gather(out_own)
        torch.cuda.synchronize()
        halo = torch.tensor([pm.halo_rows, pm.part.n_own], device=device)
        allh = [torch.zeros_like(halo) for _ in range(world)]
        dist.all_gather(allh, halo)
        if rank == 0:
            with torch.no_grad():
                ref = model(x, ei)
            err = ((full.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
            res["cases"].append({"mesh": kind, "N": int(x.shape[0]), "hidden": h, "layers": layers, "dtype": str(dtype).replace("torch.", ""),
                                 "local_aggregation": "tensor-core blocks" if blocks else "gather", "rel_max_vs_unpartitioned": err, "halo_rows_per_rank": [int(t[0]) for t in allh], "own_rows_per_rank": [int(t[1]) for t in allh]})
        dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
