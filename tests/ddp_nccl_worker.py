#!/usr/bin/env python3
"""Worker of tests/test_ddp_nccl.py - launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/ddp_nccl_worker.py OUT.json

Every rank runs the PRODUCT path that bench.py and scripts/train_gnn.py run under torchrun - ``GraphSAGEModel`` +
``MeshDataParallel`` + ``GraphedTrainStep(ddp=...)`` over NCCL, i.e. the bucketed gradient all-reduce and the masked-count
all-reduce captured inside the step's CUDA graph - on its own meshes (different node counts and different masked counts
per rank).  Rank 0 then repeats the same steps in ONE process on the union batch (reference semantics: the loss is the mean
over the masked nodes of the whole batch, ``model.py:151``; batches are disjoint unions of meshes, ``train_gnn.py:150``)
and writes the worst relative L2 distance of (i) the first-step gradients and (ii) every parameter after each step.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]

STEPS, H, L = 4, 64, 3
LR, WD = 1e-3, 1e-4  # train_gnn.py:167


def rank_batch(r, per_rank, device):
    """``per_rank`` meshes of rank ``r`` as one disjoint-union batch (sizes and masked counts differ across ranks)."""
    from deep_fem_uav_wing.gnn import synth
    from deep_fem_uav_wing.gnn.loader import Batch, Data

    datas = []
    for j in range(per_rank):
        m = synth.surface_tri_wing(2400 + 640 * r + 320 * j, seed=10 * r + j)
        m["loss_mask"][: 150 * (1 + r) + 40 * j] = False
        datas.append(Data(x=torch.from_numpy(m["x"]), edge_index=torch.from_numpy(m["edge_index"]), y=torch.from_numpy(m["y"]),
                          loss_mask=torch.from_numpy(m["loss_mask"])))
    return Batch.from_data_list(datas).to(device), datas


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    from deep_fem_uav_wing.gnn import ops
    from deep_fem_uav_wing.gnn.ddp import MeshDataParallel
    from deep_fem_uav_wing.gnn.graphed import GraphedTrainStep
    from deep_fem_uav_wing.gnn.loader import Batch
    from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss

    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
    model = GraphSAGEModel(10, H, 1, L, dropout=0.0).to(device).train()
    ddp = MeshDataParallel(model)
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD, fused=True, capturable=True)
    crit = MaskedMSELoss()
    step = GraphedTrainStep(model, crit, opt, eager_steps=1, ddp=ddp)  # step 1 eager, step 2 captured, steps 2.. replayed
    b, _ = rank_batch(rank, 2, device)
    k0 = ops.LAUNCH_COUNTER["kernels"]
    traj, grads1, losses = [], None, []
    for s in range(STEPS):
        loss = step(b.x, b.edge_index, b.y, b.loss_mask)
        losses.append(float(loss.item()))
        if s == 0:
            grads1 = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        traj.append({k: p.detach().clone() for k, p in model.named_parameters()})
    captured = bool(step._graphs)
    launches = ops.LAUNCH_COUNTER["kernels"] - k0
    torch.cuda.synchronize()
    # every rank must hold the same parameters after the all-reduced steps
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ranks_agree = bool(torch.equal(lo, hi))
    dist.barrier()

    if rank == 0:
        # ---- the same steps in one process on the union batch (eager, no DDP) ----
        ref = GraphSAGEModel(10, H, 1, L, dropout=0.0).to(device).train()
        ref.load_state_dict(init)
        ropt = torch.optim.AdamW(ref.parameters(), lr=LR, weight_decay=WD, fused=True, capturable=True)
        datas = []
        for r in range(world):
            datas += rank_batch(r, 2, device)[1]
        ub = Batch.from_data_list(datas).to(device)
        worst_grad, worst_param, ref_losses = 0.0, [], []
        for s in range(STEPS):
            ropt.zero_grad(set_to_none=True)
            l = crit(ref(ub.x, ub.edge_index, ub.batch), ub.y, ub.loss_mask)
            l.backward()
            if s == 0:
                for k, p in ref.named_parameters():
                    worst_grad = max(worst_grad, rel_l2(grads1[k], p.grad))
            ropt.step()
            ref_losses.append(float(l.item()))
            worst_param.append(max(rel_l2(traj[s][k], p.detach()) for k, p in ref.named_parameters()))
        res = {"world": world, "backend": dist.get_backend(), "steps": STEPS, "hidden": H, "layers": L,
               "graph_captured_with_nccl": captured, "ranks_hold_identical_parameters": ranks_agree,
               "worst_first_step_grad_rel_l2": worst_grad, "worst_param_rel_l2_per_step": worst_param,
               "rank0_local_losses": losses, "union_batch_losses": ref_losses, "kernel_launches_rank0": launches,
               "nodes_per_rank": [int(sum(d.x.shape[0] for d in rank_batch(r, 2, "cpu")[1])) for r in range(world)],
               "masked_per_rank": [int(sum(int(d.loss_mask.sum()) for d in rank_batch(r, 2, "cpu")[1])) for r in range(world)]}
        with open(out_path, "w") as f:
            json.dump(res, f)
        print(json.dumps(res), flush=True)
    dist.barrier()
    # the captured step holds NCCL work: leave without tearing the communicator down (see bench.py)
    step = None
    import gc

    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
