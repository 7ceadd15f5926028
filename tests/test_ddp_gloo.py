"""N>1 path on CPU: world_size-2 gloo run of the data-parallel wrapper.  The averaged gradients of
the two ranks (each with its own mesh) must equal the single-process gradient of the union batch -
including the masked-count re-weighting of the loss (model.py:151)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deep_fem_uav_wing.gnn import synth
    from deep_fem_uav_wing.gnn.ddp import MeshDataParallel
    from oracle.sage_oracle import GraphSAGEModelRef, MaskedMSELossRef

    torch.manual_seed(100 + rank)  # different init per rank: the wrapper must broadcast rank 0's weights
    model = GraphSAGEModelRef(10, 16, 1, 2, dropout=0.0)
    ddp = MeshDataParallel(model, num_buckets=3)
    meshes = [synth.surface_tri_wing(300 + 60 * r, seed=r) for r in range(world)]
    for i_, m_ in enumerate(meshes):  # unequal masked counts across ranks
        m_["loss_mask"][: 40 * (1 + i_)] = False
    mine = meshes[rank]
    x, ei = torch.from_numpy(mine["x"]), torch.from_numpy(mine["edge_index"])
    y, m = torch.from_numpy(mine["y"]), torch.from_numpy(mine["loss_mask"])
    ddp.zero_grad()
    out = ddp(x, ei)
    loss = MaskedMSELossRef()(out, y, m)
    ddp.scale_loss(loss, m.sum().float()).backward()
    ddp.finish()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    if rank == 0:
        # single-process reference on the union batch with the same (rank-0) weights
        ref = GraphSAGEModelRef(10, 16, 1, 2, dropout=0.0)
        ref.load_state_dict(model.state_dict())
        off, xs, eis, ys, ms = 0, [], [], [], []
        for mm in meshes:
            xs.append(torch.from_numpy(mm["x"])); eis.append(torch.from_numpy(mm["edge_index"]) + off)
            ys.append(torch.from_numpy(mm["y"])); ms.append(torch.from_numpy(mm["loss_mask"])); off += mm["num_nodes"]
        l = MaskedMSELossRef()(ref(torch.cat(xs), torch.cat(eis, 1)), torch.cat(ys), torch.cat(ms))
        l.backward()
        worst = max(((grads[k] - p.grad).norm() / p.grad.norm().clamp_min(1e-30)).item() for k, p in ref.named_parameters())
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ddp_two_ranks_equals_union_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst = q.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst < 1e-5, worst
