#!/usr/bin/env python3
"""Which torch (non-dfw) kernels run inside one config-2 training step, and from which aten ops (dev aid)."""
import os, sys, json
import torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss
from deep_fem_uav_wing.gnn.loader import Batch
from deep_fem_uav_wing.gnn.dataset import Data  # noqa
import bench
dev = torch.device("cuda", 0)
datas = bench.make_dataset(16, 12500) if hasattr(bench, "make_dataset") else None
ms = [synth.surface_tri_wing(12500, seed=i) for i in range(16)]
x = torch.cat([torch.from_numpy(m["x"]) for m in ms]).to(dev)
off = 0; eis = []
for m in ms:
    eis.append(torch.from_numpy(m["edge_index"]) + off); off += m["num_nodes"]
ei = torch.cat(eis, 1).to(dev)
y = torch.cat([torch.from_numpy(m["y"]) for m in ms]).to(dev)
mask = torch.cat([torch.from_numpy(m["loss_mask"]) for m in ms]).to(dev)
model = GraphSAGEModel(10, 128, 1, 4, 0.1).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
crit = MaskedMSELoss()
g = ops.get_graph(ei, x.shape[0]); g.transpose()
def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x, g, None), y, mask)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CPU and e.key.startswith("aten::") and getattr(e, "device_time_total", 0) > 0:
        rows.append((e.key, e.count, round(e.device_time_total, 1)))
rows.sort(key=lambda r: -r[2])
for r in rows[:30]: print(r)
