#!/usr/bin/env python3
"""One cfg2 training step (4 x 50k-node meshes, H=128, L=4, fp32) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --set full ... python tools/ncu_step.py"""
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import synth  # noqa: E402
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "tri"
gen = synth.surface_tri_wing if kind == "tri" else synth.tet_lattice_wing
meshes = [gen(50000, seed=42 + i) for i in range(4)]
x = torch.cat([torch.from_numpy(m["x"]) for m in meshes]).cuda()
off, eis = 0, []
for m in meshes:
    eis.append(torch.from_numpy(m["edge_index"]) + off)
    off += m["num_nodes"]
ei = torch.cat(eis, 1).cuda()
y = torch.cat([torch.from_numpy(m["y"]) for m in meshes]).cuda()
mask = torch.cat([torch.from_numpy(m["loss_mask"]) for m in meshes]).cuda()
torch.manual_seed(42)
model = GraphSAGEModel(10, 128, 1, 4, dropout=0.1).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
crit = MaskedMSELoss()


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(x, ei, None), y, mask)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
