#!/usr/bin/env python3
"""Per-parameter gradient error of the CUDA model vs an fp64 oracle (development aid): python tools/grad_accuracy.py H L [p]"""
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import synth
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss
from oracle.sage_oracle import GraphSAGEModelRef, MaskedMSELossRef
H, L = int(sys.argv[1]), int(sys.argv[2])
mesh = synth.surface_tri_wing(6000, seed=3)
x, ei = torch.from_numpy(mesh["x"]), torch.from_numpy(mesh["edge_index"])
y, m = torch.from_numpy(mesh["y"]), torch.from_numpy(mesh["loss_mask"])
torch.manual_seed(11)
ref = GraphSAGEModelRef(10, H, 1, L, dropout=0.0)
model = GraphSAGEModel(10, H, 1, L, dropout=0.0)
model.load_state_dict(ref.state_dict())
model = model.cuda()
ref64 = GraphSAGEModelRef(10, H, 1, L, dropout=0.0).double()
ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
MaskedMSELoss()(model(x.cuda(), ei.cuda()), y.cuda(), m.cuda()).backward()
MaskedMSELossRef()(ref(x, ei), y, m).backward()
MaskedMSELossRef()(ref64(x.double(), ei), y.double(), m).backward()
rl2 = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
for (k, q), qr, q64 in zip(model.named_parameters(), ref.parameters(), ref64.parameters()):
    print(f"{k:28s} kernel {rl2(q.grad.cpu(), q64.grad):.2e}   oracle32 {rl2(qr.grad, q64.grad):.2e}")
