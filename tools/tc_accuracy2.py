#!/usr/bin/env python3
"""fp32 accuracy of the tensor-core kernels vs fp64, per shape (development aid)."""
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops
torch.manual_seed(0)
def rl2(a, b): return ((a.double() - b).norm() / b.norm()).item()
def rmax(a, b): return ((a.double() - b).abs().max() / b.abs().max()).item()
n = 20000
for H in (64, 128, 256):
    a1 = torch.randn(n, H, device="cuda"); a2 = torch.randn(n, H, device="cuda")
    w1 = torch.randn(H, H, device="cuda") / H**0.5; w2 = torch.randn(H, H, device="cuda") / H**0.5
    b = torch.randn(H, device="cuda")
    out, _, _, _ = ops.linear_fwd(a1, w1, a2, w2, bias=b)
    ref = a1.double() @ w1.double().T + a2.double() @ w2.double().T + b.double()
    ref32 = a1 @ w1.T + a2 @ w2.T + b
    g = torch.randn(n, H, device="cuda")
    dw1, dw2, _ = ops.linear_bwd_weight(g, a1, a2, want_bias=False)
    dref = g.double().T @ a1.double()
    dref32 = g.T @ a1
    dx = ops.linear_bwd_input(g, w1)
    xref = g.double() @ w1.double()
    print(f"H={H}: fwd rel_l2 {rl2(out, ref):.2e} rel_max {rmax(out, ref):.2e} (torch fp32 {rl2(ref32, ref):.2e}) | dW rel_l2 {rl2(dw1, dref):.2e} (torch fp32 {rl2(dref32, dref):.2e}) | dX rel_l2 {rl2(dx, xref):.2e}")
