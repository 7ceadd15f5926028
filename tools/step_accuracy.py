#!/usr/bin/env python3
"""Accuracy of the backward building blocks vs fp64 at a given H (development aid)."""
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth
torch.manual_seed(0)
rl2 = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
mesh = synth.surface_tri_wing(6000, seed=3)
n = mesh["num_nodes"]
ei = torch.from_numpy(mesh["edge_index"]).cuda()
g = ops.get_graph(ei, n)
rp_t, col_t = g.transpose()
for H in (128, 256):
    gy = torch.randn(n, H, device="cuda") * 1e-3 + 2e-3      # gradients with a common sign component
    gout = torch.randn(n, H, device="cuda") * 1e-3
    wl = torch.randn(H, H, device="cuda") / H**0.5; wr = torch.randn(H, H, device="cuda") / H**0.5
    gt = ops.aggregate_scaled(rp_t, col_t, g.inv_deg, gy)
    src, dst = ei[0], ei[1]
    ref_gt = torch.zeros(n, H, dtype=torch.float64, device="cuda").index_add_(0, src, (gy.double() * g.inv_deg.double()[:, None])[dst])
    gx, _, _, _ = ops.linear_fwd(gt, wl.t().contiguous(), gy, wr.t().contiguous(), residual=gout)
    ref_gx = ref_gt @ wl.double() + gy.double() @ wr.double() + gout.double()
    ref_gx_from_gt = gt.double() @ wl.double() + gy.double() @ wr.double() + gout.double()
    print(f"H={H}: aggregate_scaled {rl2(gt, ref_gt):.2e} | fused dX (chain) {rl2(gx, ref_gx):.2e} | fused dX (given gt) {rl2(gx, ref_gx_from_gt):.2e} "
          f"| column-sum err {((gx.double().sum(0) - ref_gx.sum(0)).norm() / ref_gx.sum(0).norm()).item():.2e}")
    pre = torch.randn(n, H, device="cuda"); stats = torch.stack([pre.mean(1), 1.0 / torch.sqrt(pre.var(1, unbiased=False) + 1e-5)], 1).contiguous()
    gam, bet = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda")
    g_y, dgamma, dbeta, _, _, dbias = ops.epilogue_bwd(gout, n, H, gout, pre=pre, stats=stats, ln=(gam, bet), relu=True, want_bias_grad=True)
    xh = (pre.double() - stats[:, :1].double()) * stats[:, 1:].double()
    z = xh * gam.double() + bet.double()
    gg = gout.double() * (z > 0)
    ref_dbeta, ref_dgamma = gg.sum(0), (gg * xh).sum(0)
    ggam = gg * gam.double()
    ref_gy = stats[:, 1:].double() * (ggam - ggam.mean(1, keepdim=True) - xh * (ggam * xh).mean(1, keepdim=True))
    print(f"      epilogue_bwd g_y {rl2(g_y, ref_gy):.2e} dgamma {rl2(dgamma, ref_dgamma):.2e} dbeta {rl2(dbeta, ref_dbeta):.2e} dbias {rl2(dbias, ref_gy.sum(0)):.2e}")
