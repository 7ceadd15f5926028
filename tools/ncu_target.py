#!/usr/bin/env python3
"""Small single-purpose launches for ncu captures: python tools/ncu_target.py <what> [N]"""
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth  # noqa: E402

what = sys.argv[1]
torch.manual_seed(0)
if what.startswith("linear"):
    dt = torch.bfloat16 if "bf16" in what else torch.float32
    H = 256 if "h256" in what else 128
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    x = torch.randn(n, H, device="cuda").to(dt)
    agg = torch.randn(n, H, device="cuda").to(dt)
    w = (torch.randn(H, H, device="cuda") / H**0.5).to(dt)
    b = torch.randn(H, device="cuda")
    g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    for _ in range(3):
        ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=("train" in what), dropout_p=0.1 if "train" in what else 0.0, seed=7)
    if "dx" in what:
        for _ in range(3):
            ops.linear_bwd_input(x, w, row_scale=b.new_ones(n))
    if "dw" in what:
        gy = torch.randn(n, H, device="cuda").to(dt)
        for _ in range(3):
            ops.linear_bwd_weight(gy, agg, x, want_bias=False)
    if "epi" in what:
        pre = torch.randn(n, H, device="cuda").to(dt)
        stats = torch.rand(n, 2, device="cuda") + 0.5
        for _ in range(3):
            ops.epilogue_bwd(x, n, H, x, pre=pre, stats=stats, ln=(g, be), relu=True, dropout_p=0.1, seed=7, want_bias_grad=True)
elif what.startswith("agg"):
    if "cfg4" in what:
        mesh = synth.tet_lattice_wing(2_000_000, seed=42, node_order="random" if "random" in what else "native", shuffle_edges=False)
        H, dt = 256, torch.bfloat16
    else:
        mesh = synth.surface_tri_wing(200000, seed=42) if "tri" in what else synth.tet_lattice_wing(200000, seed=42)
        H, dt = 128, torch.float32
    ei = torch.from_numpy(mesh["edge_index"]).cuda()
    g = ops.get_graph(ei, mesh["num_nodes"])
    x = torch.randn(mesh["num_nodes"], H, device="cuda").to(dt)
    if what.startswith("aggs"):  # source-scaled gather over the transposed CSR (backward)
        rp_t, col_t = g.transpose()
        for _ in range(3):
            ops.aggregate_scaled(rp_t, col_t, g.inv_deg, x)
    else:
        for _ in range(3):
            ops.aggregate(g.rowptr, g.col, g.inv_deg, x)
torch.cuda.synchronize()
print("ok", what)
