#!/usr/bin/env python3
"""Encoder/decoder-shaped launches timed alone (development aid)."""
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200"), os.path.join(REPO, "tools")]
from deep_fem_uav_wing.gnn import ops  # noqa: E402
from quick_prof import timeit, report  # noqa: E402

n = 200000
x10 = torch.randn(n, 10, device="cuda")
w1 = torch.randn(64, 10, device="cuda")
b1 = torch.randn(64, device="cuda")
h64 = torch.randn(n, 64, device="cuda")
w2 = torch.randn(128, 64, device="cuda") / 8
b2 = torch.randn(128, device="cuda")
h128 = torch.randn(n, 128, device="cuda")
w3 = torch.randn(64, 128, device="cuda") / 11
w4 = torch.randn(64, device="cuda")
g64 = torch.randn(n, 64, device="cuda")
g128 = torch.randn(n, 128, device="cuda")
report("enc1 fwd 10->64", timeit(lambda: ops.linear_fwd(x10, w1, bias=b1, relu=True)), n * 74 * 4)
report("enc2 fwd 64->128", timeit(lambda: ops.linear_fwd(h64, w2, bias=b2, relu=True)), n * 192 * 4)
report("dec fwd 128->64 +rowdot", timeit(lambda: ops.linear_fwd(h128, w3, bias=b1, relu=True, dropout_p=0.1, seed=3, rowdot=(w4, None))), n * 193 * 4)
report("enc1 dW (g64^T x10)", timeit(lambda: ops.linear_bwd_weight(g64, x10, None, want_bias=False)), n * 74 * 4)
report("enc2 dW (g128^T h64)", timeit(lambda: ops.linear_bwd_weight(g128, h64, None, want_bias=False)), n * 192 * 4)
report("dec dW (g64^T h128)", timeit(lambda: ops.linear_bwd_weight(g64, h128, None, want_bias=False)), n * 192 * 4)
report("enc2 dX (g128 W2)", timeit(lambda: ops.linear_bwd_input(g128, w2)), n * 192 * 4)
report("dec dX (g64 W3)", timeit(lambda: ops.linear_bwd_input(g64, w3)), n * 192 * 4)
report("enc1 epi_bwd relu 64", timeit(lambda: ops.epilogue_bwd(g64, n, 64, g64, act=h64, relu=True, want_bias_grad=True)), n * 64 * 4 * 3)
report("enc2 epi_bwd relu 128", timeit(lambda: ops.epilogue_bwd(g128, n, 128, g128, act=h128, relu=True, want_bias_grad=True)), n * 128 * 4 * 3)
g1 = torch.randn(n, device="cuda")
report("dec epi_bwd rowdot", timeit(lambda: ops.epilogue_bwd(None, n, 64, h64, g_rowdot=g1, rowdot_w=w4, act=h64, relu=True, dropout_p=0.1, seed=3, want_bias_grad=True)), n * 64 * 4 * 2)
