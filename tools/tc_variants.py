#!/usr/bin/env python3
"""Timing of dfw_linear_fwd variants to separate mainloop and epilogue cost (development aid)."""
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200"), os.path.join(REPO, "tools")]
from deep_fem_uav_wing.gnn import ops  # noqa: E402
from quick_prof import timeit  # noqa: E402

n = 200000
for dt in (torch.float32, torch.bfloat16):
    for H in (128, 256):
        x = torch.randn(n, H, device="cuda").to(dt)
        agg = torch.randn(n, H, device="cuda").to(dt)
        w = (torch.randn(H, H, device="cuda") / H**0.5).to(dt)
        b = torch.randn(H, device="cuda")
        g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
        cases = {
            "full(train)": lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=True),
            "infer(LN+res)": lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x),
            "plain 2-input": lambda: ops.linear_fwd(agg, w, x, w, bias=b),
            "plain 1-input": lambda: ops.linear_fwd(agg, w, bias=b),
            "rowdot only": lambda: ops.linear_fwd(agg, w, bias=b, relu=True, rowdot=(b, None), want_out=False),
        }
        for name, fn in cases.items():
            t = timeit(fn, iters=10)
            print(f"{str(dt)[6:]:9s} H={H} {name:14s} {t*1e6:8.1f} us", flush=True)
