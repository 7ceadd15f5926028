#!/usr/bin/env python3
"""Knock-out timing of the persistent tensor-core linear (dev probe): run once per DFW_TC_KNOCK value (read once per process).
Results with knock != 0 are numerically wrong by construction; only the times matter."""
import json
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


n, H = 200000, 128
dt = torch.float32 if os.environ.get("LIN_DT", "f32") == "f32" else torch.bfloat16
x = torch.randn(n, H, device="cuda").to(dt); agg = torch.randn(n, H, device="cuda").to(dt)
w = (torch.randn(H, H, device="cuda") / H ** 0.5).to(dt); b = torch.randn(H, device="cuda")
g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
print(json.dumps({"knock": os.environ.get("DFW_TC_KNOCK", "0"), "dt": str(dt),
                  "fwd_train_us": t(lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=True, dropout_p=0.1, seed=3)),
                  "fwd_infer_us": t(lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x)),
                  "single_us": t(lambda: ops.linear_fwd(x, w, bias=b, relu=True))}))
