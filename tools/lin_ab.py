#!/usr/bin/env python3
"""A/B of the tensor-core linear variants (DFW_TC_PERSIST=0|1 is read once per process, so run this twice):
SAGE-layer forward / input-gradient shapes in fp32 and bf16, L2 flushed, CUDA events."""
import json
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=20):
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


res = {"DFW_TC_PERSIST": os.environ.get("DFW_TC_PERSIST", "default")}
for name, n, H, dt in (("cfg2 fp32 H128 N200k", 200000, 128, torch.float32), ("cfg5 bf16 H128 N320k", 320000, 128, torch.bfloat16),
                       ("cfg4 bf16 H256 N2M", 2_001_384, 256, torch.bfloat16), ("fp32 H64 N200k", 200000, 64, torch.float32)):
    x = torch.randn(n, H, device="cuda").to(dt); agg = torch.randn(n, H, device="cuda").to(dt)
    w = (torch.randn(H, H, device="cuda") / H ** 0.5).to(dt); b = torch.randn(H, device="cuda")
    g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    res[name] = {
        "fwd_train_us": t(lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=True, dropout_p=0.1, seed=3)),
        "fwd_infer_us": t(lambda: ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x)),
        "bwd_input_us": t(lambda: ops.linear_fwd(agg, w, x, w, residual=x, transpose_w=True)),
        "single_operand_relu_us": t(lambda: ops.linear_fwd(x, w, bias=b, relu=True)),
    }
    del x, agg
print(json.dumps(res))
