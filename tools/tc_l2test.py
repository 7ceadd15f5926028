#!/usr/bin/env python3
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200"), os.path.join(REPO, "tools")]
from deep_fem_uav_wing.gnn import ops
from quick_prof import timeit
for n in (148 * 128, 148 * 128 * 4):
    for dt in (torch.float32, torch.bfloat16):
        H = 128
        x = torch.randn(n, H, device="cuda").to(dt); agg = torch.randn(n, H, device="cuda").to(dt)
        w = (torch.randn(H, H, device="cuda") / H**0.5).to(dt); b = torch.randn(H, device="cuda")
        for flush in (True, False):
            t = timeit(lambda: ops.linear_fwd(agg, w, x, w, bias=b), iters=20, flush=flush)
            print(f"n={n} {str(dt)[6:]} flush={flush} plain2: {t*1e6:.1f} us  ({t*1e6/(n/128/148):.1f} us per wave)", flush=True)
