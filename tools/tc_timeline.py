#!/usr/bin/env python3
import os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["DFW_TC_DEBUG"] = str(dbg.data_ptr())
from deep_fem_uav_wing.gnn import ops
n, H = 200000, 128
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(n, H, device="cuda").to(dt); agg = torch.randn(n, H, device="cuda").to(dt)
    w = (torch.randn(H, H, device="cuda") / H**0.5).to(dt); b = torch.randn(H, device="cuda")
    g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    for _ in range(3):
        dbg.zero_()
        ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=True)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    t0 = d[66]
    rel = lambda v: v - t0 if v else None
    print(dt, "start->", "tma issue:", [rel(v) for v in d[0:8]])
    print("   conv start:", [rel(v) for v in d[16:24]])
    print("   conv done :", [rel(v) for v in d[32:40]])
    print("   mma start :", [rel(v) for v in d[48:56]])
    print("   accum_full:", rel(d[64]), " epilogue done:", rel(d[65]))
    print("   epi stamps: tmem_loaded", rel(d[70]), "bias_done", rel(d[71]), "pre_stored", rel(d[72]), "ln_relu_done", rel(d[73]), "out_acquired", rel(d[74]), "res_ready", rel(d[75]), "out_staged", rel(d[76]), "out_issued", rel(d[77]))
