#!/usr/bin/env python3
"""Per-phase timeline of ONE mid-grid CTA of the tcgen05 linear kernel (development aid).
Needs a probe build:  make -C deep-fem-uav-wing_b200/csrc clean && make -C deep-fem-uav-wing_b200/csrc EXTRA=-DDFW_TC_PROBE"""
import ctypes, os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, _cabi
dbg = torch.zeros(96, dtype=torch.int64, device="cuda")
_cabi.lib.dfw_tc_set_probe.argtypes = [ctypes.c_void_p]
assert _cabi.lib.dfw_tc_set_probe(dbg.data_ptr()) == 0
n, H = 200000, 128
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(n, H, device="cuda").to(dt); agg = torch.randn(n, H, device="cuda").to(dt)
    w = (torch.randn(H, H, device="cuda") / H**0.5).to(dt); b = torch.randn(H, device="cuda")
    g, be = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    for _ in range(3):
        dbg.zero_()
        ops.linear_fwd(agg, w, x, w, bias=b, ln=(g, be), relu=True, residual=x, save_pre=True, dropout_p=0.1, seed=3)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    t0 = d[0]
    rel = lambda v: (v - t0) if v else None
    print(dt, "ns since the CTA's setup barrier")
    print("   TMA data landed (chunk c):", [rel(v) for v in d[21:41]])
    print("   MMA operands ready       :", [rel(v) for v in d[1:21]])
    print("   kernel entry", rel(d[65]))
    print("   accum_full", rel(d[41]), "combined", rel(d[42]), "pass1", rel(d[43]), "pass2", rel(d[44]), "res_ready", rel(d[45]),
          "pass3", rel(d[46]), "stores_read", rel(d[47]), "exit", rel(d[48]))
    print("   pass 3 per 32-column group (box acquired, TMEM values, math done, staged+store):", [rel(v) for v in d[49:65]])
