#!/usr/bin/env python3
"""GEMM-level accuracy of the 3xTF32 tensor-core path vs the SIMT fp32 kernel vs torch fp32, against fp64."""
import os
import subprocess
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]


def run():
    from deep_fem_uav_wing.gnn import ops

    torch.manual_seed(0)
    n = int(os.environ.get("NROWS", "20000"))
    for H in (64, 128, 256):
        a1 = torch.randn(n, H, device="cuda")
        a2 = torch.randn(n, H, device="cuda")
        w1 = torch.randn(H, H, device="cuda") / H**0.5
        w2 = torch.randn(H, H, device="cuda") / H**0.5
        ref = a1.double() @ w1.double().T + a2.double() @ w2.double().T
        out, _, _, _ = ops.linear_fwd(a1, w1, a2, w2)
        tfp32 = a1 @ w1.T + a2 @ w2.T
        g = torch.randn(n, H, device="cuda")
        dw, _, _ = ops.linear_bwd_weight(g, a1, a2, want_bias=False)
        dref = g.double().T @ a1.double()
        dx = ops.linear_bwd_input(g, w1)
        dxref = g.double() @ w1.double()
        f = lambda x, r: ((x.double() - r).norm() / r.norm()).item()
        m = lambda x, r: ((x.double() - r).abs().max() / r.abs().max()).item()
        print(f"H={H} simt={os.environ.get('DFW_FORCE_SIMT','0')} fwd relL2={f(out, ref):.2e} max={m(out, ref):.2e} | torch fp32 relL2={f(tfp32, ref):.2e} | "
              f"dW relL2={f(dw, dref):.2e} (torch {f(g.T @ a1, dref):.2e}) | dx relL2={f(dx, dxref):.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for simt in ("0", "1"):
            env = dict(os.environ, DFW_FORCE_SIMT=simt)
            subprocess.run([sys.executable, __file__, "child"], env=env)
