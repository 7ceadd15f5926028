#!/usr/bin/env python3
"""Per-kernel CUDA-event timings on config-shaped inputs (development aid; prints JSON lines)."""
import json
import os
import sys
import time

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth  # noqa: E402
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss  # noqa: E402

PEAK = 6542.1e9
flush_buf = None


def flush_l2():
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_buf.zero_()


def timeit(fn, iters=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, t, bytes_=None, flops=None, **kw):
    r = {"name": name, "us": round(t * 1e6, 2)}
    if bytes_:
        r["GBps"] = round(bytes_ / t / 1e9, 1)
        r["hbm_frac"] = round(bytes_ / t / PEAK, 3)
    if flops:
        r["TFLOPs"] = round(flops / t / 1e12, 2)
    r.update(kw)
    print(json.dumps(r), flush=True)


def agg_case(name, mesh, H, dtype):
    n = mesh["num_nodes"]
    ei = torch.from_numpy(mesh["edge_index"]).cuda()
    E = ei.shape[1]
    g = ops.get_graph(ei, n)
    x = torch.randn(n, H, device="cuda").to(dtype)
    b = 4 if dtype == torch.float32 else 2
    amin = 2 * n * H * b + 4 * E + 4 * (n + 1)
    t = timeit(lambda: ops.aggregate(g.rowptr, g.col, g.inv_deg, x))
    report(f"aggregate/{name}", t, amin, N=n, E=E, H=H, dtype=str(dtype), A_min_MB=round(amin / 1e6, 1))
    t = timeit(lambda: ops.csr_build_raw(ei, n, want_perm=False), iters=5)
    report(f"csr_build/{name}", t, 16 * E, N=n, E=E)
    return g, x


def main():
    what = sys.argv[1:] or ["cfg2", "cfg4", "model"]
    torch.manual_seed(0)
    if "cfg2" in what:
        for kind, gen in (("tri", synth.surface_tri_wing), ("tet", synth.tet_lattice_wing)):
            mesh = gen(200000, seed=42)
            g, x = agg_case(f"cfg2-batch4-{kind}", mesh, 128, torch.float32)
        n, H = x.shape
        w = torch.randn(H, H, device="cuda") / H**0.5
        bias = torch.randn(H, device="cuda")
        gam, bet = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
        agg = ops.aggregate(g.rowptr, g.col, g.inv_deg, x)
        t = timeit(lambda: ops.linear_fwd(agg, w, x, w, bias=bias, ln=(gam, bet), relu=True, residual=x, save_pre=True))
        report("linear_fwd/H128-fp32-ffma", t, 5 * n * H * 4, 4 * n * H * H)
        gy = torch.randn(n, H, device="cuda")
        t = timeit(lambda: ops.linear_bwd_input(gy, w, row_scale=g.inv_deg))
        report("linear_bwd_input/H128-fp32", t, 2 * n * H * 4, 2 * n * H * H)
        t = timeit(lambda: ops.linear_bwd_weight(gy, agg, x))
        report("linear_bwd_weight/H128-fp32", t, 3 * n * H * 4, 4 * n * H * H)
        pre = torch.randn(n, H, device="cuda")
        stats = torch.rand(n, 2, device="cuda") + 0.5
        t = timeit(lambda: ops.epilogue_bwd(gy, n, H, gy, pre=pre, stats=stats, ln=(gam, bet), relu=True))
        report("epilogue_bwd/H128-fp32", t, 3 * n * H * 4)
    if "cfg4" in what:
        for order in ("native", "random"):
            t0 = time.time()
            mesh = synth.tet_lattice_wing(2_000_000, seed=42, node_order=order, shuffle_edges=False)
            print(json.dumps({"gen_s": round(time.time() - t0, 1), "dims": mesh["lattice_dims"]}), flush=True)
            g, x = agg_case(f"cfg4-{order}", mesh, 256, torch.bfloat16)
            del g, x
            ops.clear_graph_cache()
    if "model" in what:
        meshes = [synth.surface_tri_wing(50000, seed=42 + i) for i in range(4)]
        x = torch.cat([torch.from_numpy(m["x"]) for m in meshes]).cuda()
        off, eis = 0, []
        for m in meshes:
            eis.append(torch.from_numpy(m["edge_index"]) + off)
            off += m["num_nodes"]
        ei = torch.cat(eis, 1).cuda()
        y = torch.cat([torch.from_numpy(m["y"]) for m in meshes]).cuda()
        mask = torch.cat([torch.from_numpy(m["loss_mask"]) for m in meshes]).cuda()
        model = GraphSAGEModel(10, 128, 1, 4, dropout=0.1).cuda().train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
        crit = MaskedMSELoss()

        def step():
            opt.zero_grad(set_to_none=True)
            loss = crit(model(x, ei, None), y, mask)
            loss.backward()
            opt.step()

        t = timeit(step, iters=10, flush=False)
        report("train_step/cfg2-batch4-tri-H128-L4", t, meshes_per_s=round(4 / t, 1))
        model.eval()
        with torch.no_grad():
            t = timeit(lambda: model(x, ei), iters=10, flush=False)
        report("infer/cfg2-batch4-tri-H128-L4", t, meshes_per_s=round(4 / t, 1))
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        host = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        report("train_step_host_enqueue", host)


if __name__ == "__main__":
    main()
