#!/usr/bin/env python3
"""Does a locality-improving node order lift the cfg4 aggregation?  (development probe.)
Same tet lattice, same kernel, node ids relabelled: native (thickness fastest), 8x8x8 bricks, Morton (bit-interleaved)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]


def lattice_edges(nx, ny, nz):
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    nid = ((iz * ny + iy) * nx + ix).astype(np.int32)
    parts = []
    for dx, dy, dz in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)):
        a = nid[: nx - dx, : ny - dy, : nz - dz].ravel()
        b = nid[dx:, dy:, dz:].ravel()
        parts.append(np.stack([a, b]))
        parts.append(np.stack([b, a]))
    return np.concatenate(parts, axis=1)  # int32 [2, E], both directions


def orders(nx, ny, nz):
    n = nx * ny * nz
    ids = np.arange(n, dtype=np.int64)
    ix, iy, iz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    out = {"native": None}
    B = 8
    nbx, nby = -(-nx // B), -(-ny // B)
    key = (((iz // B) * nby + iy // B) * nbx + ix // B) * B**3 + ((iz % B) * B + iy % B) * B + ix % B
    out["brick8"] = key
    m = np.zeros(n, dtype=np.int64)
    for b in range(10):
        m |= ((ix >> b) & 1) << (3 * b) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b + 2)
    out["morton"] = m
    return out


def main():
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    s = (n_target / 36.0) ** (1.0 / 3.0)
    nx, ny = max(2, int(round(s))), max(2, int(round(3 * s)))
    nz = max(2, int(round(n_target / (nx * ny))))
    n = nx * ny * nz
    t0 = time.time()
    ei = lattice_edges(nx, ny, nz)
    ords = orders(nx, ny, nz)
    print(json.dumps({"dims": [nx, ny, nz], "N": n, "E": int(ei.shape[1]), "host_s": round(time.time() - t0, 1)}), flush=True)
    import torch

    from deep_fem_uav_wing.gnn import ops

    H, dt = 256, torch.bfloat16
    x = torch.randn(n, H, device="cuda").to(dt)
    ei_d = torch.from_numpy(ei).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    amin = 2 * n * H * 2 + 4 * ei.shape[1] + 4 * (n + 1)
    for name, key in ords.items():
        if key is None:
            e = ei_d.long()
        else:
            relabel = torch.empty(n, dtype=torch.int64, device="cuda")
            relabel[torch.from_numpy(np.argsort(key, kind="stable")).cuda()] = torch.arange(n, device="cuda")
            e = relabel[ei_d.long()]
        g = ops.get_graph(e.contiguous(), n)
        ts = []
        for i in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.aggregate(g.rowptr, g.col, g.inv_deg, x)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        print(json.dumps({"order": name, "aggregate_us": round(t, 1), "hbm_frac_on_A_min": round(amin / (t * 1e-6) / 6542.1e9, 3)}), flush=True)
        del g, e
        ops.clear_graph_cache()


if __name__ == "__main__":
    main()
