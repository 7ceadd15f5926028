#!/usr/bin/env python3
"""Dev probe: throughput of the prefetching DataLoader alone (no training), with and without build_graph."""
import os, sys, time, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench
from deep_fem_uav_wing.gnn.loader import DataLoader
datas = [bench.to_data(m) for m in bench.make_meshes(range(40), "tri")]
for bg in (False, True):
    ld = DataLoader(datas, batch_size=4, shuffle=False, device="cuda", build_graph=bg)
    for ep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
        for b in ld:
            n += 1
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"build_graph={bg} epoch {ep}: {dt / n * 1e3:.2f} ms per batch", flush=True)
