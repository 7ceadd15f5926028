#!/usr/bin/env python3
"""Dev probe: config-4 lattice in RANDOM numbering - time the k-d relabelled block aggregation before and after the slow paths."""
import json, os, sys
import torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench
from deep_fem_uav_wing.gnn import ops
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, iters=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)
res = {}
for order in (sys.argv[1:] or ["random"]):
    ei, pos_n, n = bench.cfg4_lattice_device(bench.CFG4_DIMS, dev, order)
    x = torch.randn(n, 256, device=dev).to(torch.bfloat16)
    paths = ops.cfg4_aggregation_paths(ei, n, pos_n, x)
    r = {}
    r["kd_first"] = t(paths["tensor_core_blocks_kd_order"][0])
    r["gather"] = t(paths["gather"][0])
    r["kd_after_gather"] = t(paths["tensor_core_blocks_kd_order"][0])
    r["info"] = paths["tensor_core_blocks_kd_order"][1]
    ig = ops.get_inference_graph(ei, n, pos=pos_n, reorder="auto")
    pl = ig.graph.plan
    r["plan"] = {"src": int(pl.plan_src.numel()), "slot": int(pl.plan_slot.numel())}
    res[order] = r
    del paths, x, ei
    ops.clear_graph_cache(); torch.cuda.empty_cache()
print(json.dumps(res))
