#!/usr/bin/env python3
"""Where does the tensor-core block aggregation start to pay?  gather vs block kernel on batch-sized bf16 graphs (dev aid)."""
import json, os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench
from deep_fem_uav_wing.gnn import ops
from deep_fem_uav_wing.gnn.loader import Batch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, iters=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return round(ts[len(ts) // 2], 1)
for kind, nm in (("tri", 1), ("tri", 4), ("tet", 4), ("tri", 16)):
    bench.NODES = 50000 if nm <= 4 else 20000
    b = Batch.from_data_list([bench.to_data(m) for m in bench.make_meshes(range(nm), kind)]).to("cuda")
    n = b.x.shape[0]
    g = ops.get_graph(b.edge_index, n)
    plan = ops.build_agg_plan(g.rowptr, g.col, n); plan.check()
    for H in (128, 256):
        x = torch.randn(n, H, device="cuda").bfloat16()
        print(json.dumps({"graph": f"{nm} x {bench.NODES} {kind}", "N": n, "H": H, "staged_rows_per_row": round(plan.staged_rows_per_row, 2),
                          "gather_us": t(lambda: ops.aggregate(g.rowptr, g.col, g.inv_deg, x)), "block_us": t(lambda: ops.aggregate_tc(plan, g.inv_deg, x, g.num_edges))}), flush=True)
