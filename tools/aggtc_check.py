#!/usr/bin/env python3
"""Quick correctness sweep of dfw_sage_aggregate_tc against the gather kernel on mid-size lattices (dev aid)."""
import sys, os, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench
from deep_fem_uav_wing.gnn import ops
for dims in ((20, 40, 80), (30, 60, 160), (38, 114, 200)):
    for order in ("native", "random"):
        ei, pos, n = bench.cfg4_lattice_device(dims, torch.device("cuda"), order)
        for H in (128, 256):
            x = torch.randn(n, H, device="cuda").bfloat16()
            for reorder in ("never", "always"):
                ops._INF_CACHE.clear()
                ig = ops.get_inference_graph(ei, n, pos=pos, reorder=reorder)
                g = ig.graph
                xp = x if ig.order is None else x.index_select(0, ig.order)
                try:
                    got = ops.aggregate_tc(g.plan, g.inv_deg, xp, g.num_edges)
                    torch.cuda.synchronize()
                    ref = ops.aggregate(g.rowptr, g.col, g.inv_deg, xp)
                    err = (got.float() - ref.float()).abs().max().item()
                    print(dims, order, H, reorder, "n", n, "staged", round(g.plan.staged_rows_per_row, 2), "maxerr", err, flush=True)
                except Exception as e:
                    print(dims, order, H, reorder, "FAILED", str(e)[:100], flush=True)
                    sys.exit(1)
