#!/usr/bin/env python3
"""Per-kernel time of the config-4 forward (2M-node tet mesh, H=256, bf16, L=3) with the ops.KernelProfiler."""
import json, os, sys, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth
from deep_fem_uav_wing.gnn.model import GraphSAGEModel
mesh = synth.tet_lattice_wing(2_000_000, seed=42, node_order="native", shuffle_edges=False)
n = mesh["num_nodes"]
ei = torch.from_numpy(mesh["edge_index"]).cuda()
x = torch.from_numpy(mesh["x"]).cuda()
torch.manual_seed(0)
model = GraphSAGEModel(10, 256, 1, 3).cuda().eval().set_compute_dtype(torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        model(x, ei)
    torch.cuda.synchronize()
    ops.PROFILER = ops.KernelProfiler()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        model(x, ei)
    b.record()
    summ = ops.PROFILER.summary()
    ops.PROFILER = None
print("forward ms", a.elapsed_time(b) / 5)
for k, r in summ.items():
    print(f"{k:20s} calls/fwd {r['calls'] / 5:4.1f}  avg {r['ms'] / r['calls'] * 1e3:8.1f} us  total/fwd {r['ms'] / 5 * 1e3:8.1f} us  {r['bytes'] / r['calls'] / (r['ms'] / r['calls'] * 1e-3) / 1e9:7.1f} GB/s")
