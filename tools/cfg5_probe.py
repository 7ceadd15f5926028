#!/usr/bin/env python3
"""Where does a config-5 launch (16 packed cases -> features + CSR -> forward -> per-case means) spend its time: host issue vs GPU."""
import os, sys, time, json
import numpy as np
import torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth
from deep_fem_uav_wing.gnn.model import GraphSAGEModel

dev = torch.device("cuda", 0)
B, nodes = 16, 20000
ms = [synth.surface_tri_wing(nodes, seed=10 + i) for i in range(B)]
pos = torch.cat([torch.from_numpy(m["pos"]) for m in ms]).to(dev)
nrm = torch.cat([torch.from_numpy(m["normal"]) for m in ms]).to(dev)
off, fs, ptr = 0, [], [0]
for m in ms:
    fs.append(torch.from_numpy(m["faces"].astype(np.int64)) + off); off += m["num_nodes"]; ptr.append(off)
faces = torch.cat(fs).to(dev)
gp = torch.rand(B, 4, device=dev)
cptr = torch.tensor(ptr, dtype=torch.int64, device=dev)
res = {}
for dt in (torch.float32, torch.bfloat16):
    model = GraphSAGEModel(10, 128, 1, 4).to(dev).eval().set_compute_dtype(dt)
    def launch():
        t0 = time.perf_counter()
        x, _ = ops.node_features_batched(pos, nrm, None, gp, cptr, nodes)
        t1 = time.perf_counter()
        g, ei = ops.faces_to_graph(faces, off)
        t2 = time.perf_counter()
        with torch.no_grad():
            out = model(x, ei)
        t3 = time.perf_counter()
        means = out.float().view(B, -1).mean(dim=1)
        ops.clear_graph_cache()
        t4 = time.perf_counter()
        return (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
    for _ in range(3): launch()
    torch.cuda.synchronize()
    host = np.zeros(4); n = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): host += np.array(launch())
    b.record(); torch.cuda.synchronize()
    res[str(dt)] = {"gpu_ms_per_launch": a.elapsed_time(b) / n, "host_ms_features,graph,forward,means": (host / n * 1e3).round(3).tolist()}
print(json.dumps(res))
