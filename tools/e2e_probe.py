#!/usr/bin/env python3
"""Dev probe: end-to-end step time (loader -> graphed step) with the CSR built inside the step vs prefetched by the loader."""
import os, sys, time, torch
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench
from deep_fem_uav_wing.gnn.graphed import GraphedTrainStep
from deep_fem_uav_wing.gnn.loader import DataLoader
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss
datas = [bench.to_data(m) for m in bench.make_meshes(range(40), "tri")]
for bg in (False, True):
    torch.manual_seed(0)
    model = GraphSAGEModel(10, 128, 1, 4, 0.1).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
    crit = MaskedMSELoss()
    gstep = GraphedTrainStep(model, crit, opt, eager_steps=2)
    ld = DataLoader(datas, batch_size=4, shuffle=False, device="cuda", build_graph=bg)
    for ep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
        for b in ld:
            gstep(b.x, b.edge_index, b.y, b.loss_mask, graph=getattr(b, "graph", None)); n += 1
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = torch.cuda.memory_stats()
        print(f"build_graph={bg} epoch {ep}: {dt / n * 1e3:.2f} ms per step; cudaMalloc calls so far {st['num_device_alloc']}, reserved {st['reserved_bytes.all.current'] >> 20} MB", flush=True)
