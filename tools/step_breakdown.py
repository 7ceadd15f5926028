#!/usr/bin/env python3
"""Per-call breakdown of one cfg2 training step (development aid): every C-ABI launch of the step, keyed by
(label, algorithmic bytes) so the SAGE layers, the encoder and the decoder show up as separate lines."""
import json
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth  # noqa: E402
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss  # noqa: E402

PEAK = 6542.1e9


def main():
    hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
    meshes = [synth.surface_tri_wing(50000, seed=42 + i) for i in range(4)]
    x = torch.cat([torch.from_numpy(m["x"]) for m in meshes]).cuda()
    off, eis = 0, []
    for m in meshes:
        eis.append(torch.from_numpy(m["edge_index"]) + off)
        off += m["num_nodes"]
    ei = torch.cat(eis, 1).cuda()
    y = torch.cat([torch.from_numpy(m["y"]) for m in meshes]).cuda()
    mask = torch.cat([torch.from_numpy(m["loss_mask"]) for m in meshes]).cuda()
    model = GraphSAGEModel(10, hidden, 1, 4, dropout=0.1).cuda().train()
    if dtype == torch.bfloat16:
        model = model.to(torch.bfloat16)
        x = x.to(torch.bfloat16)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
    crit = MaskedMSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x, ei, None).float(), y, mask)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    steps = 5
    ops.PROFILER = ops.KernelProfiler()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    recs = ops.PROFILER.records
    ops.PROFILER = None
    out = {}
    for name, ea, eb, nbytes, flops in recs:
        r = out.setdefault((name, nbytes), [0, 0.0])
        r[0] += 1
        r[1] += ea.elapsed_time(eb)
    total = 0.0
    for (name, nbytes), (calls, ms) in sorted(out.items(), key=lambda kv: -kv[1][1]):
        us = ms / calls * 1e3
        total += ms / steps
        print(json.dumps({"kernel": name, "MB": round(nbytes / 1e6, 1), "calls_per_step": calls / steps, "avg_us": round(us, 1),
                          "us_per_step": round(ms / steps * 1e3, 1), "hbm_frac": round(nbytes / (us * 1e-6) / PEAK, 3) if nbytes else None}))
    print(json.dumps({"sum_ms_per_step": round(total, 3), "eager_step_ms_with_profiler": round(a.elapsed_time(b) / steps, 3)}))


if __name__ == "__main__":
    main()
