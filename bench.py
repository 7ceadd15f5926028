#!/usr/bin/env python3
"""Headline benchmark: GraphSAGE fwd+bwd+AdamW training throughput (meshes/s) on BASELINE.json
config 2 - 200 synthetic 50k-node wing meshes, hidden 128, 4 layers, fp32, batch 4 - at N GPUs,
data-parallel by mesh (weak scaling: every rank steps through its own 4-mesh batches).

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the CPU path (oracle port of the PyG model) on host cores

One JSON line on rank 0.  `value` = device-resident throughput (pre-collated batches, CSR cached);
`e2e` = the same steps through the public API from pinned HOST batches (H2D + on-device CSR build +
fwd/bwd/AdamW + loss.item() per step).  `roofline` = dominant kernel of the step by measured time
(CUDA events on the launching stream, separate instrumented pass over the same steps), algorithmic
bytes or flops per launch / average duration, against MEASURED_PEAKS.json.  `cpu_baseline` = the
oracle on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]

HIDDEN, LAYERS, BATCH, NODES, N_MESHES, DROPOUT = 128, 4, 4, 50000, 200, 0.1
LR, WD = 1e-3, 1e-4  # train_gnn.py:167
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def make_meshes(ids, kind="tri"):
    from deep_fem_uav_wing.gnn import synth

    gen = synth.surface_tri_wing if kind == "tri" else synth.tet_lattice_wing
    out = []
    for i in ids:
        m = gen(NODES, seed=42 + i)
        out.append({k: m[k] for k in ("x", "edge_index", "y", "loss_mask", "num_nodes")})
    return out


def to_data(m):
    from deep_fem_uav_wing.gnn.loader import Data

    return Data(x=torch.from_numpy(m["x"]), edge_index=torch.from_numpy(m["edge_index"]), y=torch.from_numpy(m["y"]),
                loss_mask=torch.from_numpy(m["loss_mask"]))


class ClockSampler:
    """nvidia-smi style clock / throttle-reason samples during the timed region (via NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.ok:
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": float(self.max),
                    "reasons": sorted(self.reasons), "samples": len(self.samples)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}


def timed_region(step_fn, steps, dist_on, device):
    """barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks."""
    import torch.distributed as dist

    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step_fn(i)
    b.record()
    torch.cuda.synchronize(device)
    if dist_on:
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
    if dist_on:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# --------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's PyG model)
# --------------------------------------------------------------------------------------------
def cpu_training_throughput(meshes, steps, warmup, threads=None):
    from oracle.sage_oracle import GraphSAGEModelRef, MaskedMSELossRef

    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(42)
    model = GraphSAGEModelRef(10, HIDDEN, 1, LAYERS, DROPOUT).train()
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD)
    crit = MaskedMSELossRef()
    batches = []
    for s in range(0, len(meshes) - BATCH + 1, BATCH):
        off, xs, eis, ys, ms = 0, [], [], [], []
        for m in meshes[s:s + BATCH]:
            xs.append(torch.from_numpy(m["x"])); eis.append(torch.from_numpy(m["edge_index"]) + off)
            ys.append(torch.from_numpy(m["y"])); ms.append(torch.from_numpy(m["loss_mask"])); off += m["num_nodes"]
        batches.append((torch.cat(xs), torch.cat(eis, 1), torch.cat(ys), torch.cat(ms)))

    def step(i):
        x, ei, y, m = batches[i % len(batches)]
        opt.zero_grad()
        loss = crit(model(x, ei, None), y, m)
        loss.backward()
        opt.step()
        return loss.item()

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return BATCH * steps / dt, dt / steps * 1e3


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"os_cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cpu_model": model}


def run_reference(args, rank, emit):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 1))
    meshes = make_meshes(range(BATCH * min(2, steps)), args.mesh)
    value, ms = cpu_training_throughput(meshes, steps, warmup)
    info = cpu_info()
    sample = f"{warmup} warm-up + {steps} timed fwd+bwd+AdamW steps of one 4x50k-node batch (cycling {len(meshes) // BATCH} batch(es)); oracle port of the PyG model, torch CPU fp32"
    line = {
        "impl": "reference", "metric": "train_meshes_per_sec", "value": value, "unit": "meshes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "meshes/s", "cores": info["torch_threads"], "kind": "port", "sample": sample, **info},
        "e2e": {"value": value, "unit": "meshes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"cfg2: GraphSAGE fwd+bwd+AdamW, {N_MESHES} synthetic {NODES}-node {'surface-triangle' if args.mesh == 'tri' else 'tet-lattice'} "
                    f"wing meshes, hidden {HIDDEN}, {LAYERS} SAGE layers, fp32, dropout {DROPOUT}, batch {BATCH} meshes/step/GPU",
        "hidden": HIDDEN, "layers": LAYERS, "batch_meshes_per_gpu": BATCH, "nodes_per_mesh": NODES, "global_batch_meshes": BATCH * world,
        "parallelism": f"dp{world} by mesh" if world > 1 else "single GPU",
        "l2_policy": "no flush: every step reads a different 4-mesh batch and streams > 1.5 GB of activations (L2 = 126 MB)",
    }


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", default="tri", choices=["tri", "tet"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying CUDA graphs (both arms)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner) write to fd 1 too, so
    # everything else is routed to stderr and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return

    import torch.distributed as dist

    from deep_fem_uav_wing.gnn import ops
    from deep_fem_uav_wing.gnn.ddp import MeshDataParallel
    from deep_fem_uav_wing.gnn.loader import Batch, DataLoader
    from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: the product path has no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or not dist_on, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # ---- data: this rank's shard of the 200 meshes (mesh id = rank + k*world) -----------------
    steps, warmup = args.steps, args.warmup
    n_local = min(N_MESHES // world, BATCH * (steps + warmup))
    n_local = max(BATCH, n_local // BATCH * BATCH)
    ids = [rank + k * world for k in range(n_local)]
    meshes = make_meshes(ids, args.mesh)
    datas = [to_data(m) for m in meshes]
    n_batches = n_local // BATCH

    torch.manual_seed(42)
    model = GraphSAGEModel(10, HIDDEN, 1, LAYERS, DROPOUT).to(device).train()
    ddp = MeshDataParallel(model) if dist_on else None
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD, fused=True, capturable=True)
    crit = MaskedMSELoss()

    def train_step(b, read_loss=False, return_loss=False):
        if ddp is not None:
            ddp.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        out = model(b.x, b.edge_index, b.batch)
        loss = crit(out, b.y, b.loss_mask)
        if ddp is not None:
            ddp.scale_loss(loss, crit.last_count).backward()
            ddp.finish()
        else:
            loss.backward()
        opt.step()
        if return_loss:
            return loss.detach()
        return loss.item() if read_loss else None

    # ---- arm 1: device-resident batches, CSR cached (one-time build per batch) -----------------
    resident = [Batch.from_data_list(datas[i * BATCH:(i + 1) * BATCH]).to(device) for i in range(n_batches)]
    for b in resident:  # one-time CSR builds (forward + transposed) outside the timed region
        ops.get_graph(b.edge_index, b.x.shape[0]).transpose()
    for i in range(warmup):
        train_step(resident[i % n_batches])
    # One CUDA graph per resident batch (captured ON the batch's tensors: no staging copies, the cached CSR stays outside
    # the graph): a step costs one launch, so the device-resident number is not bounded by Python's ~2 ms of launch work
    # per step (--no-graph: kernels launched one by one).
    replays = None
    if not args.no_graph:
        from deep_fem_uav_wing.gnn.graphed import GraphedTrainStep

        rstep = GraphedTrainStep(model, crit, opt, ddp=ddp)
        replays = [rstep.capture_resident(b.x, b.edge_index, b.y, b.loss_mask)[0] for b in resident]
        for i in range(min(3, n_batches)):
            replays[i]()
    sampler = ClockSampler(local_rank)
    sampler.start()
    k0 = ops.LAUNCH_COUNTER["kernels"]
    if replays is not None:
        ms_total = timed_region(lambda i: replays[(warmup + i) % n_batches](), steps, dist_on, device)
    else:
        ms_total = timed_region(lambda i: train_step(resident[(warmup + i) % n_batches]), steps, dist_on, device)
    launches = ops.LAUNCH_COUNTER["kernels"] - k0
    clocks = sampler.stop()
    value = BATCH * steps * world / (ms_total * 1e-3)

    # ---- inference arm (BASELINE metric: "train/infer"): eval-mode forward over the same resident batches ----
    model.eval()
    with torch.no_grad():
        for i in range(3):
            model(resident[i % n_batches].x, resident[i % n_batches].edge_index)
        ms_inf = timed_region(lambda i: model(resident[(warmup + i) % n_batches].x, resident[(warmup + i) % n_batches].edge_index), steps,
                              dist_on, device)
    model.train()
    infer = {"value": BATCH * steps * world / (ms_inf * 1e-3), "unit": "meshes/s", "ms_per_step": ms_inf / steps,
             "nodes_per_sec": BATCH * steps * world * NODES / (ms_inf * 1e-3), "what": "eval-mode forward, device-resident batches, CSR cached"}

    # ---- roofline pass: same steps with per-launch CUDA events ---------------------------------
    prof_steps = min(steps, 8)
    ops.PROFILER = ops.KernelProfiler()
    overlap, ops.OVERLAP_DW = ops.OVERLAP_DW, False  # every kernel alone on its stream while its duration is measured
    for i in range(prof_steps):
        train_step(resident[i % n_batches])
    summ = ops.PROFILER.summary()
    ops.PROFILER = None
    ops.OVERLAP_DW = overlap
    pk = peaks()
    tot_ms = sum(r["ms"] for r in summ.values())
    kernels = {}
    for name, r in summ.items():
        avg_s = r["ms"] * 1e-3 / r["calls"]
        ent = {"calls_per_step": r["calls"] / prof_steps, "avg_us": round(avg_s * 1e6, 2), "share_of_kernel_time": round(r["ms"] / tot_ms, 4),
               "achieved_GBps": round(r["bytes"] / r["calls"] / avg_s / 1e9, 1),
               "hbm_frac": round(r["bytes"] / r["calls"] / avg_s / 1e9 / pk["hbm_gbs"], 4)}
        if r["flops"]:
            ent["achieved_TFLOPs"] = round(r["flops"] / r["calls"] / avg_s / 1e12, 2)
        kernels[name] = ent
    dom = max(summ, key=lambda k: summ[k]["ms"])
    d = kernels[dom]
    # DRAM traffic per launch of the same kernel on the same shapes, from the committed `ncu --set full` captures
    # (tools/ncu_capture_all.sh -> tools/ncu_summary.py --json); null when no capture of that kernel is on file
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "r01_ncu_traffic.json")
    cap = {"aggregate": "r01b_agg_cfg2", "aggregate_bwd": "r01b_aggs_cfg2", "linear_fwd": "r01b_lin_fwd_fp32",
           "linear_bwd_input": "r01b_lin_fwd_fp32", "linear_bwd_weight": "r01b_dw_fp32", "epilogue_bwd": "r01b_epi_bwd_fp32"}
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        for name, ent in kernels.items():
            c = tj.get(cap.get(name, ""))
            if c:
                ent["ncu_dram_bytes_per_launch"] = c["dram_bytes"]
        if cap.get(dom) in tj:
            traffic, traffic_src = tj[cap[dom]]["dram_bytes"], f"profiles/r01_ncu_traffic.json:{cap[dom]}"
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": d["hbm_frac"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": summ[dom]["bytes"] / summ[dom]["calls"], "peak_source": pk["_source"],
                "note": "algorithmic bytes per launch (DESIGN.md) / mean CUDA-event duration; every kernel of the step is listed under `kernels`"}

    # ---- arm 2: end to end through the public API from pinned host memory ----------------------
    e2e, gstep = None, None
    if not args.no_e2e:
        host_loader = DataLoader(datas, batch_size=BATCH, shuffle=False, device=device)
        h2d = sum(int(getattr(datas[j], k).numel() * getattr(datas[j], k).element_size()) for j in range(BATCH)
                  for k in ("x", "edge_index", "y", "loss_mask")) + 8 * (BATCH + 1)

        def epoch_iter():
            while True:
                for b in host_loader:
                    yield b

        it = epoch_iter()
        ops.clear_graph_cache()
        for _ in range(warmup):
            train_step(next(it), read_loss=True)
        # every step's loss is read back on the host inside the timed region; the read of step i happens after
        # step i+1 has been enqueued (one-step software pipeline), so the GPU is never idle while Python launches
        loss_pin = torch.empty(steps, dtype=torch.float32).pin_memory()
        loss_evs, loss_vals = [], []

        use_graph = not args.no_graph
        if use_graph:
            from deep_fem_uav_wing.gnn.graphed import GraphedTrainStep

            gstep = GraphedTrainStep(model, crit, opt, eager_steps=2, ddp=ddp)
            for _ in range(4):  # first calls of the shape run eagerly, then the step is captured
                b = next(it)
                gstep(b.x, b.edge_index, b.y, b.loss_mask)

        def e2e_step(i):
            if use_graph:
                b = next(it)
                loss = gstep(b.x, b.edge_index, b.y, b.loss_mask)
            else:
                loss = train_step(next(it), return_loss=True)
            loss_pin[i:i + 1].copy_(loss.reshape(1).float(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            loss_evs.append(ev)
            if i > 0:
                loss_evs[i - 1].synchronize()
                loss_vals.append(float(loss_pin[i - 1]))
            if i == steps - 1:
                ev.synchronize()
                loss_vals.append(float(loss_pin[i]))

        ms_e2e = timed_region(e2e_step, steps, dist_on, device)
        assert len(loss_vals) == steps and all(np.isfinite(loss_vals)), "e2e losses were not all read back"
        e2e = {"value": BATCH * steps * world / (ms_e2e * 1e-3), "unit": "meshes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / steps, "cuda_graph": bool(use_graph),
               "path": "DataLoader(pinned host Data, device=cuda) -> H2D on a copy stream (1 batch prefetch) -> GraphSAGEModel(x, edge_index, batch) "
                       "incl. on-device CSR build -> MaskedMSELoss -> backward -> AdamW (one CUDA graph per batch shape, gradient all-reduce captured with it) -> loss copied to pinned host memory and read "
                       "(every step, one step behind the launch front)"}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms = cpu_training_throughput(meshes[:BATCH], steps=2, warmup=1)
        info = cpu_info()
        cpu = {"value": v, "unit": "meshes/s", "cores": info["torch_threads"], "kind": "port", "ms_per_step": ms,
               "sample": "1 warm-up + 2 timed fwd+bwd+AdamW steps on the first 4x50k-node batch of this workload (oracle port of the PyG model, torch CPU fp32)",
               **info}

    if rank == 0:
        line = {
            "metric": "train_meshes_per_sec", "value": value, "unit": "meshes/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "nodes_per_sec": value * NODES, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "infer": infer,
            "aggregation": {"kernel": "dfw_sage_aggregate (forward mean, this workload)", "achieved_GBps": kernels.get("aggregate", {}).get("achieved_GBps"),
                            "hbm_frac": kernels.get("aggregate", {}).get("hbm_frac"),
                            "config4": "2M nodes / 27M edges / H=256 bf16: profiles/r01_cfg4_2M_bf16.jsonl (tools/bench_configs.py cfg4)"},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        }
        emit(line)
    if dist_on:
        # the captured steps hold NCCL work: release the graphs, agree that everybody is done, and leave without
        # tearing the communicator down (destroying a communicator that graphs still reference can block)
        gstep = rstep = replays = None
        import gc

        gc.collect()
        torch.cuda.synchronize(device)
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
