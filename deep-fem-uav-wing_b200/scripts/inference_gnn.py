#!/usr/bin/env python3
"""GNN inference - B200-native counterpart of the reference's ``scripts/inference_gnn.py``.

Same flags (``inference_gnn.py:332-338``) and the same per-case artefacts the Gradio app reads: ``wing_pred.glb``,
``wing_error.glb``, ``inference_report.json`` (``:314-326``) and ``checkpoints/inference_summary.json`` (``:416-423``).
With ``torchrun`` the case list is sharded ``case_ids[rank::world]`` - no communication.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

PKG_ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(PKG_ROOT))

import torch  # noqa: E402

from deep_fem_uav_wing.gnn import caseio  # noqa: E402
from deep_fem_uav_wing.gnn.dataset import build_graph_data, build_graph_data_device, graph_from_raw_device  # noqa: E402
from deep_fem_uav_wing.gnn.glb import hot_rgb, viridis_rgb, write_glb  # noqa: E402
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, compute_metrics  # noqa: E402


def load_model(checkpoint_path: Path, device: torch.device) -> GraphSAGEModel:
    """``inference_gnn.py:42-59``: ``model_config`` if present, else the H=128 / L=4 defaults; strict load; eval."""
    ck = torch.load(checkpoint_path, map_location=device, weights_only=False)
    config = ck.get("model_config", {"in_channels": 10, "hidden_channels": 128, "out_channels": 1, "num_layers": 4, "dropout": 0.1})
    model = GraphSAGEModel(**config).to(device)
    model.load_state_dict(ck["model_state_dict"])
    model.eval()
    return model


def _surface_faces(node_id, faces) -> np.ndarray:
    ids = np.asarray(node_id, dtype=np.int64)
    order = np.argsort(ids, kind="stable")
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    pos = np.minimum(np.searchsorted(ids[order], f), len(ids) - 1)
    ok = (ids[order][pos] == f).all(axis=1)
    return order[pos][ok].astype(np.uint32)


def run_inference(model, case_id: str, device, paths: dict, *, log_scale_stress: bool = True, deform_scale: float = 10.0,
                  writer=None) -> dict:
    fem_dir, mesh_dir, geometry_dir = paths["fem_dir"], paths["mesh_dir"], paths["geometry_dir"]
    npz_p, bs_p, par_p = fem_dir / case_id / "surface_results.npz", mesh_dir / case_id / "boundary_sets.json", geometry_dir / case_id / "params.json"
    for q, name in ((npz_p, "surface_results.npz"), (bs_p, "boundary_sets.json"), (par_p, "params.json")):
        if not q.exists():
            return {"status": "failed", "reason": f"{name} not found"}
    # the case is read ONCE: from its single-file cache when that is up to date (gnn/caseio.py: no JSON face list to parse),
    # else from the reference's three files; graph construction (faces -> CSR) and feature assembly run on the GPU
    # (dfw_faces_to_csr, dfw_node_features)
    packed = paths.get("packed_dir") and paths["packed_dir"] / f"{case_id}.npz"
    if packed and packed.exists() and packed.stat().st_mtime >= max(q.stat().st_mtime for q in (npz_p, bs_p, par_p)):
        raw = caseio.load_packed_case(packed)
    else:
        raw = caseio.read_case(npz_p, bs_p, par_p)
    g = graph_from_raw_device(raw, device=device, log_scale_stress=log_scale_stress, normalize_pos=True)
    x, edge_index, y, loss_mask = g["x"], g["edge_index"], g["y"], g["loss_mask"]
    g["loss_mask"] = loss_mask.cpu().numpy()
    with torch.no_grad():
        pred_log = model(x, edge_index)
    metrics = compute_metrics(pred_log, y, loss_mask, log_scale=log_scale_stress)
    pred_np = pred_log.float().cpu().numpy().flatten()
    pred_stress = np.expm1(pred_np) if log_scale_stress else pred_np
    gt = g["stress_vm_raw"]
    error = np.abs(gt - pred_stress)

    npz = raw
    faces = _surface_faces(raw["node_id"], raw["faces"])
    pred_glb, err_glb = fem_dir / case_id / "wing_pred.glb", fem_dir / case_id / "wing_error.glb"
    ok = len(faces) > 0
    if ok:
        m = g["loss_mask"]
        valid = gt[m] if m is not None and np.any(m) else gt  # unified colour scale with the FEM view (inference_gnn.py:134-141)
        vmin = float(valid.min()) if valid.size else 0.0
        vmax = float(np.percentile(valid, 98)) if valid.size else 1.0
        norm = (np.clip(pred_stress, vmin, vmax) - vmin) / max(1e-12, vmax - vmin)
        write_glb(pred_glb, npz["pos"], faces, viridis_rgb(norm))
        emax = float(error.max())
        write_glb(err_glb, npz["pos"], faces, hot_rgb(error / emax if emax >= 1e-10 else np.zeros_like(error)))
    report = {
        "case_id": case_id, "status": "success" if ok else "partial", "metrics": metrics,
        "pred_glb": str(pred_glb) if ok else None, "error_glb": str(err_glb) if ok else None,
        "pred_stress_range": [float(pred_stress.min()), float(pred_stress.max())],
        "gt_stress_range": [float(gt.min()), float(gt.max())], "error_range": [float(error.min()), float(error.max())],
    }
    if writer is not None:
        writer.add(report)  # written in one go at the end of the run (caseio.ReportWriter)
    else:
        (fem_dir / case_id / "inference_report.json").write_text(json.dumps(report, indent=2), encoding="utf-8")
    return report


def main():
    p = argparse.ArgumentParser(description="GNN Inference for Wing Stress Prediction (B200-native)")
    p.add_argument("--case-id", type=str, default=None)
    p.add_argument("--all", action="store_true")
    p.add_argument("--checkpoint", type=str, default="checkpoints/best_model.pt")
    p.add_argument("--deform-scale", type=float, default=10.0)
    p.add_argument("--device", type=str, default="auto")
    p.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    p.add_argument("--root", type=str, default=os.environ.get("DFW_PROJECT_ROOT", str(Path.cwd())))
    p.add_argument("--pack", action="store_true", help="write/refresh the single-file case caches under data/packed first (gnn/caseio.py)")
    args = p.parse_args()
    if not args.case_id and not args.all:
        p.error("give --case-id or --all")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available() or args.device == "cpu":
        raise SystemExit("deep_fem_uav_wing (B200 build) is CUDA-only: no CPU fallback exists")
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    root = Path(args.root)
    raw = root / "data" / "raw"
    paths = {"fem_dir": raw / "fem", "mesh_dir": raw / "mesh", "geometry_dir": raw / "geometry", "packed_dir": root / "data" / "packed"}
    if args.pack and rank == 0:
        print(f"[Inference] packed {len(caseio.pack_dataset(root))} case(s) under {paths['packed_dir']}")
    model = load_model(root / args.checkpoint, device)
    if args.dtype == "bf16":
        model.set_compute_dtype(torch.bfloat16)
    case_ids = [args.case_id] if args.case_id else sorted(d.name for d in paths["fem_dir"].iterdir() if d.is_dir())
    mine = case_ids[rank::world]
    results = []
    writer = caseio.ReportWriter(paths["fem_dir"])
    for cid in mine:
        r = run_inference(model, cid, device, paths, deform_scale=args.deform_scale, writer=writer)
        results.append(r)
        if r.get("status") in ("success", "partial"):
            mm = r["metrics"]["masked_nodes"]
            print(f"[Inference] {cid}: masked MAE {mm['mae']:.2e} Pa, RMSE {mm['rmse']:.2e} Pa")
        else:
            print(f"[Inference] {cid}: {r.get('status')} ({r.get('reason')})")
    writer.flush()
    good = [r for r in results if r.get("status") in ("success", "partial")]
    print(f"\n[Inference] Completed: {len(good)}/{len(results)} successful")
    if world == 1:  # same schema as the reference (inference_gnn.py:416-423)
        summary = {"n_processed": len(results), "n_successful": len(good), "results": results}
        (root / "checkpoints").mkdir(parents=True, exist_ok=True)
        (root / "checkpoints" / "inference_summary.json").write_text(json.dumps(summary, indent=2), encoding="utf-8")
    else:  # sharded run: every rank writes its part; rank 0's file name is the reference's
        (root / "checkpoints").mkdir(parents=True, exist_ok=True)
        name = "inference_summary.json" if rank == 0 else f"inference_summary.rank{rank}.json"
        (root / "checkpoints" / name).write_text(json.dumps({"n_processed": len(results), "n_successful": len(good), "results": results,
                                                             "rank": rank, "world_size": world}, indent=2), encoding="utf-8")


if __name__ == "__main__":
    main()
