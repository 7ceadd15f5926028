#!/usr/bin/env python3
"""Train the GraphSAGE surrogate - B200-native counterpart of the reference's ``scripts/train_gnn.py``.

Same CLI flags (``train_gnn.py:113-125``), same loop semantics (``train_epoch`` ``:44-63``, ``evaluate`` ``:66-109``), same
artefacts: ``checkpoints/best_model.pt`` (``:224-231``), ``final_model.pt`` (``:272-283``), ``training_log.json`` (``:255-268``).
Additions: ``--dtype {fp32,bf16}``, ``--root`` (project root holding ``data/raw``), and mesh-level data parallelism
when launched with ``torchrun`` (one process per GPU, NCCL gradient all-reduce overlapped with backward).

    python deep-fem-uav-wing_b200/scripts/train_gnn.py --epochs 100 --batch-size 4
    torchrun --nproc-per-node 8 deep-fem-uav-wing_b200/scripts/train_gnn.py --epochs 100
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from datetime import datetime, timezone
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(PKG_ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.optim as optim  # noqa: E402

from deep_fem_uav_wing.gnn.dataset import WingStressDataset  # noqa: E402
from deep_fem_uav_wing.gnn.ddp import MeshDataParallel  # noqa: E402
from deep_fem_uav_wing.gnn.loader import DataLoader  # noqa: E402
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss, compute_metrics  # noqa: E402


def train_epoch(model, loader, optimizer, criterion, device, ddp=None):
    """One epoch; epoch loss is weighted by ``num_graphs`` (``train_gnn.py:60-63``).  Losses stay on the device and
    are read once per epoch (the reference syncs every step with ``loss.item()``)."""
    model.train()
    losses, counts = [], []
    for data in loader:
        data = data.to(device)
        if ddp is not None:
            ddp.zero_grad()
        else:
            optimizer.zero_grad(set_to_none=True)
        out = model(data.x, data.edge_index, data.batch)
        loss = criterion(out, data.y, data.loss_mask)
        if ddp is not None:
            ddp.scale_loss(loss, criterion.last_count).backward()
            ddp.finish()
        else:
            loss.backward()
        optimizer.step()
        losses.append(loss.detach() * data.num_graphs)
        counts.append(data.num_graphs)
    if not counts:
        return 0.0
    return float(torch.stack(losses).sum().item()) / sum(counts)


@torch.no_grad()
def evaluate(model, loader, criterion, device, log_scale=True):
    """Mean loss and per-batch-averaged metrics (``train_gnn.py:66-109``).

    Same numbers as the reference's loop, but nothing synchronises inside it: the per-batch loss and the 8 metric
    statistics (``dfw_stress_metrics``) stay on the device and ONE copy at the end brings them all back (the
    reference does ``loss.item()`` and three full D2H copies per batch, ``train_gnn.py:80,85``)."""
    from deep_fem_uav_wing.gnn import ops

    model.eval()
    losses, stats, counts = [], [], []
    for data in loader:
        data = data.to(device)
        out = model(data.x, data.edge_index, data.batch)
        losses.append(criterion(out, data.y, data.loss_mask).detach().double().reshape(1))
        stats.append(ops.stress_metrics(out, data.y, data.loss_mask, log_scale=log_scale))
        counts.append(data.num_graphs)
    if not counts:
        zero = {"mae": 0.0, "rmse": 0.0, "max_error": 0.0}
        return 0.0, {"all_nodes": dict(zero), "masked_nodes": dict(zero)}
    host = torch.cat([torch.cat(losses), torch.cat(stats)]).cpu().tolist()  # the only synchronisation
    nb = len(counts)
    loss_v, st = host[:nb], host[nb:]
    avg_loss = sum(l * c for l, c in zip(loss_v, counts)) / sum(counts)
    avg = {}
    for k, o in (("all_nodes", 0), ("masked_nodes", 4)):
        mae = [st[8 * i + o] for i in range(nb)]
        rmse = [st[8 * i + o + 1] for i in range(nb)]
        mx = [st[8 * i + o + 2] for i in range(nb)]
        avg[k] = {"mae": sum(mae) / nb, "rmse": sum(rmse) / nb, "max_error": max(mx)}
    return avg_loss, avg


def main():
    p = argparse.ArgumentParser(description="Train GNN for Wing Stress Prediction (B200-native)")
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--batch-size", type=int, default=4)
    p.add_argument("--lr", type=float, default=1e-3)
    p.add_argument("--weight-decay", type=float, default=1e-4)
    p.add_argument("--hidden-channels", type=int, default=128)
    p.add_argument("--num-layers", type=int, default=4)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--patience", type=int, default=20)
    p.add_argument("--device", type=str, default="auto", help="auto/cuda (there is no CPU path)")
    p.add_argument("--checkpoint-dir", type=str, default="checkpoints")
    p.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    p.add_argument("--root", type=str, default=os.environ.get("DFW_PROJECT_ROOT", str(Path.cwd())))
    args = p.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or args.device == "cpu":
        raise SystemExit("deep_fem_uav_wing (B200 build) is CUDA-only: no CPU fallback exists")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(args.seed)
    log = print if rank == 0 else (lambda *a, **k: None)
    log(f"[Train] Using device: {device} (world size {world})")

    root = Path(args.root)
    ckpt_dir = root / args.checkpoint_dir
    if rank == 0:
        ckpt_dir.mkdir(parents=True, exist_ok=True)
        WingStressDataset(root, split="train", seed=args.seed)  # builds the processed splits once
    if world > 1:
        dist.barrier()
    train_ds, val_ds, test_ds = (WingStressDataset(root, split=s, seed=args.seed) for s in ("train", "val", "test"))
    log(f"[Train] Train: {len(train_ds)}, Val: {len(val_ds)}, Test: {len(test_ds)}")
    # only the fields the step reads cross PCIe (pos, disp, stress_vm_raw, global_params* stay on the host)
    keys = ("x", "edge_index", "y", "loss_mask")
    train_loader = DataLoader(train_ds, batch_size=args.batch_size, shuffle=True, device=device, rank=rank, world_size=world, seed=args.seed,
                              keys=keys)
    val_loader = DataLoader(val_ds, batch_size=args.batch_size, shuffle=False, device=device, keys=keys)
    test_loader = DataLoader(test_ds, batch_size=args.batch_size, shuffle=False, device=device, keys=keys)

    model = GraphSAGEModel(10, args.hidden_channels, 1, args.num_layers, args.dropout).to(device)
    if args.dtype == "bf16":
        model.set_compute_dtype(torch.bfloat16)
    ddp = MeshDataParallel(model) if world > 1 else None
    log(f"[Train] Model parameters: {sum(q.numel() for q in model.parameters()):,}")
    criterion = MaskedMSELoss()
    optimizer = optim.AdamW(model.parameters(), lr=args.lr, weight_decay=args.weight_decay, fused=True)
    scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", patience=10, factor=0.5)

    best_val, patience, train_log, best_epoch = float("inf"), 0, [], 0
    start = time.time()
    for epoch in range(1, args.epochs + 1):
        t0 = time.time()
        train_loader.set_epoch(epoch)
        train_loss = train_epoch(model, train_loader, optimizer, criterion, device, ddp)
        val_loss, val_metrics = evaluate(model, val_loader, criterion, device)
        scheduler.step(val_loss)
        dt, lr = time.time() - t0, optimizer.param_groups[0]["lr"]
        train_log.append({"epoch": epoch, "train_loss": train_loss, "val_loss": val_loss,
                          "val_mae_all": val_metrics["all_nodes"]["mae"], "val_mae_masked": val_metrics["masked_nodes"]["mae"],
                          "val_rmse_all": val_metrics["all_nodes"]["rmse"], "val_rmse_masked": val_metrics["masked_nodes"]["rmse"],
                          "lr": lr, "epoch_time_s": dt})
        log(f"[Epoch {epoch:03d}] Train Loss: {train_loss:.4f} | Val Loss: {val_loss:.4f} | MAE(all/masked): "
            f"{val_metrics['all_nodes']['mae']:.2e}/{val_metrics['masked_nodes']['mae']:.2e} | LR: {lr:.2e} | Time: {dt:.1f}s")
        if val_loss < best_val:
            best_val, patience, best_epoch = val_loss, 0, epoch
            if rank == 0:
                torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                            "val_loss": val_loss, "val_metrics": val_metrics, "args": vars(args)}, ckpt_dir / "best_model.pt")
                log(f"  -> Saved best model to {ckpt_dir / 'best_model.pt'}")
        else:
            patience += 1
            if patience >= args.patience:
                log(f"[Train] Early stopping at epoch {epoch}")
                break
    total = time.time() - start
    log(f"\n[Train] Training completed in {total:.1f}s")
    if world > 1:
        dist.barrier()
    ck = torch.load(ckpt_dir / "best_model.pt", map_location=device, weights_only=False)
    model.load_state_dict(ck["model_state_dict"])
    test_loss, test_metrics = evaluate(model, test_loader, criterion, device)
    if rank == 0:
        log(f"[Train] Final Test Loss: {test_loss:.4f}  masked MAE {test_metrics['masked_nodes']['mae']:.2e} Pa")
        (ckpt_dir / "training_log.json").write_text(json.dumps({
            "args": vars(args), "device": str(device), "total_time_s": total, "best_epoch": best_epoch, "best_val_loss": best_val,
            "test_loss": test_loss, "test_metrics": test_metrics, "train_log": train_log,
            "completed_at": datetime.now(timezone.utc).isoformat()}, indent=2), encoding="utf-8")
        torch.save({"model_state_dict": model.state_dict(),
                    "model_config": {"in_channels": 10, "hidden_channels": args.hidden_channels, "out_channels": 1,
                                     "num_layers": args.num_layers, "dropout": args.dropout},
                    "test_metrics": test_metrics, "completed_at": datetime.now(timezone.utc).isoformat()}, ckpt_dir / "final_model.pt")
        log(f"[Train] Saved final model to {ckpt_dir / 'final_model.pt'}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
