"""Callers on either side of the hot path: the train / inference drivers and the GLB writer."""
import json
import os
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "deep-fem-uav-wing_b200"


def test_glb_writer_produces_a_valid_container(tmp_path):
    from deep_fem_uav_wing.gnn.glb import hot_rgb, viridis_rgb, write_glb

    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3]])
    rgb = viridis_rgb(np.array([0.0, 0.5, 1.0, 2.0]))
    assert rgb.tolist()[0] == [68, 1, 84] and rgb.tolist()[2] == [253, 231, 37] and rgb.tolist()[3] == [253, 231, 37]
    assert hot_rgb(np.array([0.0, 1.0])).tolist()[1] == [255, 255, 255]
    out = tmp_path / "m.glb"
    write_glb(out, v, f, rgb)
    raw = out.read_bytes()
    magic, version, total = struct.unpack("<4sII", raw[:12])
    assert magic == b"glTF" and version == 2 and total == len(raw)
    jlen, jtype = struct.unpack("<I4s", raw[12:20])
    doc = json.loads(raw[20:20 + jlen])
    assert jtype == b"JSON" and doc["accessors"][1]["count"] == 4 and doc["accessors"][0]["count"] == 6
    blen, btype = struct.unpack("<I4s", raw[20 + jlen:28 + jlen])
    assert btype == b"BIN\x00" and blen == doc["buffers"][0]["byteLength"]


@pytest.mark.gpu
def test_train_then_inference_scripts_end_to_end(tmp_path):
    from deep_fem_uav_wing.gnn import synth

    for s in range(10):
        synth.write_case_files(synth.surface_tri_wing(400 + 20 * s, seed=s), tmp_path)
    env = dict(os.environ, PYTHONPATH=str(PKG))
    r = subprocess.run([sys.executable, str(PKG / "scripts" / "train_gnn.py"), "--root", str(tmp_path), "--epochs", "3", "--batch-size", "2",
                        "--hidden-channels", "32", "--num-layers", "2"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ck = tmp_path / "checkpoints"
    log = json.loads((ck / "training_log.json").read_text())
    assert {"args", "device", "total_time_s", "best_epoch", "best_val_loss", "test_loss", "test_metrics", "train_log", "completed_at"} <= set(log)
    assert len(log["train_log"]) == 3 and {"epoch", "train_loss", "val_loss", "val_mae_all", "val_mae_masked", "lr", "epoch_time_s"} <= set(log["train_log"][0])
    import torch

    final = torch.load(ck / "final_model.pt", map_location="cpu", weights_only=False)
    assert final["model_config"] == {"in_channels": 10, "hidden_channels": 32, "out_channels": 1, "num_layers": 2, "dropout": 0.1}
    best = torch.load(ck / "best_model.pt", map_location="cpu", weights_only=False)
    assert {"epoch", "model_state_dict", "optimizer_state_dict", "val_loss", "val_metrics", "args"} <= set(best)
    r = subprocess.run([sys.executable, str(PKG / "scripts" / "inference_gnn.py"), "--root", str(tmp_path), "--all", "--checkpoint",
                        "checkpoints/final_model.pt"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    summ = json.loads((ck / "inference_summary.json").read_text())
    assert summ["n_processed"] == 10 and summ["n_successful"] == 10
    case = sorted((tmp_path / "data" / "raw" / "fem").iterdir())[0]
    rep = json.loads((case / "inference_report.json").read_text())
    assert rep["status"] == "success" and {"all_nodes", "masked_nodes"} == set(rep["metrics"])
    assert (case / "wing_pred.glb").read_bytes()[:4] == b"glTF" and (case / "wing_error.glb").exists()
    # the same run from the single-file case caches (gnn/caseio.py): identical reports, byte for byte
    reports = {c.name: (c / "inference_report.json").read_text() for c in (tmp_path / "data" / "raw" / "fem").iterdir()}
    r = subprocess.run([sys.executable, str(PKG / "scripts" / "inference_gnn.py"), "--root", str(tmp_path), "--all", "--pack", "--checkpoint",
                        "checkpoints/final_model.pt"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "packed 10 case(s)" in r.stdout and len(list((tmp_path / "data" / "packed").glob("*.npz"))) == 10
    for name, text in reports.items():
        assert (tmp_path / "data" / "raw" / "fem" / name / "inference_report.json").read_text() == text, name
