"""Single-file case format and batched report writer: bit-identical to the reference's three-file path."""
import json
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import load_golden
from test_dataset_golden import _write_case

from deep_fem_uav_wing.gnn import caseio
from deep_fem_uav_wing.gnn.dataset import WingStressDataset, build_graph_data, build_graph_data_packed


def _case_paths(raw, cid):
    return raw / "fem" / cid / "surface_results.npz", raw / "mesh" / cid / "boundary_sets.json", raw / "geometry" / cid / "params.json"


def test_packed_case_round_trip_and_identical_graph():
    g = load_golden("build_graph_case")
    with tempfile.TemporaryDirectory() as td:
        cid, raw = _write_case(g, td)
        src = _case_paths(raw, cid)
        packed = caseio.pack_case(*src, caseio.packed_path(td, cid))
        a, b = caseio.read_case(*src), caseio.load_packed_case(packed)
        assert a.keys() == b.keys() and a["case_id"] == b["case_id"] and a["params"] == b["params"]
        for k in ("node_id", "pos", "normal", "stress_vm", "disp", "loss_mask", "faces"):
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
        assert np.load(packed)["faces"].dtype == np.int32  # half the bytes of the int64 ids, a fraction of the JSON text
        ref, got = build_graph_data(*src), build_graph_data_packed(packed)
        assert ref.keys() == got.keys()
        for k, v in ref.items():
            if isinstance(v, np.ndarray):
                assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k  # bit-identical, edge order included
            else:
                assert got[k] == v, k
        # the reference fixture itself (written by the reference's own dataset.py)
        for k in ("x", "y", "loss_mask", "pos", "stress_vm_raw", "global_params", "global_params_raw"):
            assert np.array_equal(got[k], g[k]), k


def test_pack_dataset_is_incremental_and_the_dataset_prefers_packed_cases(capsys):
    g = load_golden("build_graph_case")
    with tempfile.TemporaryDirectory() as td:
        cid, raw = _write_case(g, td)
        assert caseio.pack_dataset(td) == [caseio.packed_path(td, cid)]
        t0 = caseio.packed_path(td, cid).stat().st_mtime_ns
        assert caseio.pack_dataset(td) == [caseio.packed_path(td, cid)] and caseio.packed_path(td, cid).stat().st_mtime_ns == t0
        (raw / "fem" / "incomplete").mkdir()  # no mesh / geometry files: not a case
        assert len(caseio.pack_dataset(td, overwrite=True)) == 1
        ds = WingStressDataset(td, split="train", split_ratio=(1.0, 0.0))
        assert len(ds) == 1 and torch.equal(ds[0].x, torch.from_numpy(g["x"]))
        # a stale packed file (older than its sources) is ignored in favour of the three-file form
        bad = caseio.load_packed_case(caseio.packed_path(td, cid))
        assert bad["case_id"] == cid


def test_empty_face_list_and_wide_ids():
    with tempfile.TemporaryDirectory() as td:
        raw = Path(td) / "data" / "raw"
        for sub in ("fem", "mesh", "geometry"):
            (raw / sub / "c0").mkdir(parents=True)
        n = 5
        ids = np.arange(n, dtype=np.int64) + 2**33  # node ids beyond int32: faces stay int64 in the packed file
        np.savez(raw / "fem" / "c0" / "surface_results.npz", node_id=ids, pos=np.random.rand(n, 3), normal=np.random.rand(n, 3),
                 stress_vm=np.random.rand(n), disp=np.zeros((n, 3)), loss_mask=np.ones(n, bool))
        (raw / "geometry" / "c0" / "params.json").write_text(json.dumps(
            {"case_id": "c0", "span_m": 1.5, "chord_m": 0.3, "sweep_deg": 10.0, "thickness_ratio": 0.1}))
        for faces in ([], [[int(ids[0]), int(ids[1]), int(ids[2])], [int(ids[2]), int(ids[3]), 7]]):
            (raw / "mesh" / "c0" / "boundary_sets.json").write_text(json.dumps({"surf_all_faces": faces}))
            src = _case_paths(raw, "c0")
            p = caseio.pack_case(*src, Path(td) / "c0.npz")
            a, b = build_graph_data(*src), build_graph_data_packed(p)
            assert np.array_equal(a["edge_index"], b["edge_index"]) and a["edge_index"].shape == (2, 6 if faces else 0)
            assert np.array_equal(a["x"], b["x"])
        assert np.load(Path(td) / "c0.npz")["faces"].dtype == np.int64
        with pytest.raises(ValueError, match="version"):
            z = dict(np.load(Path(td) / "c0.npz"))
            z["pack_version"] = np.int64(99)
            np.savez(Path(td) / "c1.npz", **z)
            caseio.load_packed_case(Path(td) / "c1.npz")


def test_report_writer_batches_reference_schema_reports():
    with tempfile.TemporaryDirectory() as td:
        fem = Path(td) / "fem"
        (fem / "a").mkdir(parents=True)
        (fem / "b").mkdir()
        reports = [{"case_id": c, "status": "success", "metrics": {"all_nodes": {"mae": 1.0}, "masked_nodes": {"mae": 2.0}},
                    "pred_glb": None, "error_glb": None, "pred_stress_range": [0.0, 1.0], "gt_stress_range": [0.0, 2.0],
                    "error_range": [0.0, 1.0]} for c in ("a", "b", "gone")]
        with caseio.ReportWriter(fem, merged_path=Path(td) / "out" / "reports.jsonl") as w:
            for r in reports:
                w.add(r)
            assert not (fem / "a" / "inference_report.json").exists()  # nothing is written inside the loop
        assert w.written == 3
        assert json.loads((fem / "a" / "inference_report.json").read_text()) == reports[0]
        assert json.loads((fem / "b" / "inference_report.json").read_text()) == reports[1]
        lines = (Path(td) / "out" / "reports.jsonl").read_text().splitlines()
        assert [json.loads(x)["case_id"] for x in lines] == ["a", "b", "gone"]
