#!/usr/bin/env python3
"""Host time to get one case from disk to raw arrays: the reference's three files (JSON face list) vs the packed
single-file form (gnn/caseio.py).  CPU only; prints one JSON line."""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import caseio, synth  # noqa: E402
from deep_fem_uav_wing.gnn.dataset import build_graph_data, build_graph_data_packed  # noqa: E402


def med(fn, n=7):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2] * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    m = synth.surface_tri_wing(n, seed=42)
    nn = m["num_nodes"]
    ids = np.arange(nn, dtype=np.int64) + 1
    with tempfile.TemporaryDirectory() as td:
        raw = Path(td) / "data" / "raw"
        for sub in ("fem", "mesh", "geometry"):
            (raw / sub / "c").mkdir(parents=True)
        rng = np.random.default_rng(0)
        np.savez(raw / "fem" / "c" / "surface_results.npz", node_id=ids, pos=m["pos"], normal=rng.standard_normal((nn, 3)).astype(np.float32),
                 stress_vm=np.expm1(m["y"][:, 0]), disp=np.zeros((nn, 3), np.float32), loss_mask=m["loss_mask"])
        faces = (m["faces"] + 1).tolist()
        (raw / "mesh" / "c" / "boundary_sets.json").write_text(json.dumps({"surf_all_faces": faces}))
        (raw / "geometry" / "c" / "params.json").write_text(json.dumps({"case_id": "c", "span_m": 1.5, "chord_m": 0.3, "sweep_deg": 10.0, "thickness_ratio": 0.1}))
        src = (raw / "fem" / "c" / "surface_results.npz", raw / "mesh" / "c" / "boundary_sets.json", raw / "geometry" / "c" / "params.json")
        packed = caseio.pack_case(*src, caseio.packed_path(td, "c"))
        a, b = build_graph_data(*src), build_graph_data_packed(packed)
        assert all(np.array_equal(a[k], b[k]) for k in a if isinstance(a[k], np.ndarray))
        out = {"config": "case I/O (8f-3)", "N": nn, "F": len(faces),
               "three_files_bytes": sum(p.stat().st_size for p in src), "packed_bytes": packed.stat().st_size,
               "read_case_three_files_ms": round(med(lambda: caseio.read_case(*src)), 2),
               "load_packed_case_ms": round(med(lambda: caseio.load_packed_case(packed)), 2),
               "build_graph_data_three_files_ms": round(med(lambda: build_graph_data(*src)), 2),
               "build_graph_data_packed_ms": round(med(lambda: build_graph_data_packed(packed)), 2),
               "host": f"{os.cpu_count()} cores (build container, no GPU)"}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
