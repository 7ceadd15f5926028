"""Graph dataset builder for wing stress prediction - drop-in for the reference's
``src/deep_fem_uav_wing/gnn/dataset.py`` without PyTorch Geometric.

* ``_faces_to_edge_index(faces, node_id_to_idx)``   (``dataset.py:26-63``)
* ``build_graph_data(surface_npz, boundary_sets, params, ...)`` (``dataset.py:66-164``) - same
  keys, dtypes and normalisations; reads the reference's three-file case format.
* ``WingStressDataset(root, split, split_ratio, seed, ...)`` (``dataset.py:168-328``) - same
  directory layout (``data/raw/{fem,mesh,geometry}``, ``data/processed/gnn``), same seeded
  70/15/15 split and ``split_info_s{seed}.json``.  The processed ``.pt`` files hold a plain list
  of tensor dicts instead of PyG's collated ``(data, slices)`` pair.

The edge SET is identical to the reference's; the edge ORDER is canonical (sorted) here, while
the reference emits Python-``set`` iteration order.  The model is order-independent (the CSR
build sorts), and ``tests/test_dataset_golden.py`` compares after canonicalisation.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any

import numpy as np
import torch

from .caseio import load_packed_case, packed_path, read_case
from .loader import Data


def _faces_to_edge_index(faces, node_id_to_idx) -> np.ndarray:
    """Triangle faces (node ids) -> symmetric ``int64 [2, E]`` edge_index (``dataset.py:26-63``).

    Faces that mention an id missing from ``node_id_to_idx`` are skipped (``:43-46``); no faces
    -> ``[2, 0]`` (``:60-61``).  ``node_id_to_idx`` may be a dict or ``(sorted_ids, order)``.
    """
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3) if len(faces) else np.zeros((0, 3), np.int64)
    if f.shape[0] == 0:
        return np.zeros((2, 0), dtype=np.int64)
    if isinstance(node_id_to_idx, dict):
        ids = np.fromiter(node_id_to_idx.keys(), dtype=np.int64, count=len(node_id_to_idx))
        idx = np.fromiter(node_id_to_idx.values(), dtype=np.int64, count=len(node_id_to_idx))
        order = np.argsort(ids, kind="stable")
        sorted_ids, sorted_idx = ids[order], idx[order]
    else:
        sorted_ids, sorted_idx = node_id_to_idx
    if sorted_ids.size == 0:
        return np.zeros((2, 0), dtype=np.int64)
    pos = np.searchsorted(sorted_ids, f)
    pos_c = np.minimum(pos, sorted_ids.size - 1)
    known = sorted_ids[pos_c] == f
    f = sorted_idx[pos_c][known.all(axis=1)]
    if f.shape[0] == 0:
        return np.zeros((2, 0), dtype=np.int64)
    n = int(sorted_idx.max()) + 1
    a = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    lo, hi = a.min(axis=1), a.max(axis=1)
    key = np.unique(lo * n + hi)
    lo, hi = key // n, key % n
    out = np.empty((2, 2 * key.size), dtype=np.int64)
    out[0, 0::2], out[1, 0::2] = lo, hi
    out[0, 1::2], out[1, 1::2] = hi, lo
    return out


def build_graph_data(surface_npz_path: Path, boundary_sets_path: Path, params_path: Path, *,
                     log_scale_stress: bool = True, normalize_pos: bool = True) -> dict[str, Any]:
    """Build graph data from FEM results (``dataset.py:66-164``).

    Returns a dict with ``x [N,10] f32`` (pos_norm | unit normal | 4 scaled globals),
    ``edge_index [2,E] i64``, ``y [N,1] f32`` (= log1p(stress) by default), ``loss_mask [N] bool``,
    ``pos``, ``disp``, ``stress_vm_raw``, ``case_id``, ``global_params``, ``global_params_raw``.
    """
    return _graph_from_raw(read_case(surface_npz_path, boundary_sets_path, params_path), log_scale_stress, normalize_pos)


def build_graph_data_packed(packed_npz_path: Path, *, log_scale_stress: bool = True, normalize_pos: bool = True) -> dict[str, Any]:
    """:func:`build_graph_data` from the single-file case format of ``gnn/caseio.py``: same result,
    bit for bit, without parsing the JSON face list."""
    return _graph_from_raw(load_packed_case(packed_npz_path), log_scale_stress, normalize_pos)


def _graph_from_raw(raw: dict, log_scale_stress: bool, normalize_pos: bool) -> dict[str, Any]:
    params = {**raw["params"], "case_id": raw["case_id"]}
    node_ids = raw["node_id"]
    pos = raw["pos"].astype(np.float32)
    normal = raw["normal"].astype(np.float32)
    stress_vm = raw["stress_vm"].astype(np.float32)
    disp = raw["disp"].astype(np.float32)
    loss_mask = raw["loss_mask"].astype(bool)
    n = len(node_ids)

    ids = np.asarray(node_ids, dtype=np.int64)
    order = np.argsort(ids, kind="stable")
    # duplicate ids: the reference's dict keeps the LAST index (dataset.py:109)
    sorted_ids, sorted_idx = ids[order], order.astype(np.int64)
    if n > 1 and np.any(sorted_ids[1:] == sorted_ids[:-1]):
        keep = np.append(sorted_ids[1:] != sorted_ids[:-1], True)
        sorted_ids, sorted_idx = sorted_ids[keep], sorted_idx[keep]
    edge_index = _faces_to_edge_index(raw["faces"], (sorted_ids, sorted_idx))

    span_m, chord_m = params["span_m"], params["chord_m"]
    sweep_deg, thickness_ratio = params["sweep_deg"], params["thickness_ratio"]
    global_params = np.array(
        [(span_m - 1.0) / 1.0, (chord_m - 0.2) / 0.3, sweep_deg / 30.0, (thickness_ratio - 0.05) / 0.10], dtype=np.float32
    )

    pos_normalized = pos.copy()
    if normalize_pos:
        pos_min, pos_max = pos.min(axis=0), pos.max(axis=0)
        pos_range = pos_max - pos_min
        pos_range[pos_range < 1e-8] = 1.0
        pos_normalized = (pos - pos_min) / pos_range

    norm_lengths = np.linalg.norm(normal, axis=1, keepdims=True)
    norm_lengths[norm_lengths < 1e-8] = 1.0
    normal_normalized = normal / norm_lengths

    x = np.concatenate([pos_normalized, normal_normalized, np.tile(global_params, (n, 1))], axis=1)

    y = stress_vm.copy()
    if log_scale_stress:
        y = np.log1p(y)
    y = y.reshape(-1, 1)

    return {
        "x": x.astype(np.float32),
        "edge_index": edge_index.astype(np.int64),
        "y": y.astype(np.float32),
        "loss_mask": loss_mask,
        "pos": pos.astype(np.float32),
        "disp": disp.astype(np.float32),
        "stress_vm_raw": stress_vm.astype(np.float32),
        "case_id": params["case_id"],
        "global_params": global_params.astype(np.float32),
        "global_params_raw": np.array([span_m, chord_m, sweep_deg, thickness_ratio], dtype=np.float32),
    }


def build_graph_data_device(surface_npz_path: Path, boundary_sets_path: Path, params_path: Path, *, device="cuda",
                            log_scale_stress: bool = True, normalize_pos: bool = True) -> dict[str, Any]:
    """``build_graph_data`` with the graph construction and the feature assembly on the GPU.

    Same keys as :func:`build_graph_data`; ``x``, ``edge_index``, ``y``, ``loss_mask`` and ``pos`` are CUDA tensors
    (``edge_index`` in canonical (dst, src) order with its CSR already registered, so the first forward does not
    rebuild it), the rest stays numpy.  The host only parses the three files and copies the RAW arrays:
    ``dfw_faces_to_csr`` replaces the Python set loop of ``_faces_to_edge_index`` (``dataset.py:26-63``) and
    ``dfw_node_features`` the numpy block of ``dataset.py:129-151``."""
    return _graph_from_raw_device(read_case(surface_npz_path, boundary_sets_path, params_path), device, log_scale_stress,
                                  normalize_pos)


def build_graph_data_device_packed(packed_npz_path: Path, *, device="cuda", log_scale_stress: bool = True,
                                   normalize_pos: bool = True) -> dict[str, Any]:
    """:func:`build_graph_data_device` from the single-file case format (``gnn/caseio.py``): no JSON on the way to the GPU."""
    return _graph_from_raw_device(load_packed_case(packed_npz_path), device, log_scale_stress, normalize_pos)


def graph_from_raw_device(raw: dict, *, device="cuda", log_scale_stress: bool = True, normalize_pos: bool = True) -> dict[str, Any]:
    """Device-side graph build from the raw arrays of ``caseio.read_case`` / ``caseio.load_packed_case`` (a caller that
    also needs the faces, e.g. for the GLB export, reads the case once)."""
    return _graph_from_raw_device(raw, device, log_scale_stress, normalize_pos)


def _graph_from_raw_device(raw: dict, device, log_scale_stress: bool, normalize_pos: bool) -> dict[str, Any]:
    from . import ops

    dev = torch.device(device)
    npz, params = raw, {**raw["params"], "case_id": raw["case_id"]}
    node_ids = np.asarray(npz["node_id"], dtype=np.int64)
    n = len(node_ids)
    if n > 1 and np.unique(node_ids).size != n:  # duplicate ids (last one wins in the reference's dict): host path
        g = _graph_from_raw(raw, log_scale_stress, normalize_pos)
        for k in ("x", "edge_index", "y", "loss_mask", "pos"):
            g[k] = torch.from_numpy(np.asarray(g[k])).to(dev)
        return g
    pos = npz["pos"].astype(np.float32)
    stress_vm = npz["stress_vm"].astype(np.float32)
    faces = np.ascontiguousarray(raw["faces"], dtype=np.int64).reshape(-1, 3)
    span_m, chord_m = params["span_m"], params["chord_m"]
    sweep_deg, thickness_ratio = params["sweep_deg"], params["thickness_ratio"]
    global_params = np.array(
        [(span_m - 1.0) / 1.0, (chord_m - 0.2) / 0.3, sweep_deg / 30.0, (thickness_ratio - 0.05) / 0.10], dtype=np.float32
    )
    pos_d = torch.from_numpy(pos).to(dev, non_blocking=True)
    x, y = ops.node_features(pos_d, torch.from_numpy(npz["normal"].astype(np.float32)).to(dev, non_blocking=True),
                             torch.from_numpy(stress_vm).to(dev, non_blocking=True), global_params,
                             normalize_pos=normalize_pos, log_scale=log_scale_stress)
    _, edge_index = ops.faces_to_graph(torch.from_numpy(faces).to(dev, non_blocking=True), n, torch.from_numpy(node_ids).to(dev))
    return {
        "x": x,
        "edge_index": edge_index,
        "y": y,
        "loss_mask": torch.from_numpy(npz["loss_mask"].astype(bool)).to(dev),
        "pos": pos_d,
        "disp": npz["disp"].astype(np.float32),
        "stress_vm_raw": stress_vm,
        "case_id": params["case_id"],
        "global_params": global_params,
        "global_params_raw": np.array([span_m, chord_m, sweep_deg, thickness_ratio], dtype=np.float32),
    }


def graph_dict_to_data(g: dict) -> Data:
    return Data(
        x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]), y=torch.from_numpy(g["y"]),
        loss_mask=torch.from_numpy(np.asarray(g["loss_mask"], dtype=bool)), pos=torch.from_numpy(g["pos"]),
        disp=torch.from_numpy(g["disp"]) if "disp" in g else None,
        stress_vm_raw=torch.from_numpy(g["stress_vm_raw"]), case_id=g["case_id"],
        global_params=torch.from_numpy(g["global_params"]) if "global_params" in g else None,
        global_params_raw=torch.from_numpy(g["global_params_raw"]) if "global_params_raw" in g else None,
    )


PROCESSED_FORMAT = "dfw_b200/records/v1"  # tag of the processed cache files written by WingStressDataset.process


class WingStressDataset:
    """In-memory dataset of wing graphs (``dataset.py:168-328``), PyG-free.

    Args: ``root`` (directory containing ``data/raw/{fem,mesh,geometry}``), ``split`` in
    {'train','val','test'}, ``split_ratio`` (train, val), ``seed``, ``log_scale_stress``,
    ``normalize_pos``, ``transform``, ``pre_transform``.
    """

    def __init__(self, root, split: str = "train", split_ratio: tuple[float, float] = (0.7, 0.15), seed: int = 42,
                 log_scale_stress: bool = True, normalize_pos: bool = True, transform=None, pre_transform=None):
        self.split, self.split_ratio, self.seed = split, split_ratio, seed
        self.log_scale_stress, self.normalize_pos = log_scale_stress, normalize_pos
        self.transform, self.pre_transform = transform, pre_transform
        self.root = str(Path(root))
        split_idx = {"train": 0, "val": 1, "test": 2}[split]
        if not all(Path(p).exists() for p in self.processed_paths):
            self.process()
        records = self._load_records(self.processed_paths[split_idx])
        if records is None:  # not written by this package (or an older layout): rebuild our own cache files
            self.process()
            records = self._load_records(self.processed_paths[split_idx])
            if records is None:
                raise RuntimeError(f"{self.processed_paths[split_idx]} is not a {PROCESSED_FORMAT} file even after re-processing")
        self._data = [Data(**r) for r in records]

    @staticmethod
    def _load_records(path):
        """Records of one of OUR processed files, or None when the payload is something else (e.g. PyG's collated
        ``(data, slices)`` tuple, which the reference writes to ``train_s{seed}.pt`` in the same directory)."""
        try:
            payload = torch.load(path, weights_only=False)
        except Exception:
            return None
        if isinstance(payload, dict) and payload.get("format") == PROCESSED_FORMAT and isinstance(payload.get("records"), list):
            return payload["records"]
        return None

    @property
    def raw_dir(self) -> str:
        return str(Path(self.root) / "data" / "raw")

    @property
    def processed_dir(self) -> str:
        return str(Path(self.root) / "data" / "processed" / "gnn")

    @property
    def raw_file_names(self) -> list[str]:
        return ["fem", "mesh", "geometry"]

    @property
    def processed_file_names(self) -> list[str]:
        # NOT the reference's ``train_s{seed}.pt`` (dataset.py:219-224): those hold PyG's collated ``(data, slices)`` pair and
        # need torch_geometric to unpickle; ours hold a tagged list of tensor dicts.  Both can live in the same
        # ``data/processed/gnn`` directory without breaking each other.
        return [f"train_s{self.seed}_dfw.pt", f"val_s{self.seed}_dfw.pt", f"test_s{self.seed}_dfw.pt"]

    @property
    def processed_paths(self) -> list[str]:
        return [str(Path(self.processed_dir) / f) for f in self.processed_file_names]

    def __len__(self) -> int:
        return len(self._data)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        d = self._data[i]
        return self.transform(d) if self.transform is not None else d

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def process(self):
        raw = Path(self.raw_dir)
        fem_dir, mesh_dir, geometry_dir = raw / "fem", raw / "mesh", raw / "geometry"
        case_ids = []
        if fem_dir.is_dir():
            for case_dir in sorted(fem_dir.iterdir()):
                if not case_dir.is_dir():
                    continue
                cid = case_dir.name
                if ((fem_dir / cid / "surface_results.npz").exists() and (mesh_dir / cid / "boundary_sets.json").exists()
                        and (geometry_dir / cid / "params.json").exists()):
                    case_ids.append(cid)
        print(f"[Dataset] Found {len(case_ids)} valid cases")

        data_list = []
        for cid in case_ids:
            try:
                src = (fem_dir / cid / "surface_results.npz", mesh_dir / cid / "boundary_sets.json", geometry_dir / cid / "params.json")
                packed = packed_path(self.root, cid)  # single-file cache of the case (gnn/caseio.py), if it is up to date
                if packed.exists() and packed.stat().st_mtime >= max(q.stat().st_mtime for q in src):
                    g = build_graph_data_packed(packed, log_scale_stress=self.log_scale_stress, normalize_pos=self.normalize_pos)
                else:
                    g = build_graph_data(*src, log_scale_stress=self.log_scale_stress, normalize_pos=self.normalize_pos)
                d = graph_dict_to_data(g)
                if self.pre_transform is not None:
                    d = self.pre_transform(d)
                data_list.append(d)
            except Exception as e:  # dataset.py:284-285: bad cases are skipped, not fatal
                print(f"[Dataset] Failed to process {cid}: {e}")
        print(f"[Dataset] Successfully processed {len(data_list)} cases")

        np.random.seed(self.seed)  # dataset.py:290-298
        indices = np.random.permutation(len(data_list))
        n_train = int(len(data_list) * self.split_ratio[0])
        n_val = int(len(data_list) * self.split_ratio[1])
        parts = [indices[:n_train], indices[n_train:n_train + n_val], indices[n_train + n_val:]]
        print(f"[Dataset] Split: train={len(parts[0])}, val={len(parts[1])}, test={len(parts[2])}")

        Path(self.processed_dir).mkdir(parents=True, exist_ok=True)
        for part, path in zip(parts, self.processed_paths):
            torch.save({"format": PROCESSED_FORMAT, "records": [data_list[i].to_dict() for i in part]}, path)

        split_info = {
            "seed": self.seed, "split_ratio": self.split_ratio, "n_total": len(data_list),
            "n_train": len(parts[0]), "n_val": len(parts[1]), "n_test": len(parts[2]),
            "train_case_ids": [data_list[i].case_id for i in parts[0]],
            "val_case_ids": [data_list[i].case_id for i in parts[1]],
            "test_case_ids": [data_list[i].case_id for i in parts[2]],
        }
        p = Path(self.processed_dir) / f"split_info_s{self.seed}.json"
        p.write_text(json.dumps(split_info, indent=2), encoding="utf-8")
        print(f"[Dataset] Saved split info to {p}")
