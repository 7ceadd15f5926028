"""Host-side graph construction vs fixtures produced by the reference's own dataset.py."""
import json
import re
import tempfile
from pathlib import Path

import numpy as np
import torch

from helpers import canon_edges, load_golden

from deep_fem_uav_wing.gnn import synth
from deep_fem_uav_wing.gnn.dataset import WingStressDataset, _faces_to_edge_index, build_graph_data
from deep_fem_uav_wing.gnn.loader import Batch, Data, DataLoader


def test_faces_to_edge_index_matches_reference():
    g = load_golden("faces_to_edge_index")
    ids = {i + 10: i for i in range(8)}
    assert np.array_equal(canon_edges(_faces_to_edge_index(g["faces"].tolist(), ids)), canon_edges(g["e_box"]))
    got = _faces_to_edge_index(g["faces"].tolist() + [[10, 11, 999]], ids)  # unknown id -> face skipped
    assert np.array_equal(canon_edges(got), canon_edges(g["e_skip"]))
    e = _faces_to_edge_index([], ids)
    assert e.shape == (2, 0) and e.dtype == np.int64
    assert _faces_to_edge_index([[1, 2, 3]], {}).shape == (2, 0)


def _write_case(g, root):
    cid = "goldencase"
    raw = Path(root) / "data" / "raw"
    for sub in ("fem", "mesh", "geometry"):
        (raw / sub / cid).mkdir(parents=True, exist_ok=True)
    n = len(g["in_node_id"])
    np.savez_compressed(raw / "fem" / cid / "surface_results.npz", node_id=g["in_node_id"], pos=g["in_pos"],
                        normal=g["in_normal"], stress_vm=g["in_stress"], disp=np.zeros((n, 3)), loss_mask=g["in_loss_mask"])
    (raw / "mesh" / cid / "boundary_sets.json").write_text(json.dumps({"surf_all_faces": g["in_faces"].tolist()}))
    p = dict(zip(("span_m", "chord_m", "sweep_deg", "thickness_ratio"), map(float, g["in_params"])))
    p["case_id"] = cid
    (raw / "geometry" / cid / "params.json").write_text(json.dumps(p))
    return cid, raw


def test_build_graph_data_matches_reference():
    g = load_golden("build_graph_case")
    with tempfile.TemporaryDirectory() as td:
        cid, raw = _write_case(g, td)
        got = build_graph_data(raw / "fem" / cid / "surface_results.npz", raw / "mesh" / cid / "boundary_sets.json",
                               raw / "geometry" / cid / "params.json")
    for k in ("x", "y", "loss_mask", "pos", "stress_vm_raw", "global_params", "global_params_raw"):
        assert got[k].dtype == g[k].dtype, k
        assert np.array_equal(got[k], g[k]), k  # same numpy expressions: bit-equal
    assert got["edge_index"].dtype == np.int64 and got["edge_index"].shape == g["edge_index"].shape
    assert np.array_equal(canon_edges(got["edge_index"]), canon_edges(g["edge_index"]))
    assert got["case_id"] == "goldencase"


def test_synth_meshes_are_valid_graphs():
    for m, deg_lo, deg_hi in ((synth.surface_tri_wing(2000, seed=3), 5.5, 6.5), (synth.tet_lattice_wing(3000, seed=3), 9.0, 14.0)):
        ei, n = m["edge_index"], m["num_nodes"]
        assert m["x"].shape == (n, 10) and m["x"].dtype == np.float32 and m["y"].shape == (n, 1)
        assert ei.dtype == np.int64 and ei.min() >= 0 and ei.max() < n
        assert not np.any(ei[0] == ei[1])  # no self loops
        key = ei[0] * n + ei[1]
        assert np.unique(key).size == key.size  # no duplicates
        assert np.array_equal(np.sort(key), np.sort(ei[1] * n + ei[0]))  # symmetric
        assert deg_lo <= ei.shape[1] / n <= deg_hi
        assert 0.0 <= m["x"][:, :3].min() and m["x"][:, :3].max() <= 1.0
    a = synth.surface_tri_wing(50000, seed=1)
    assert a["num_nodes"] == 50000
    r = synth.tet_lattice_wing(1000, seed=2, node_order="random")
    assert r["edge_index"].max() < r["num_nodes"]


def test_dataset_and_loader_roundtrip():
    with tempfile.TemporaryDirectory() as td:
        for s in range(8):
            synth.write_case_files(synth.surface_tri_wing(120 + 8 * s, seed=s), Path(td))
        ds = {sp: WingStressDataset(td, split=sp, seed=42) for sp in ("train", "val", "test")}
        assert [len(ds[s]) for s in ("train", "val", "test")] == [5, 1, 2]  # int(8*.7), int(8*.15), rest
        info = json.loads((Path(td) / "data" / "processed" / "gnn" / "split_info_s42.json").read_text())
        np.random.seed(42)
        perm = np.random.permutation(8)
        ids = sorted(f"synthtri_{s:06d}" for s in range(8))
        assert info["train_case_ids"] == [ids[i] for i in perm[:5]]
        loader = DataLoader(ds["train"], batch_size=2, shuffle=False)
        batches = list(loader)
        assert len(batches) == 3 and batches[0].num_graphs == 2 and batches[2].num_graphs == 1
        b = batches[0]
        d0, d1 = ds["train"][0], ds["train"][1]
        assert b.x.shape[0] == d0.num_nodes + d1.num_nodes
        assert torch.equal(b.edge_index[:, : d0.num_edges], d0.edge_index)
        assert torch.equal(b.edge_index[:, d0.num_edges:], d1.edge_index + d0.num_nodes)
        assert torch.equal(b.batch, torch.cat([torch.zeros(d0.num_nodes), torch.ones(d1.num_nodes)]).long())
        assert b.ptr.tolist() == [0, d0.num_nodes, d0.num_nodes + d1.num_nodes]


def test_processed_cache_coexists_with_reference_format_files():
    """The reference writes PyG's collated ``(data, slices)`` tuple to ``data/processed/gnn/{split}_s{seed}.pt``
    (dataset.py:219-224); ours are separate, tagged files, and a foreign payload under OUR name is rebuilt, not crashed on."""
    from deep_fem_uav_wing.gnn.dataset import PROCESSED_FORMAT

    with tempfile.TemporaryDirectory() as td:
        for s in range(4):
            synth.write_case_files(synth.surface_tri_wing(120 + 8 * s, seed=s), Path(td))
        proc = Path(td) / "data" / "processed" / "gnn"
        proc.mkdir(parents=True)
        for sp in ("train", "val", "test"):  # what a project root already processed by the reference holds
            torch.save(({"x": torch.zeros(3, 10)}, {"x": torch.tensor([0, 3])}), proc / f"{sp}_s42.pt")
        ds = WingStressDataset(td, split="train", seed=42)
        assert len(ds) == 2
        assert isinstance(torch.load(proc / "train_s42.pt", weights_only=False), tuple)  # the reference's file is untouched
        ours = torch.load(proc / "train_s42_dfw.pt", weights_only=False)
        assert ours["format"] == PROCESSED_FORMAT and len(ours["records"]) == 2
        torch.save(({"x": torch.zeros(3, 10)}, {}), proc / "train_s42_dfw.pt")  # a foreign payload under our name
        assert len(WingStressDataset(td, split="train", seed=42)) == 2


def test_loader_sharding_is_a_partition():
    data = [Data(x=torch.zeros(3, 10), edge_index=torch.zeros(2, 0, dtype=torch.long), tag=i) for i in range(10)]
    seen = []
    for r in range(4):
        ld = DataLoader(data, batch_size=1, shuffle=True, rank=r, world_size=4, seed=7)
        ld.set_epoch(3)
        seen.append([b.tag[0] for b in ld])
    assert all(len(s) == 3 for s in seen)  # ceil(10/4), wrap-around padding
    flat = [t for s in seen for t in s]
    assert set(flat) == set(range(10))
