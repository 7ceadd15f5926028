"""Tensor-core block-sparse aggregation (GPU, through the C ABI): the device-built block plan against the numpy design oracle
(``oracle/blocked_agg_oracle.py``) bit for bit, and ``dfw_sage_aggregate_tc`` against the plain-C CSR aggregation
(``oracle/csr_oracle.c``) - mean and sum, H in {64, 128, 256}, ragged last blocks, isolated rows, duplicate edges, long rows,
random graphs with poor locality (many chunks per block)."""
import numpy as np
import pytest
import torch

from helpers import TOL_BF16, record, rel_max
from oracle import csr_aggregate_c, csr_oracle_c
from oracle.blocked_agg_oracle import build_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from deep_fem_uav_wing.gnn import ops as _ops

    return _ops


def _graphs():
    from deep_fem_uav_wing.gnn import synth

    rng = np.random.default_rng(0)
    out = {}
    out["tet_5k"] = (synth.tet_lattice_wing(5000, seed=1)["edge_index"], None)
    m = synth.surface_tri_wing(3000, seed=2)
    out["tri_3k"] = (m["edge_index"], m["num_nodes"])
    n = 1000
    ei = rng.integers(0, n, size=(2, 9000)).astype(np.int64)  # random: duplicates, self loops, ~14 chunks per block
    ei[1, :300] = 5  # one long row
    ei = ei[:, ei[1] != 77]  # one isolated row
    out["random_1k"] = (ei, n)
    out["single_row"] = (np.array([[0, 0, 0], [0, 0, 0]], dtype=np.int64), 1)
    out["no_edges"] = (np.zeros((2, 0), dtype=np.int64), 300)
    out["ragged_129"] = (rng.integers(0, 129, size=(2, 700)).astype(np.int64), 129)
    return out


@pytest.mark.parametrize("name", ["tet_5k", "tri_3k", "random_1k", "single_row", "no_edges", "ragged_129"])
def test_block_plan_matches_design_oracle(ops, name):
    ei, n = _graphs()[name]
    n = int(ei.max()) + 1 if n is None else n
    rowptr, col, _, _ = csr_oracle_c(ei, n)
    plan = ops.build_agg_plan(torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda(), n)
    assert plan.check()
    blk_ptr, blk_src, _ = build_blocks(rowptr, col, 128)
    meta = plan.blk_meta.cpu().numpy().reshape(-1, 4)
    src = plan.plan_src.cpu().numpy()
    rec = plan.plan_rec.cpu().numpy().view(np.uint16).reshape(-1, 136)
    slots = plan.plan_slot.cpu().numpy().view(np.uint16)
    nb = (n + 127) // 128
    assert meta.shape[0] >= nb
    total = 0
    for b in range(nb):
        s_off, S, t_off, nent = (int(v) for v in meta[b])
        r0, r1 = 128 * b, min(128 * b + 128, n)
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        # slots: the block's own rows hold 0 .. 127 (slot = row - r0, used or not), the distinct halo sources follow ascending
        distinct = blk_src[blk_ptr[b]:blk_ptr[b + 1]].astype(np.int64)  # the design oracle's ascending distinct sources
        halo = distinct[(distinct < r0) | (distinct >= r1)]
        assert S == 128 + halo.size and s_off % 64 == 0 and t_off % 8 == 0
        assert np.array_equal(src[s_off:s_off + 128], np.minimum(r0 + np.arange(128), n - 1))
        assert np.array_equal(src[s_off + 128:s_off + S], halo)
        pad = (S + 63) // 64 * 64
        assert np.all(src[s_off + S:s_off + pad] == (halo[-1] if halo.size else min(r0 + 127, n - 1)))
        # adjacency entries: one per distinct (slot, row) pair with its multiplicity, ordered by (slot, row)
        cols_b = col[e0:e1].astype(np.int64)
        own = (cols_b >= r0) & (cols_b < r1)
        slot_b = np.where(own, cols_b - r0, 128 + np.searchsorted(halo, cols_b))
        rows_of_edges = np.repeat(np.arange(r1 - r0), np.diff(rowptr[r0:r1 + 1]))
        pairs, counts = np.unique(np.stack([slot_b, rows_of_edges]), axis=1, return_counts=True) if e1 > e0 else (np.zeros((2, 0), np.int64), np.zeros(0, np.int64))
        want = ((counts - 1) << 13 | pairs[1] << 6 | (pairs[0] & 63)).astype(np.uint16)
        assert nent == want.size and rec[b, 0] == S and rec[b, 1] == nent
        assert np.array_equal(slots[t_off:t_off + nent], want)
        nch = pad // 64
        cptr = np.searchsorted(pairs[0], 64 * np.arange(nch + 1))
        assert np.array_equal(rec[b, 2:3 + nch], cptr.astype(np.uint16))
        total += S
    assert abs(plan.staged_rows_per_row * n - total) < 0.5


@pytest.mark.parametrize("h", [64, 128, 256])
@pytest.mark.parametrize("name", ["tet_5k", "tri_3k", "random_1k", "single_row", "no_edges", "ragged_129"])
def test_aggregate_tc_matches_c_oracle(ops, name, h):
    ei, n = _graphs()[name]
    n = int(ei.max()) + 1 if n is None else n
    rng = np.random.default_rng(h)
    x = torch.from_numpy(rng.standard_normal((n, h)).astype(np.float32)).bfloat16()
    rowptr, col, _, inv = csr_oracle_c(ei, n)
    g = ops.get_graph(torch.from_numpy(ei).cuda(), n)
    plan = ops.build_agg_plan(g.rowptr, g.col, n)
    assert plan.check()
    for scale, inv_np in ((g.inv_deg, inv), (None, None)):  # mean, then plain sum
        ref = torch.from_numpy(csr_aggregate_c(rowptr, col, inv_np, x.float().numpy()))
        got = ops.aggregate_tc(plan, scale, x.cuda(), g.num_edges)
        assert got.dtype == torch.bfloat16 and got.shape == (n, h)
        gf = got.float().cpu()
        assert torch.isfinite(gf).all()
        err = rel_max(gf, ref) if ref.abs().max() > 0 else float(gf.abs().max())
        mism = int((got.cpu() != ref.bfloat16()).sum())
        record("aggregate_tc_vs_c_oracle", graph=name, H=h, mean=scale is not None, rel_max=err, bf16_mismatches=mism, elements=n * h)
        assert err < TOL_BF16
        # association order of fp32 additions only: at most one bf16 ulp of the result's magnitude
        assert (gf - ref).abs().max().item() <= 2.0 ** -7 * max(ref.abs().max().item(), 1e-30)
        assert torch.equal(got, ops.aggregate_tc(plan, scale, x.cuda(), g.num_edges))  # deterministic
        same = ops.aggregate(g.rowptr, g.col, scale, x.cuda())
        assert rel_max(gf, same.float().cpu()) < 2.0 ** -7 if ref.abs().max() > 0 else True


def test_aggregate_tc_rejects_what_it_cannot_take(ops):
    from deep_fem_uav_wing.gnn import _cabi

    ei = np.array([[0, 1], [1, 0]], dtype=np.int64)
    g = ops.get_graph(torch.from_numpy(ei).cuda(), 2)
    plan = ops.build_agg_plan(g.rowptr, g.col, 2)
    with pytest.raises(TypeError):
        ops.aggregate_tc(plan, None, torch.zeros(2, 64, device="cuda"))
    with pytest.raises(_cabi.DfwError, match="64, 128 or 256"):
        ops.aggregate_tc(plan, None, torch.zeros(2, 96, device="cuda", dtype=torch.bfloat16))
    # a block with more edges than the plan holds is reported, not mis-computed
    n = 200
    hub = np.stack([np.arange(5000) % n, np.zeros(5000, dtype=np.int64)]).astype(np.int64)
    gh = ops.get_graph(torch.from_numpy(hub).cuda(), n)
    ph = ops.build_agg_plan(gh.rowptr, gh.col, n)
    assert not ph.check()
    # so is an edge repeated more often than the 3-bit multiplicity field holds
    rep = np.stack([np.zeros(9, dtype=np.int64), np.ones(9, dtype=np.int64)])
    gr = ops.get_graph(torch.from_numpy(rep).cuda(), 2)
    assert not ops.build_agg_plan(gr.rowptr, gr.col, 2).check()
    rep8 = rep[:, :8].copy()
    g8 = ops.get_graph(torch.from_numpy(rep8).cuda(), 2)
    p8 = ops.build_agg_plan(g8.rowptr, g8.col, 2)
    assert p8.check()
    out8 = ops.aggregate_tc(p8, None, torch.ones(2, 64, device="cuda", dtype=torch.bfloat16))
    assert out8[1].float().eq(8).all() and out8[0].float().eq(0).all()


def test_large_mesh_inference_path_relabels_nodes_and_matches_the_plain_path(ops, monkeypatch):
    """GraphSAGEModel bf16 inference on a randomly numbered tet mesh: the block plan + k-d relabelling path (input rows
    permuted once, output restored) must give the plain path's predictions in the CALLER's node order."""
    from deep_fem_uav_wing.gnn import synth
    from deep_fem_uav_wing.gnn.model import GraphSAGEModel

    mesh = synth.tet_lattice_wing(120000, seed=3, node_order="random")
    n = mesh["num_nodes"]
    x, ei = torch.from_numpy(mesh["x"]).cuda(), torch.from_numpy(mesh["edge_index"]).cuda()
    torch.manual_seed(5)
    model = GraphSAGEModel(10, 128, 1, 2, dropout=0.0).cuda().eval().set_compute_dtype(torch.bfloat16)
    model.mesh_plan = "always"  # ('auto' prepares a mesh from its second forward on: covered at the end of this test)
    monkeypatch.setattr(ops, "TC_AGG_MIN_NODES", 1000)
    monkeypatch.setattr(ops, "TC_AGG_MIN_REUSE", 0.0)  # (the policy would leave a 120k-node mesh with ~4.5 staged rows per row to the gather kernel)
    tc_calls = []
    real_tc = ops.aggregate_tc
    monkeypatch.setattr(ops, "aggregate_tc", lambda *a, **k: (tc_calls.append(1), real_tc(*a, **k))[1])
    with torch.no_grad():
        out_plan = model(x, ei)
    assert len(tc_calls) == 2, "every SAGE layer of the large-mesh path must run the tensor-core block aggregation"
    ig = ops.get_inference_graph(ei, n, pos=x[:, :3], reorder="auto")
    assert ig.order is not None and ig.graph.plan.usable
    assert ig.staged_rows_per_row_given > 8 and ig.staged_rows_per_row < 4.8  # random numbering -> compact blocks
    assert torch.equal(torch.sort(ig.new_id).values, torch.arange(n, device="cuda"))
    record("kd_relabel_random_tet_120k", staged_given=ig.staged_rows_per_row_given, staged_kd=ig.staged_rows_per_row)
    monkeypatch.setattr(ops, "TC_AGG_MIN_NODES", 10 ** 12)
    with torch.no_grad():
        out_plain = model(x, ei)
    assert len(tc_calls) == 2  # the plain path stays on the gather kernel
    assert out_plan.shape == out_plain.shape == (n, 1)
    err = rel_max(out_plan.float().cpu(), out_plain.float().cpu())
    record("large_mesh_inference_path_vs_plain", rel_max=err)
    assert err < TOL_BF16
    # 'never': plan in the given numbering, no permutation
    monkeypatch.setattr(ops, "TC_AGG_MIN_NODES", 1000)
    model.node_reorder = "never"
    with torch.no_grad():
        out_never = model(x, ei)
    assert rel_max(out_never.float().cpu(), out_plain.float().cpu()) < TOL_BF16
    # 'auto': a mesh is prepared from its SECOND forward on (a screening loop sees every graph once)
    model.mesh_plan, model.node_reorder = "auto", "auto"
    ei_new = ei.clone()
    n0 = len(tc_calls)
    with torch.no_grad():
        first = model(x, ei_new)
        assert len(tc_calls) == n0
        second = model(x, ei_new)
        assert len(tc_calls) == n0 + 2
    assert rel_max(first.float().cpu(), out_plain.float().cpu()) < TOL_BF16 and rel_max(second.float().cpu(), out_plain.float().cpu()) < TOL_BF16


@pytest.mark.parametrize("order,reorder", [("native", "never"), ("random", "never"), ("random", "always")])
@pytest.mark.parametrize("h", [128, 256])
def test_aggregate_tc_long_pipelines_match_the_gather_kernel(ops, order, reorder, h):
    """Many blocks per CTA and many chunks per block (a randomly numbered lattice stages ~14 rows per output row): every
    barrier of the pipeline wraps its phase many times.  (A phase race in the own-row path passed every small-graph test and
    only showed beyond ~30 chunks per CTA.)  Reference: the gather kernel on the same CSR; the two differ by the association
    order of fp32 additions only (<= 1 bf16 ulp)."""
    import bench

    ei, pos, n = bench.cfg4_lattice_device((30, 60, 160), torch.device("cuda"), order)
    ops._INF_CACHE.clear()
    ig = ops.get_inference_graph(ei, n, pos=pos, reorder=reorder)
    g = ig.graph
    assert g.plan.usable and (ig.order is not None) == (reorder == "always")
    x = torch.randn(n, h, device="cuda", generator=torch.Generator(device="cuda").manual_seed(h)).bfloat16()
    xp = x if ig.order is None else x.index_select(0, ig.order)
    got = ops.aggregate_tc(g.plan, g.inv_deg, xp, g.num_edges)
    ref = ops.aggregate(g.rowptr, g.col, g.inv_deg, xp)
    err = (got.float() - ref.float()).abs().max().item()
    mism = int((got != ref).sum())
    record("aggregate_tc_vs_gather_288k", order=order, reorder=reorder, H=h, max_abs=err, bf16_mismatches=mism, elements=n * h,
           staged_rows_per_row=g.plan.staged_rows_per_row)
    assert err <= 2.0 ** -7 * max(ref.float().abs().max().item(), 1.0)
    assert mism < 1e-3 * n * h
    ops._INF_CACHE.clear()
    ops.clear_graph_cache()


@pytest.mark.parametrize("h", [64, 256])
def test_aggregate_tc_rectangular_sources_beyond_the_destination_rows(ops, h):
    """The partitioned layout: n_own destination rows whose sources index [own | halo] (more rows than destinations).  The block
    kernel with a plan of that rectangular CSR must match the gather kernel on the same CSR."""
    from deep_fem_uav_wing.gnn import synth

    mesh = synth.tet_lattice_wing(60000, seed=2)
    n_ext = mesh["num_nodes"]
    n_own = (n_ext * 2 // 3) // 128 * 128 + 37  # ragged last block
    ei = torch.from_numpy(mesh["edge_index"]).cuda()
    g = ops.get_graph(ei, n_ext)
    rowptr = g.rowptr[: n_own + 1].contiguous()
    e_loc = int(rowptr[-1].item())
    col = g.col[:e_loc].contiguous()
    assert int(col.max()) >= n_own  # there ARE halo sources behind the destination rows
    inv_deg = g.inv_deg[:n_own].contiguous()
    x = torch.randn(n_ext, h, device="cuda").bfloat16()
    plan = ops.build_agg_plan(rowptr, col, n_own)
    assert plan.check()
    got = ops.aggregate_tc(plan, inv_deg, x, e_loc, n_rows=n_own)
    want = ops.aggregate_rows(rowptr, col, inv_deg, x, n_own)
    assert got.shape == want.shape == (n_own, h)
    err = (got.float() - want.float()).abs().max().item() / want.float().abs().max().item()
    assert err < 1e-2, err
    mism = (got != want).float().mean().item()
    assert mism < 0.05, mism  # same fp32 sums up to association order: almost all elements round to the same bf16
    with pytest.raises(ValueError):
        ops.aggregate_tc(plan, inv_deg, x[: n_own - 1], e_loc, n_rows=n_own)
