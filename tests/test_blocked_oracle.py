"""Design oracle of the brick + halo staged aggregation (oracle/blocked_agg_oracle.py): the block structure reproduces
the CSR oracle bit for bit, and Morton order is what makes the blocks small."""
import numpy as np

from oracle import csr_aggregate_c, csr_oracle_c
from oracle.blocked_agg_oracle import aggregate_blocked, build_blocks, morton_order, relabel_edge_index

from deep_fem_uav_wing.gnn import synth


def test_blocked_aggregation_equals_the_csr_oracle_bit_for_bit():
    rng = np.random.default_rng(0)
    m = synth.tet_lattice_wing(1500, seed=3)
    n = m["num_nodes"]
    cases = [(m["edge_index"], n), (rng.integers(0, 300, size=(2, 4000)).astype(np.int64), 300),  # duplicates, self loops
             (np.zeros((2, 0), dtype=np.int64), 5)]
    for ei, nn in cases:
        rowptr, col, _, inv = csr_oracle_c(ei, nn)
        x = rng.standard_normal((nn, 16)).astype(np.float32)
        ref = csr_aggregate_c(rowptr, col, inv, x)
        for block_rows in (1, 7, 64):
            blk_ptr, blk_src, slot = build_blocks(rowptr, col, block_rows)
            assert blk_ptr[-1] == blk_src.size and slot.size == col.size
            for b in range(blk_ptr.size - 1):  # every list ascending and duplicate-free
                s = blk_src[blk_ptr[b]:blk_ptr[b + 1]]
                assert np.all(np.diff(s) > 0)
            got = aggregate_blocked(rowptr, blk_ptr, blk_src, slot, inv, x, block_rows)
            assert np.array_equal(got, ref), (nn, block_rows)


def test_morton_order_is_a_permutation_and_shrinks_the_blocks():
    m = synth.surface_tri_wing(20000, seed=5)
    n, ei = m["num_nodes"], m["edge_index"]
    new_id = morton_order(m["pos"], ei)
    assert np.array_equal(np.sort(new_id), np.arange(n))
    staged = {}
    for name, e in (("native", ei), ("morton", relabel_edge_index(ei, new_id))):
        rowptr, col, _, _ = csr_oracle_c(e, n)
        blk_ptr, _, _ = build_blocks(rowptr, col, 64)
        staged[name] = blk_ptr[-1] / n
    assert staged["morton"] < 0.75 * staged["native"] < 0.75 * 6.0  # distinct sources per output row; 6 gathers per row either way
    # equivariance: aggregating the relabelled graph on relabelled features is the relabelled result
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 8)).astype(np.float32)
    rowptr, col, _, inv = csr_oracle_c(ei, n)
    ref = csr_aggregate_c(rowptr, col, inv, x)
    e2 = relabel_edge_index(ei, new_id)
    rp2, col2, _, inv2 = csr_oracle_c(e2, n)
    x2 = np.empty_like(x)
    x2[new_id] = x
    out2 = csr_aggregate_c(rp2, col2, inv2, x2)
    # the set of neighbours is the same, their CSR (summation) order changes with the labels: equal up to fp32 rounding
    assert np.allclose(out2[new_id], ref, rtol=1e-5, atol=1e-6)
    assert morton_order(np.zeros((0, 3)), np.zeros((2, 0), dtype=np.int64)).size == 0
