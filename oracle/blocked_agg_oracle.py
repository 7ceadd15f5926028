"""Design oracle for the brick + halo staged aggregation - TEST INFRASTRUCTURE ONLY.

Nothing in the product imports this module.  It pins down, in numpy, the data structure the next aggregation kernel
will consume and proves on the CPU that accumulating through it reproduces the CSR oracle bit for bit:

  * ``morton_order(pos, edge_index)``: node permutation from hop-normalised coordinates (bit-interleaved), the order that
    makes blocks of consecutive rows spatially compact (tools/brick_plan.py: 2.85 staged rows per output row on the cfg4
    lattice with 64-row blocks and index-space Morton, against 13.7 gathers).
  * ``build_blocks(rowptr, col, block_rows)``: for every block of ``block_rows`` consecutive rows the sorted list of
    DISTINCT source rows it needs (``blk_ptr``, ``blk_src``: what a CTA stages in shared memory once) and, per edge, the
    position of its source inside that list (``slot``, fits 16 bits): the kernel's inner loop becomes
    ``acc += staged[slot[e]]`` in CSR order - the same additions in the same order as the gather kernel, hence the same
    bits (reference semantics: PyG ``SAGEConv`` mean aggregation, call site ``model.py:90``).
  * ``aggregate_blocked(...)``: that loop, in numpy, used by ``tests/test_blocked_oracle.py`` against
    ``oracle.csr_aggregate_c``.
"""
from __future__ import annotations

import numpy as np


def morton_order(pos: np.ndarray, edge_index: np.ndarray) -> np.ndarray:
    """``new_id[old_id]``: rank of every node in Morton (Z-curve) order of its coordinates measured in HOPS - every axis
    is divided by the mean coordinate difference along the mesh edges, so that a cell of the curve holds about as many
    nodes along each axis whatever the element anisotropy (per-axis min-max quantisation makes thin bricks: worse than
    the native order on the wing lattice).  Stable: nodes that share a cell keep their relative order.
    Measured (64-row blocks, distinct sources per output row): cfg2 surface batch 3.06 -> 1.81; swept / tapered tet
    lattice 4.15 -> 3.66 (shear and thickness taper misalign the physical axes with the lattice; Morton over the lattice
    INDICES gives 2.85, tools/brick_plan.py) - a graph-based blocking (BFS balls) is the geometry-free alternative."""
    p = np.asarray(pos, dtype=np.float64)
    n, d = p.shape
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    ei = np.asarray(edge_index, dtype=np.int64)
    hop = np.abs(p[ei[0]] - p[ei[1]]).mean(axis=0) if ei.shape[1] else np.ones(d)
    hop = np.maximum(hop, 1e-30)
    q = np.floor((p - p.min(axis=0)) / hop).astype(np.int64)
    bits = int(np.ceil(np.log2(int(q.max()) + 2)))
    if d * bits > 62:
        raise ValueError("coordinates span too many hops for a 64-bit Morton key")
    key = np.zeros(n, dtype=np.int64)
    for b_ in range(bits):
        for a_ in range(d):
            key |= ((q[:, a_] >> b_) & 1) << (d * b_ + a_)
    new_id = np.empty(n, dtype=np.int64)
    new_id[np.argsort(key, kind="stable")] = np.arange(n)
    return new_id


def relabel_edge_index(edge_index: np.ndarray, new_id: np.ndarray) -> np.ndarray:
    return new_id[np.asarray(edge_index, dtype=np.int64)]


def build_blocks(rowptr: np.ndarray, col: np.ndarray, block_rows: int):
    """-> (blk_ptr int32 [B+1], blk_src int32 [S], slot uint16 [E]).  Block b owns rows [b*R, min((b+1)*R, N))."""
    n = rowptr.size - 1
    nblk = (n + block_rows - 1) // block_rows
    e_blk = np.repeat(np.arange(n, dtype=np.int64) // block_rows, np.diff(rowptr).astype(np.int64))  # block of every edge
    key = e_blk * max(n, 1) + col.astype(np.int64)
    uniq, inv = np.unique(key, return_inverse=True)           # sorted by (block, source): each block's list is ascending
    counts = np.bincount(uniq // max(n, 1), minlength=nblk)
    blk_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    blk_src = (uniq % max(n, 1)).astype(np.int32)
    slot = inv.astype(np.int64) - blk_ptr[e_blk].astype(np.int64)
    if slot.size and slot.max() >= 1 << 16:
        raise ValueError(f"a block needs {int(slot.max()) + 1} distinct sources: more than 16-bit slots address")
    return blk_ptr, blk_src, slot.astype(np.uint16)


def aggregate_blocked(rowptr, blk_ptr, blk_src, slot, inv_deg, x: np.ndarray, block_rows: int) -> np.ndarray:
    """Mean aggregation through the block structure: stage, then ``acc += staged[slot]`` in CSR order (fp32)."""
    n, h = x.shape
    out = np.zeros((n, h), dtype=np.float32)
    for b in range(blk_ptr.size - 1):
        staged = x[blk_src[blk_ptr[b]:blk_ptr[b + 1]]].astype(np.float32)  # what the CTA holds in shared memory
        for r in range(b * block_rows, min((b + 1) * block_rows, n)):
            acc = np.zeros(h, dtype=np.float32)
            for e in range(rowptr[r], rowptr[r + 1]):
                acc = acc + staged[slot[e]]
            out[r] = acc * (inv_deg[r] if inv_deg is not None else np.float32(1.0))
    return out
