#!/usr/bin/env python3
"""Sizing of the brick + halo staging design for the aggregation.  CPU only.

Rows are relabelled in Morton order and cut into blocks of R consecutive rows; a block stages the UNION of its rows'
neighbours once in shared memory and accumulates from there.  For the cfg4 lattice and a cfg2 surface mesh this prints,
per block size: staged rows per output row (the L2->SM traffic relative to one row read per output row; the gather
kernel moves `degree` rows per output row), and the shared memory a block needs for 512-byte rows."""
import json
import os
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200"), os.path.join(REPO, "tools")]
from reorder_prof import lattice_edges, orders  # noqa: E402


def plan(ei, n, block_rows, row_bytes=512):
    dst, src = ei[1].astype(np.int64), ei[0].astype(np.int64)
    blk = dst // block_rows
    pair = np.unique(blk * n + src)              # (block, neighbour) pairs, each neighbour counted once per block
    per_block = np.bincount(pair // n, minlength=(n + block_rows - 1) // block_rows)
    return {"block_rows": block_rows, "staged_rows_per_output_row": round(float(pair.size) / n, 2),
            "gathers_per_output_row": round(float(dst.size) / n, 2), "max_staged_rows": int(per_block.max()),
            "smem_KB_max": round(per_block.max() * row_bytes / 1024, 1), "smem_KB_mean": round(per_block.mean() * row_bytes / 1024, 1)}


def relabel(ei, key):
    n = key.size
    new_id = np.empty(n, dtype=np.int64)
    new_id[np.argsort(key, kind="stable")] = np.arange(n)
    return new_id[ei]


def main():
    nx, ny, nz = 38, 114, 462  # cfg4 lattice (2 001 384 nodes)
    ei = lattice_edges(nx, ny, nz)
    ords = orders(nx, ny, nz)
    n = nx * ny * nz
    for name in ("native", "morton"):
        e = ei if ords[name] is None else relabel(ei, ords[name])
        for r in (64, 128, 256):
            print(json.dumps({"mesh": "cfg4 tet lattice", "order": name, **plan(e, n, r)}), flush=True)
    from deep_fem_uav_wing.gnn import synth

    m = synth.surface_tri_wing(200000, seed=42)
    e, n = m["edge_index"], m["num_nodes"]
    for r in (64, 128, 256):
        print(json.dumps({"mesh": "cfg2 surface batch (native order)", **plan(e, n, r)}), flush=True)


if __name__ == "__main__":
    main()
