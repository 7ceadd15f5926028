"""Host logic of the partitioned single-mesh execution (gnn/partition.py) on the CPU: the rank-local graphs plus
the halo exchange - emulated here by indexing the global tensor - must reproduce the global mean aggregation exactly."""
import numpy as np
import torch

import bench
from deep_fem_uav_wing.gnn import ops
from deep_fem_uav_wing.gnn.partition import build_partition, partition_bounds, send_lists_from_halos
from oracle import csr_aggregate_c, csr_oracle_c


def test_partition_with_emulated_halo_exchange_equals_global_aggregation():
    ei, pos, n = bench.cfg4_lattice_device((9, 20, 41), torch.device("cpu"), "random")
    new_id = ops.locality_order(pos, ei)
    ei2 = new_id[ei]
    world = 3
    bounds = partition_bounds(n, world)
    assert bounds[0] == 0 and bounds[-1] == n and all(int(b) % 128 == 0 for b in bounds[:-1])
    parts = [build_partition(ei2, bounds, r) for r in range(world)]
    send_lists_from_halos(parts)
    rng = np.random.default_rng(0)
    h = rng.standard_normal((n, 16)).astype(np.float32)  # global tensor in relabelled order
    rowptr, col, _, inv = csr_oracle_c(ei2.numpy(), n)
    want = csr_aggregate_c(rowptr, col, inv, h)
    halo_total = 0
    for p in parts:
        lo, hi = int(bounds[p.rank]), int(bounds[p.rank + 1])
        assert p.n_own == hi - lo and sum(p.recv_counts) == p.halo_ids.numel()
        assert torch.all((p.halo_ids < lo) | (p.halo_ids >= hi)) and torch.all(p.halo_ids[1:] > p.halo_ids[:-1])
        # the exchange: peer q sends h_q[send_index slice for p]; concatenated in peer order it must be h[halo_ids]
        recv = []
        for q in parts:
            off = sum(q.send_counts[: p.rank])
            idx = q.send_index[off: off + q.send_counts[p.rank]].numpy()
            recv.append(h[int(bounds[q.rank]) + idx])
        halo = np.concatenate(recv) if recv else np.zeros((0, 16), np.float32)
        assert np.array_equal(halo, h[p.halo_ids.numpy()])
        h_ext = np.concatenate([h[lo:hi], halo])
        n_ext = h_ext.shape[0]
        lr, lc, _, linv = csr_oracle_c(p.local_edge_index.numpy(), n_ext)
        got = csr_aggregate_c(lr, lc, linv, h_ext)[: p.n_own]
        # same rows and the same neighbours; a row's halo sources are numbered behind its own-range sources, so the fp32
        # additions happen in a different order than in the global CSR: equal up to rounding
        np.testing.assert_allclose(got, want[lo:hi], rtol=0, atol=2e-6)
        halo_total += p.halo_ids.numel()
    assert halo_total < 0.6 * n  # k-d order keeps the halos thin (a random 1-D split would need almost every row)
