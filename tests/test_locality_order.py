"""``ops.locality_order`` (k-d relabelling for the blocked aggregation) on the CPU: it is plain torch, so its host logic is
checked here - a valid permutation, compact blocks on a randomly numbered lattice, identity-sized inputs."""
import numpy as np
import torch

import bench
from deep_fem_uav_wing.gnn import ops


def _staged_rows_per_row(ei, n, block_rows=128):
    blk = ei[1] // block_rows
    return np.unique(blk * n + ei[0]).size / n


def test_locality_order_is_a_permutation_and_compacts_blocks():
    dims = (12, 30, 64)
    for order in ("native", "random"):
        ei, pos, n = bench.cfg4_lattice_device(dims, torch.device("cpu"), order)
        new_id = ops.locality_order(pos, ei)
        assert new_id.dtype == torch.int64 and torch.equal(torch.sort(new_id).values, torch.arange(n))
        before = _staged_rows_per_row(ei.numpy(), n)
        after = _staged_rows_per_row(new_id[ei].numpy(), n)
        assert after < 3.6, (order, before, after)
        if order == "random":
            assert before > 10 and after < 0.35 * before
    # tiny graphs: nothing to do
    assert torch.equal(ops.locality_order(torch.rand(5, 3), torch.zeros(2, 0, dtype=torch.long)), torch.arange(5))
    # no edges: falls back to raw coordinates, still a permutation
    p = torch.rand(1000, 3)
    nid = ops.locality_order(p, torch.zeros(2, 0, dtype=torch.long))
    assert torch.equal(torch.sort(nid).values, torch.arange(1000))


def test_static_mesh_is_recognised_from_its_second_use_and_by_identity_only():
    """Large-mesh preparation policy (``model.mesh_plan = 'auto'``): a mesh counts as static from the SECOND forward on the same
    edge_index tensor OBJECT; a new tensor - even one the allocator places at the address of a freed one - starts over."""
    ops._SEEN_MESHES.clear()
    a = torch.zeros(2, 10, dtype=torch.long)
    assert not ops.static_mesh_seen_before(a)
    assert ops.static_mesh_seen_before(a) and ops.static_mesh_seen_before(a)
    b = a.clone()
    assert not ops.static_mesh_seen_before(b) and ops.static_mesh_seen_before(b)
    ident = id(a)
    del a
    c = torch.zeros(2, 10, dtype=torch.long)  # may or may not reuse the Python object slot of `a`
    if id(c) == ident:
        assert not ops.static_mesh_seen_before(c)  # the weak reference of the dead tensor does not vouch for the new one
    for _ in range(40):  # bounded
        ops.static_mesh_seen_before(torch.zeros(2, 1, dtype=torch.long))
    assert len(ops._SEEN_MESHES) <= 16
