"""bench.py's host-side helpers (CPU): the device-side generator of the config-4 lattice must produce the graph family of
``synth.tet_lattice_wing`` (BASELINE.json config 4), and the reference arm must not load the product library."""
import os
import subprocess
import sys

import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)


def _edge_set(ei, n):
    ei = np.asarray(ei)
    return np.unique(ei[0].astype(np.int64) * n + ei[1])


def test_cfg4_lattice_generator_matches_synth():
    import bench
    from deep_fem_uav_wing.gnn import synth

    dims = (5, 7, 9)
    m = synth.tet_lattice_wing(0, seed=1, dims=dims, params={"span_m": 1.5, "chord_m": 0.35, "sweep_deg": 15.0, "thickness_ratio": 0.10})
    ei, pos_n, n = bench.cfg4_lattice_device(dims, torch.device("cpu"), "native")
    assert n == m["num_nodes"] == 5 * 7 * 9
    assert np.array_equal(_edge_set(ei.numpy(), n), _edge_set(m["edge_index"], n))
    np.testing.assert_allclose(pos_n.numpy(), m["x"][:, :3], atol=2e-6)
    # random order: the same graph up to a node relabelling (degree multiset and edge count are invariants)
    ei_r, pos_r, _ = bench.cfg4_lattice_device(dims, torch.device("cpu"), "random")
    assert ei_r.shape == ei.shape
    assert np.array_equal(np.sort(np.bincount(ei_r[1].numpy(), minlength=n)), np.sort(np.bincount(ei[1].numpy(), minlength=n)))
    assert np.allclose(np.sort(pos_r.numpy(), axis=0), np.sort(pos_n.numpy(), axis=0))
    assert bench.CFG4_DIMS[0] * bench.CFG4_DIMS[1] * bench.CFG4_DIMS[2] == 2_001_384


def test_reference_arm_does_not_load_the_product_library():
    """`bench.py --impl reference` is the CPU arm: its process must not map libdfw_b200.so."""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import bench; bench.NODES=600; bench.main();"
            "maps=open('/proc/self/maps').read(); assert 'libdfw_b200' not in maps, 'product library mapped in the CPU arm';"
            "assert 'deep_fem_uav_wing' not in sys.modules")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=REPO, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json

    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 1 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)  # all host threads although OMP_NUM_THREADS=1 was exported
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
