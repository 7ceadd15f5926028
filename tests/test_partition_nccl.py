"""Partitioned single-mesh execution on GPUs (gnn/partition.py): one mesh split over 2 ranks with a halo
all-to-all per SAGE layer over NCCL must reproduce the unpartitioned forward (fp32 1e-5, bf16 2e-2 of max|ref|).
Needs >= 2 GPUs: ``gpurun --gpus 2 -- python -m pytest tests/test_partition_nccl.py -m gpu``."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_partitioned_forward_equals_unpartitioned(tmp_path):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    out = tmp_path / "partition.json"
    port = 29700 + (os.getpid() % 1500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tests", "partition_nccl_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    log_dir = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"partition_nccl_n{world}.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert len(res["cases"]) == 4 and res["cases"][3]["local_aggregation"] == "tensor-core blocks"
    for c in res["cases"]:
        tol = 1e-5 if c["dtype"] == "float32" else 2e-2
        assert c["rel_max_vs_unpartitioned"] < tol, c
        assert sum(c["own_rows_per_rank"]) == c["N"]
        assert max(c["halo_rows_per_rank"]) < 0.35 * c["N"], c  # k-d order: thin halos even for a randomly numbered mesh
