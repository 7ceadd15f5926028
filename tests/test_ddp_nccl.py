"""N-rank NCCL parity on GPUs: the product model under ``MeshDataParallel`` + ``GraphedTrainStep(ddp=...)``
- the graph-captured, bucketed NCCL all-reduce that bench.py and SCALE runs use - must follow the single-process run on the
union batch (``model.py:151``, ``train_gnn.py:150``): first-step gradients and the 4-step AdamW parameter trajectory within
rel-L2 1e-5.  Needs >= 2 GPUs (``gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl.py -m gpu``); the run log is
kept under profiles/."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
TOL = 1e-5


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2])
def test_nccl_ranks_follow_the_union_batch_run(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    out = tmp_path / "ddp_nccl.json"
    port = 29600 + (os.getpid() % 1500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tests", "ddp_nccl_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    log_dir = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"ddp_nccl_parity_n{world}.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert res["backend"] == "nccl" and res["world"] == world
    assert res["graph_captured_with_nccl"], "the step was never replayed from a CUDA graph"
    assert res["ranks_hold_identical_parameters"]
    assert res["worst_first_step_grad_rel_l2"] < TOL, res
    assert max(res["worst_param_rel_l2_per_step"]) < TOL, res
