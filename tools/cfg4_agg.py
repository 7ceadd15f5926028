#!/usr/bin/env python3
"""BASELINE.json config 4 aggregation roofline only (the `aggregation_cfg4` block of bench.py), for quick GPU iterations:
    python tools/cfg4_agg.py [iters] > gpurun_out/cfg4_agg.json"""
import json
import os
import sys

import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
import bench  # noqa: E402

if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    print(json.dumps(bench.aggregation_cfg4(torch.device("cuda", 0), bench.peaks(), iters=iters)))
