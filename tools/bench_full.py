"""Whole-codec GOP block of bench.py on its own:  python tools/bench_full.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import learned_pmctf_b200 as pkg  # noqa: E402

print(json.dumps(bench.run_full_codec(pkg, torch.device("cuda:0")), indent=1))
