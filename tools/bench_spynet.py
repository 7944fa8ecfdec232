"""SpyNet block of bench.py on its own (one JSON object):  python tools/bench_spynet.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import learned_pmctf_b200 as pkg  # noqa: E402

print(json.dumps(bench.run_spynet(pkg, torch.device("cuda:0"), bench.peaks()), indent=1))
