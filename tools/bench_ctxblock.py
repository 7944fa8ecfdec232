"""context_fusion block of bench.py on its own:  python tools/bench_ctxblock.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import learned_pmctf_b200 as pkg  # noqa: E402

r = bench.run_context_fusion(pkg, torch.device("cuda:0"), bench.peaks())
print(json.dumps({k: r[k] for k in ("ms_per_plane", "ll_sequential")}, indent=1))
