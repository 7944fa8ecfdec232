"""Bitstream path of pWave (pWave.py:381-529) on one 1080p luma plane: compress -> file -> decompress, wall clock (host entropy coder
included), with a coarse breakdown from torch's profiler-free timers.
    python tools/bench_bitstream.py  ->  one JSON line"""
import json, os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
import numpy as np
import torch
import learned_pmctf_b200 as pkg
from test_pwave_coder import _randomise

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = _randomise(pkg.pWave(entropy_model=True)).to(dev).eval()
m.update()
H, W = 1152, 1920
g = torch.Generator(device=dev).manual_seed(5)
x = torch.nn.functional.avg_pool2d(torch.rand((1, 1, H + 4, W + 4), device=dev, generator=g) * 255, 5, 1).round()
out = {"plane": [H, W]}
with torch.no_grad():
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x_hat = m.compress(x, sideinfo=(1, 1, H, W), file_name="/tmp/plane.bin", q_index=12)
        torch.cuda.synchronize()
        out["compress_ms"] = (time.perf_counter() - t0) * 1e3
    out["bytes"] = os.path.getsize("/tmp/plane.bin")
    for rep in range(1 if os.environ.get("QUICK") else 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        back = m.decompress("/tmp/plane.bin", padding=64, q_index=12)["x_hat"]
        torch.cuda.synchronize()
        out["decompress_ms"] = (time.perf_counter() - t0) * 1e3
    out["round_trip_exact"] = bool(torch.equal(back, x_hat))
out["coder"] = {"ec_thread": os.environ.get("PMCTF_EC_THREAD", "0") == "1", "stream_part": int(os.environ.get("PMCTF_STREAM_PART", "1"))}
print(json.dumps(out))
