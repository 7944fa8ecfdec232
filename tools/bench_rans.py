"""Host throughput of the native rANS coder (symbols per second) by sub-stream count:  python tools/bench_rans.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_pmctf_b200 as P  # noqa: E402,F401
from learned_pmctf_b200.entropy_models.entropy_models import GaussianEncoder  # noqa: E402
from learned_pmctf_b200.models import MLCodec_rans  # noqa: E402

g = GaussianEncoder("laplace")
g.update()
cdf, ln, off = g.get_cdf_info()
r = np.random.default_rng(0)
n = 8_000_000
idx = r.integers(0, 256, n).astype(np.int16)
sym = np.round(r.laplace(0, np.exp(np.linspace(np.log(0.01), np.log(64.0), 256))[idx])).clip(-30000, 30000).astype(np.int16)
print("cores", os.cpu_count())
for parts in (1, 2, 4, 8, 16):
    e = MLCodec_rans.RansEncoder(parts > 1, parts)
    for _ in range(2):      # second pass: buffers warm
        e.reset()
        t0 = time.perf_counter()
        e.encode_with_indexes(sym, idx, cdf, ln, off)
        e.flush()
        st = e.get_encoded_stream()
        t1 = time.perf_counter()
    d = MLCodec_rans.RansDecoder(parts)
    d.set_stream(st)
    t2 = time.perf_counter()
    out = d.decode_stream(idx, cdf, ln, off)
    t3 = time.perf_counter()
    print(f"parts {parts:2d}: encode {n / (t1 - t0) / 1e6:6.1f} Msym/s, decode {n / (t3 - t2) / 1e6:6.1f} Msym/s, exact {np.array_equal(out, sym)}")
