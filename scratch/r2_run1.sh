#!/bin/bash
# round 2, GPU call 1: new full-size parity tests, baseline bench, what-if timing builds, int8 peak probe
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_gop.py -m gpu -x -q -k "config2 or standalone" > gpurun_out/r2a_pytest_new.log 2>&1; tail -3 gpurun_out/r2a_pytest_new.log
python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err || tail -5 gpurun_out/r2a_bench.err
for w in 2 4 6 8 12 16; do
  PMCTF_LIB=$PWD/learned-pmctf_b200/lib/libpmctf_b200_w$w.so python bench.py --steps 2 --warmup 3 --frames 32 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_w$w.json 2> gpurun_out/r2a_bench_w$w.err
done
python bench.py --steps 2 --warmup 3 --frames 32 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_w0.json 2> gpurun_out/r2a_bench_w0.err
python - > gpurun_out/r2a_int8_peak.log 2>&1 <<'PY'
import torch, time
a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device="cuda")
for _ in range(3): torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("int8 burst TOPS", 2 * 8192**3 / (best * 1e-3) / 1e12)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; e0.record(); t0 = time.time()
while time.time() - t0 < 4:
    for _ in range(20): torch._int_mm(a, b)
    n += 20; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print("int8 sustained TOPS", n * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
PY
cat gpurun_out/r2a_int8_peak.log
python - <<'PY'
import json
for w in (0, 2, 4, 6, 8, 12, 16):
    try:
        d = json.load(open(f"gpurun_out/r2a_bench_w{w}.json")); print("whatif", w, "frames/s", round(d["value"], 1), "single-stream ms", round(d["roofline"]["single_stream_ms_per_step"], 1))
    except Exception as e: print("whatif", w, "failed", e)
d = json.load(open("gpurun_out/r2a_bench.json")); print("bench", d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
