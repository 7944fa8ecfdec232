#!/bin/bash
# round 2, GPU call 4: batched coder (all H frames in one batch), host symbols, full bench with uvg + baselines
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -6 gpurun_out/r2d_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err || tail -20 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench.json")); r = d["roofline"]
print("bench", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], "frac", r["frac"], "single", r["single_stream_ms_per_step"], "launches", d["gpu_launches"])
print("uvg", d["uvg"]); print("torch", d["torch_gpu_baseline"]); print("int8", r.get("int8_dense_peak_tops_measured"), r.get("executed_int8_frac_of_measured_sustained"))
PY
