#!/bin/bash
# round 2, GPU call 2: continuation-tile kernel + boundary hardening: full GPU suite, bench
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -15 gpurun_out/r2b_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err || tail -5 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2b_bench.json")); print("bench", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["single_stream_ms_per_step"])
PY
