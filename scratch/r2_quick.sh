#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_gop.py -m gpu -x -q -k "config2 or standalone or config1 or concurrent" > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 2 --warmup 3 --frames 32 --no-e2e --no-cpu-baseline --no-uvg --no-torch-baseline --no-int8-peak > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err || tail -5 gpurun_out/r2q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print('frames/s', d['value'], 'single', d['roofline']['single_stream_ms_per_step'], 'frac', d['roofline']['frac'])"
