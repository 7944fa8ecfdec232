#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gop.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 120 python scratch/tc_phases.py > gpurun_out/r2f_phases.log 2>&1; head -4 gpurun_out/r2f_phases.log
timeout 300 python bench.py --steps 2 --warmup 3 --frames 32 --no-e2e --no-cpu-baseline --no-uvg --no-torch-baseline --no-int8-peak > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err || tail -5 gpurun_out/r2f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print('frames/s', d['value'], 'single', d['roofline']['single_stream_ms_per_step'])"
