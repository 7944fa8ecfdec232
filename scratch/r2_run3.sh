#!/bin/bash
cd "$(dirname "$0")/.."
python scratch/tc_phases.py > gpurun_out/r2c_phases.log 2>&1; cat gpurun_out/r2c_phases.log
python bench.py --steps 2 --warmup 3 --frames 32 --no-e2e --no-cpu-baseline --no-uvg --no-torch-baseline --no-int8-peak > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err || tail -5 gpurun_out/r2c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench.json')); print('frames/s', d['value'], 'single', d['roofline']['single_stream_ms_per_step'])"
