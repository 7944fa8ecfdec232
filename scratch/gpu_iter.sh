#!/bin/bash
# one optimisation iteration on the GPU box: parity tests, phase timing, short bench (tag = $1)
tag=${1:-iter}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
# phase timing needs the timing build: see scratch/tc_phases.py
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print('frames/s', d['value'], 'frac', d['roofline']['frac'])"
