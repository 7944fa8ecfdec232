#!/bin/bash
cd "$(dirname "$0")/.."
python scratch/prof_tc.py > gpurun_out/r2e_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lift_step_tc -s 10 -c 3 -f -o gpurun_out/r2e_tc_prof python scratch/prof_tc.py > gpurun_out/r2e_ncu2.log 2>&1
tail -3 gpurun_out/r2e_ncu2.log
