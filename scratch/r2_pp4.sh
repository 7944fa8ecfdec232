#!/bin/bash
cd "$(dirname "$0")/.."
python scratch/prof_pp.py > gpurun_out/r2p_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pp_conv64 -s 16 -c 2 -f -o gpurun_out/r2p_pp_prof python scratch/prof_pp.py > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
