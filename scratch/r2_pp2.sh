#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_postprocess.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
bash scratch/r2_pp3.sh 2>&1 | grep -E "ms_per_plane|achieved|frac\"|max_abs"
