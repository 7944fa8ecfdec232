#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_postprocess.py -m gpu -x -q -s > gpurun_out/r2p_pytest.log 2>&1; tail -25 gpurun_out/r2p_pytest.log
