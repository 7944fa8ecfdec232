#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python - > gpurun_out/r2p_pp.json 2> gpurun_out/r2p_pp.err <<'PY'
import sys, json; sys.path.insert(0, ".")
import torch, bench, learned_pmctf_b200 as P
print(json.dumps(bench.run_postprocess(P, torch.device("cuda"), bench.peaks()), indent=1))
PY
cat gpurun_out/r2p_pp.json; tail -3 gpurun_out/r2p_pp.err
