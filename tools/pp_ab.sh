python -m pytest tests/test_postprocess.py tests/test_gpu_ctx.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import json, torch, bench
import learned_pmctf_b200 as pkg
r = bench.run_postprocess(pkg, torch.device("cuda:0"), bench.peaks())
print("PAIR  ", json.dumps({k: r[k] for k in ("ms_per_plane",)}), r["roofline"]["frac"], r["torch_gpu_baseline"]["tf32"])
PY
PMCTF_PP_SINGLE_CTA=1 python - <<'PY'
import json, torch, bench
import learned_pmctf_b200 as pkg
r = bench.run_postprocess(pkg, torch.device("cuda:0"), bench.peaks())
print("SINGLE", json.dumps({k: r[k] for k in ("ms_per_plane",)}), r["roofline"]["frac"])
PY
