#!/bin/bash
# ncu evidence for the four-step entropy-parameter network (run under gpurun; outputs in gpurun_out/<tag>_ctx_*):
#   launch list of one module pass over the 12 subbands of a 1080p luma plane, and one --set full capture of the CTA-pair layer
tag=${1:-r2z}
python tools/bench_ctx.py > gpurun_out/${tag}_ctx_bench.json 2> gpurun_out/${tag}_ctx_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${tag}_ctx_launches.csv \
    python tools/bench_ctx.py --profile > gpurun_out/${tag}_ctx_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctx_conv112 -c 2 -f -o gpurun_out/${tag}_ctx_conv112 \
    python tools/bench_ctx.py --profile > gpurun_out/${tag}_ctx_ncu2.log 2>&1
ls -la gpurun_out/${tag}_ctx_*
