#!/bin/bash
# A/B runs of library variants (PMCTF_LIB) and host switches on the headline bench: one line per variant in gpurun_out/ab_<tag>.txt
tag=${1:-ab}; shift
out=gpurun_out/ab_${tag}.txt; : > $out
LL="--steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-torch-baseline --no-uvg --no-int8-peak --no-postprocess --no-full-codec --no-spynet --no-train-block --no-context-fusion"
run() {  # name, env assignments...
    name=$1; shift
    v=$(env "$@" timeout 300 python bench.py $LL 2>gpurun_out/ab_${tag}_${name}.err | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('%.2f frames/s  frac %.4f  single-stream kernel %.1f TFLOP/s' % (d['value'], d['roofline']['frac'], d['roofline']['achieved']))")
    echo "$name: $v" | tee -a $out
}
L=learned-pmctf_b200/lib
run base X=1
run base_again X=1
run lp_serial PMCTF_CONCURRENT_LP=0
for v in "$@"; do
    [ -f $L/libpmctf_b200_$v.so ] && run $v PMCTF_LIB=$PWD/$L/libpmctf_b200_$v.so
done
