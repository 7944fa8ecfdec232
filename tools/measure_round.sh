#!/bin/bash
# final measurement set of a round (tag = $1): parity tests, full bench, ncu launch list, ncu full capture of the top kernel
tag=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
LL="--steps 1 --warmup 3 --frames 16 --no-e2e --no-cpu-baseline --single-stream --no-torch-baseline --no-uvg --no-int8-peak"
python bench.py $LL > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $LL > gpurun_out/${tag}_ncu1.log 2>&1
[ -f learned-pmctf_b200/lib/libpmctf_b200_timing.so ] && python tools/tc_phases.py > gpurun_out/${tag}_phases.log 2>&1
python tools/prof_tc.py > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lift_step_tc -c 3 -f -o gpurun_out/${tag}_tc_prof python tools/prof_tc.py > gpurun_out/${tag}_ncu2.log 2>&1
# secondary rows (section 8f): four-step network + SpyNet evidence; every ncu run is bounded (-c)
bash tools/prof_ctx.sh ${tag}
python tools/bench_spynet.py > gpurun_out/${tag}_spynet_bench.json 2>> gpurun_out/${tag}_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:pair_conv_kernel -s 25 -c 5 -f -o gpurun_out/${tag}_pair_conv python tools/bench_spynet.py > gpurun_out/${tag}_ncu3.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print('frames/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
