#!/bin/bash
# Regenerates the tracked profiles/<tag>_* summaries from the raw outputs of scratch/gpu_final.sh in gpurun_out/ (tag = $1).
set -e
tag=${1:-r1z}
cd "$(dirname "$0")/.."
cuobjdump -sass learned-pmctf_b200/lib/libpmctf_b200.so 2>/dev/null | awk '/Function : .*lift_step_tc_kernelILi1/{f=1} /Function : /{if(!/lift_step_tc_kernelILi1/)f=0} f' > /tmp/sass_warp.txt
for op in UTCIMMA UTCBAR LDTM STTM UTCATOMSWS UBLKCP FFMA2 FMUL2 FADD2 SYNCS LDCU I2F.S64; do echo "#   $(grep -c "$op" /tmp/sass_warp.txt) $op"; done > /tmp/sasscounts.txt
{
echo "# ncu --set full --clock-control none --import-source on, round 1 final ($tag): lift_step_tc_kernel after this round's work on the shared-memory"
echo "# pipe and the instruction count: conv4 as per-tap partials in the conv3 epilogue, conv1 residual stashed (not recomputed), one in-place"
echo "# set of digit planes, conv1/conv4 weights + biases as kernel arguments (constant-bank FFMA2 operands), second-order tanh, 14 MMAs per"
echo "# 128-pixel block, operand images and tanh table staged by TMA bulk copies, register-resident skip-filter source, word-arithmetic recombination."
echo "# command: python scratch/prof_tc.py (scratch/gpu_final.sh); launches 1,2: 1080p luma temporal forward MCTF (<1> WARP source, 2.21 Mpx);"
echo "# launch 3: first row step of the 2-D lifting (<2> SKIP3, 1.1 Mpx).  Persistent grid 296 CTAs = 2 per SM, 288 threads, 112.9 KB smem."
echo "# Algorithmic HBM bytes of launch 1/2: 44 MB; measured dram r+w below that (outputs stay in L2)."
echo "# Reading: the shared-memory data pipe (LSU wavefronts + tensor-core operand fetch) is the busiest unit -- l1tex__data_pipe_{lsu,tc}_wavefronts"
echo "# sum to ~80 % of peak over ncu's (cold, serialised) duration and more over the un-profiled one; issue slots ~57 %; tensor pipe ~25 %."
echo
python tools/summarize_ncu.py full gpurun_out/${tag}_tc_prof.ncu-rep
echo
echo "# SASS evidence (cuobjdump -sass, lift_step_tc_kernel<WARP>): tcgen05.mma -> UTCIMMA, tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM,"
echo "# tcgen05.alloc/dealloc -> UTCATOMSWS, cp.async.bulk (TMA) -> UBLKCP, fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2 (weights as uniform-register"
echo "# operands loaded by LDCU)"
cat /tmp/sasscounts.txt
} > profiles/${tag}_lift_step_tc_ncu.txt
{
echo "# ncu launch list summary, round 1 final ($tag): \`ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900\`"
echo "# command: python bench.py --steps 1 --warmup 3 --frames 16 --no-e2e --no-cpu-baseline --single-stream  (1080p GOP-16s, tensor-core kernel;"
echo "# single stream so that the list is the serial launch order; the bench itself overlaps the luma and chroma chains on two streams)"
echo "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes"
echo
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv
} > profiles/${tag}_ncu_launches_summary.txt
cp gpurun_out/${tag}_launches.csv profiles/${tag}_ncu_launches_1gop.csv
cp gpurun_out/${tag}_bench.json profiles/${tag}_bench_1gpu_tensor.json
cp gpurun_out/${tag}_bench_ref.json profiles/${tag}_bench_reference_arm.json
[ -f gpurun_out/${tag}_hbm_kernels.json ] && cp gpurun_out/${tag}_hbm_kernels.json profiles/${tag}_hbm_kernels.json
[ -f gpurun_out/${tag}_bench_8gpu.json ] && grep "^{" gpurun_out/${tag}_bench_8gpu.json > profiles/${tag}_bench_8gpu.json
echo refreshed profiles/${tag}_*
