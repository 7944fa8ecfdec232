#!/bin/bash
# profiles/<tag>_ctx_* from the raw outputs of tools/prof_ctx.sh in gpurun_out/ (tag = $1)
set -e
tag=${1:-r2z}
cd "$(dirname "$0")/.."
cuobjdump -sass learned-pmctf_b200/lib/libpmctf_b200.so 2>/dev/null | awk '/Function : .*ctx_conv112_kernel/{f=1} /Function : /{if(!/ctx_conv112_kernel/)f=0} f' > /tmp/sass_ctx.txt
{
echo "# ncu --set full --clock-control none --import-source on ($tag): ctx_conv112_kernel, the 112 -> 112 3x3 layer of the four-step"
echo "# entropy-parameter network (csrc/pmctf_ctx.cu) as a tcgen05 CTA-pair implicit GEMM: cta_group::2, M = 256 (two 4x30-pixel tiles), N = 112,"
echo "# K = 16 bf16, 63 MMAs per tile pair, fp32 accumulators in TMEM (4 buffers), inputs by TMA tensor loads (one 43 KB box per tile and CTA, zero"
echo "# fill = padding), all 9 x 7 weight slabs resident in shared memory (113 KB = half of the layer per CTA), 8 epilogue warps."
echo "# command: python tools/bench_ctx.py --profile; one level-0 subband of a 1080p luma plane (1 x 576 x 960, 0.55 Mpx, 125 GFLOP per launch)."
echo "# launch 1: bf16 in -> LeakyReLU -> bf16 out (ContextResidual.conv1; 448 B/px algorithmic HBM traffic);"
echo "# launch 2: bf16 in + fp32 skip -> fp32 + bf16 out (ContextResidual.conv2; 1344 B/px = 743 MB per launch)."
echo "# Reading: launch 1 is tensor-bound (tensor pipe busy 60 % of ncu's cold, serialised duration; 1.11-1.19 PFLOP/s = 80-85 % of the measured"
echo "# bf16 peak when timed warm with CUDA events); launch 2 moves 708 MB through DRAM (= its algorithmic bytes, no re-reads) and is bounded by"
echo "# HBM latency / bandwidth in the epilogue (long_scoreboard), not by the MMAs."
echo
python tools/summarize_ncu.py full gpurun_out/${tag}_ctx_conv112.ncu-rep
echo
echo "# SASS evidence (cuobjdump -sass, ctx_conv112_kernel): tcgen05.mma.cta_group::2 -> UTCHMMA.2CTA, cp.async.bulk.tensor.4d.cta_group::2 -> UTMALDG.4D.2CTA,"
echo "# tcgen05.commit...multicast -> UTCBAR.2CTA.MULTICAST, tcgen05.alloc.cta_group::2 -> UTCATOMSWS.2CTA, barrier.cluster -> UCGABAR_ARV / UCGABAR_WAIT"
for op in UTCHMMA.2CTA UTMALDG.4D.2CTA UTCBAR.2CTA.MULTICAST UTCATOMSWS.2CTA UCGABAR LDTM UBLKCP MAPA; do echo "#   $(grep -c "$op" /tmp/sass_ctx.txt) $op"; done
} > profiles/${tag}_ctx_conv112_ncu.txt
{
echo "# ncu launch list ($tag): \`ncu --metrics gpu__time_duration.sum --clock-control none\` over python tools/bench_ctx.py --profile"
echo "# (two single layers, then two passes of ContextFusionFourStep over the 12 high-pass subbands of one 1080p luma plane, the first of which"
echo "# packs the weights: ctx_pack_kernel is not part of the steady state).  Cold-cache, serialised times: compare SHARES."
echo
python tools/summarize_ncu.py launches gpurun_out/${tag}_ctx_launches.csv
} > profiles/${tag}_ctx_launches_summary.txt
cp gpurun_out/${tag}_ctx_bench.json profiles/${tag}_ctx_bench.json
if [ -f gpurun_out/${tag}_pair_conv.ncu-rep ]; then
{
echo "# ncu --set full --clock-control none --import-source on ($tag): pair_conv_kernel (csrc/pmctf_pairconv.cu), the run-time-shaped CTA-pair"
echo "# convolution, on SpyNet's 7x7 layers at the finest pyramid level of a 1080p pair (1 x 1152 x 1920; python tools/bench_spynet.py, launches"
echo "# 26-30 of the kernel = the five layers 8->32->64->32->16->2 of the finest level).  N <= 64 output channels per layer: the MMAs (M = 256, N = 16..64, K = 16) are"
echo "# bounded by the A-operand fetch from shared memory, not by the tensor pipe; the 8 -> 32 layer pads K from 8 to 16."
echo
python tools/summarize_ncu.py full gpurun_out/${tag}_pair_conv.ncu-rep
} > profiles/${tag}_pair_conv_ncu.txt
cp gpurun_out/${tag}_spynet_bench.json profiles/${tag}_spynet_bench.json
fi
echo refreshed profiles/${tag}_ctx_*
