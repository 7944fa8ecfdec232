#!/bin/bash
# Regenerates the tracked profiles/<tag>_* summaries from the raw outputs of tools/measure_round.sh in gpurun_out/ (tag = $1).
set -e
tag=${1:-r1z}
cd "$(dirname "$0")/.."
cuobjdump -sass learned-pmctf_b200/lib/libpmctf_b200.so 2>/dev/null | awk '/Function : .*lift_step_tc_kernelILi1/{f=1} /Function : /{if(!/lift_step_tc_kernelILi1/)f=0} f' > /tmp/sass_warp.txt
for op in UTCIMMA UTCBAR LDTM STTM UTCATOMSWS UBLKCP FFMA2 FMUL2 FADD2 SYNCS LDCU I2F.S64; do echo "#   $(grep -c "$op" /tmp/sass_warp.txt) $op"; done > /tmp/sasscounts.txt
{
echo "# ncu --set full --clock-control none --import-source on ($tag): lift_step_tc_kernel of this round -- continuation tiles (a CTA walks"
echo "# down a column strip; the rows two tiles share travel in registers: conv1 on 18 instead of 22 rows, conv2 / conv3 on 5 instead of 6"
echo "# 128-pixel blocks), conv1 -> conv2 -> conv3 pipelined through mbarriers (no CTA barrier between them: the MMA warp starts conv2 while"
echo "# conv1's later passes run and conv3 while conv2's last epilogues run), everything of round 1 (exact int8 digit-split MMAs, conv4 as"
echo "# per-tap partials, conv1/conv4 weights as kernel arguments, TMA-staged operand images)."
echo "# command: python tools/prof_tc.py (bench.py's weights, 4 planes per launch); launches 1,2: 1080p luma temporal forward MCTF"
echo "# (<1> WARP source, 4 x 2.21 Mpx = 17 280 tiles); launch 3: first row step of the 2-D lifting (<2> SKIP3, 4 x 1.1 Mpx = 8 640 tiles)."
echo "# Persistent grid 296 CTAs = 2 per SM, 288 threads, 113 KB smem.  Algorithmic HBM bytes of launch 1: 4 x 44.2 MB (20 B/px)."
echo "# Reading: no unit is saturated -- issue slots ~52-55 %, shared-memory data pipe (LSU + tensor-core operand fetch) ~67-78 %, tensor"
echo "# pipe ~21-25 %, DRAM 3 %.  The MMAs are hidden behind the CUDA-core work now (epilogue warps wait <10 % of their time for"
echo "# accumulators, profiles/${tag}_phases.txt); what bounds the kernel is the instruction stream of tanh / digit split / exact"
echo "# recombination at 18 warps per SM (two CTAs, 113 KB shared memory each)."
echo
python tools/summarize_ncu.py full gpurun_out/${tag}_tc_prof.ncu-rep
echo
echo "# SASS evidence (cuobjdump -sass, lift_step_tc_kernel<WARP>): tcgen05.mma -> UTCIMMA, tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM,"
echo "# tcgen05.alloc/dealloc -> UTCATOMSWS, cp.async.bulk (TMA) -> UBLKCP, fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2 (weights as uniform-register"
echo "# operands loaded by LDCU)"
cat /tmp/sasscounts.txt
} > profiles/${tag}_lift_step_tc_ncu.txt
{
echo "# ncu launch list summary ($tag): \`ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400\`"
echo "# command: python bench.py --steps 1 --warmup 3 --frames 16 --no-e2e --no-cpu-baseline --single-stream --no-torch-baseline --no-uvg --no-int8-peak  (1080p GOP-16s, tensor-core kernel;"
echo "# single stream so that the list is the serial launch order; the bench itself overlaps the luma and chroma chains on two streams)"
echo "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes"
echo
python tools/summarize_ncu.py launches gpurun_out/${tag}_launches.csv
} > profiles/${tag}_ncu_launches_summary.txt
cp gpurun_out/${tag}_launches.csv profiles/${tag}_ncu_launches_1gop.csv
cp gpurun_out/${tag}_bench.json profiles/${tag}_bench_1gpu_tensor.json
cp gpurun_out/${tag}_bench_ref.json profiles/${tag}_bench_reference_arm.json
[ -f gpurun_out/${tag}_hbm_kernels.json ] && cp gpurun_out/${tag}_hbm_kernels.json profiles/${tag}_hbm_kernels.json
[ -f gpurun_out/${tag}_phases.log ] && cp gpurun_out/${tag}_phases.log profiles/${tag}_phases.txt
python - "$tag" <<'PY'
import csv, io, json, subprocess, sys
tag = sys.argv[1]
out = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_tc_prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
r = data[0]
d = {"dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
     "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
     "what": "launch 1 of tools/prof_tc.py: temporal lifting step (WARP source) on 4 luma planes 1152x1920, ncu --set full",
     "algorithmic_bytes": 4 * 1152 * 1920 * 20, "source": f"gpurun_out/{tag}_tc_prof.ncu-rep via tools/refresh_profiles.sh"}
json.dump(d, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("traffic", d["dram_bytes_per_launch"], "algorithmic", d["algorithmic_bytes"])
PY
[ -f gpurun_out/${tag}_bench_8gpu.json ] && grep "^{" gpurun_out/${tag}_bench_8gpu.json > profiles/${tag}_bench_8gpu.json
echo refreshed profiles/${tag}_*
