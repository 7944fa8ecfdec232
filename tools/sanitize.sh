#!/bin/bash
# compute-sanitizer memcheck over smoke() and the small-shape parity tests (tag = $1): out-of-bounds / misaligned accesses of any kernel of the
# library, incl. the tcgen05 / TMA ones, show up as errors and a non-zero exit code.  Logs: gpurun_out/<tag>_memcheck_*.log
tag=${1:-san}
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
run() {  # name, timeout, command...
    name=$1; t=$2; shift 2
    timeout $t $CS "$@" > gpurun_out/${tag}_memcheck_${name}.log 2>&1
    echo "$name: rc=$? $(grep -c 'Invalid\|misaligned\|Error:' gpurun_out/${tag}_memcheck_${name}.log) error lines; $(grep 'ERROR SUMMARY' gpurun_out/${tag}_memcheck_${name}.log | tail -1); $(tail -1 gpurun_out/${tag}_memcheck_${name}.log | cut -c1-120)"
}
run smoke 400 python -c "import __graft_entry__ as g; g.smoke()"
run parity 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "flow_warp_vs_oracle or chroma_mv or predict_update_vs_oracle or iwave1d or lift2d or quantize_dequantize or random_shapes or errors"
run gop 400 python -m pytest tests/test_gpu_gop.py -x -q -m gpu -k "unpack_u8 or quantize_stats or per_plane_steps or code_gop_vs_oracle"
run llar 400 python -m pytest tests/test_gpu_llar.py -x -q -m gpu -k "oracle or parallel_encoder"
run ctx 400 python -m pytest tests/test_gpu_ctx.py tests/test_gpu_spynet.py tests/test_postprocess.py -x -q -m gpu
