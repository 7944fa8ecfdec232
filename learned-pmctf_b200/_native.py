"""ctypes binding of the C-ABI library (include/pmctf_b200.h) -- the only way the Python host
reaches the GPU kernels.  There is no fallback: if the library is missing or a call is rejected
a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("PMCTF_LIB") or os.path.join(_HERE, "lib", "libpmctf_b200.so")  # PMCTF_LIB: profiling builds only
SOURCES = [os.path.join(_HERE, "csrc", "pmctf_kernels.cu"), os.path.join(_HERE, "csrc", "pmctf_umma_test.cu"),
           os.path.join(_HERE, "csrc", "pmctf_lift_tc.cu"), os.path.join(_HERE, "csrc", "pmctf_train.cu"),
           os.path.join(_HERE, "csrc", "pmctf_pp.cu"), os.path.join(_HERE, "csrc", "pmctf_rans.cu"), os.path.join(_HERE, "csrc", "pmctf_ctx.cu"), os.path.join(_HERE, "csrc", "pmctf_pairconv.cu"), os.path.join(_HERE, "csrc", "pmctf_llar.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "pmctf_umma.cuh"), os.path.join(_HERE, "csrc", "pmctf_common.cuh")]
INCLUDE = os.path.join(ROOT, "include")

PU_PACKED_FLOATS = 10128
CONV_DEFAULT, CONV_FFMA, CONV_TENSOR = 0, 1, 2
E_TIMEOUT = -4
SRC_PLANE, SRC_WARP, SRC_SKIP3 = 0, 1, 2
MODE_ACCUM, MODE_FILTER, MODE_PU = 0, 1, 2

_f = C.c_float
_fp = C.c_void_p  # device pointers travel as integers


class Plane(C.Structure):
    _fields_ = [("p", C.c_void_p), ("bs", C.c_longlong), ("rs", C.c_longlong), ("cs", C.c_longlong),
                ("gs", C.c_longlong), ("group_n", C.c_int)]


class Step(C.Structure):
    _fields_ = [("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("src_kind", C.c_int), ("mode", C.c_int),
                ("src", Plane), ("src_div1", _f), ("src_div2", _f),
                ("mv", _fp), ("mv_n", C.c_int), ("mv_down", C.c_int), ("mv_sign", _f),
                ("lin_x", _fp), ("lin_y", _fp), ("round_src", C.c_int),
                ("tap0", _f), ("tap1", _f), ("tap2", _f), ("tap_bias", _f),
                ("pu_packed", _fp), ("in_mul", _f), ("post_mul", _f), ("out_mul", _f), ("round_tmp", C.c_int),
                ("base", Plane), ("base_div1", _f), ("base_div2", _f), ("sign", _f), ("final_mul", _f),
                ("out", Plane), ("pred", Plane), ("aux", Plane), ("aux_mul", _f), ("conv_mode", C.c_int)]


class IWave(C.Structure):
    _fields_ = [("tap", (_f * 3) * 4), ("bias", _f * 4), ("pu_packed", _fp), ("scale_l", _f), ("scale_h", _f),
                ("dynamic_range", _f), ("lossy", C.c_int), ("conv_mode", C.c_int)]


class UmmaOp(C.Structure):
    _fields_ = [(n, C.c_uint) for n in ("a_off", "a_lbo", "a_sbo", "b_off", "b_lbo", "b_sbo", "n", "d_col", "accumulate", "a_unsigned")]


class PostProcessD(C.Structure):
    _fields_ = [("conv1_w", _fp), ("conv1_b", _fp), ("res_w", _fp * 12), ("res_b", _fp * 12), ("conv2_w", _fp), ("conv2_b", _fp),
                ("conv3_w", _fp), ("conv3_b", _fp)]


class CtxDcb(C.Structure):
    _fields_ = [(n, _fp) for n in ("dw_w", "dw_b", "pw_w", "pw_b", "ad_w", "ad_b", "f1_w", "f1_b", "f2_w", "f2_b")]


class CtxStep(C.Structure):
    _fields_ = [("x", _fp), ("dec_sym", C.c_void_p), ("scales", _fp), ("means", _fp), ("x_hat", _fp), ("x_q", _fp), ("s_hat", _fp),
                ("x_res", _fp), ("sym16", C.c_void_p), ("idx16", C.c_void_p), ("log_scale_min", _f), ("log_scale_step", _f),
                ("scale_levels", C.c_int), ("step", C.c_int), ("lossy", C.c_int), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int)]


class LLar(C.Structure):
    _fields_ = [("w_in", _fp), ("b_in", _fp), ("w", _fp * 5), ("b", _fp * 5), ("w1", _fp * 2), ("b1", _fp * 2), ("w_out", _fp), ("b_out", _fp),
                ("Y", _fp), ("hist", _fp * 5), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("log_scale_min", _f), ("log_scale_step", _f),
                ("scale_levels", C.c_int)]


class Temporal(C.Structure):
    _fields_ = [("P_t_packed", _fp), ("U_t_packed", _fp), ("scale_p", _f), ("scale_u", _f), ("lossy", C.c_int),
                ("conv_mode", C.c_int)]


# name -> argtypes; every entry returns int except the ones listed in _RESTYPES.  This table is
# also what tests/test_cabi.py checks against include/pmctf_b200.h.
_I, _LL, _P = C.c_int, C.c_longlong, C.c_void_p
SIGNATURES = {
    "pmctf_abi_version": [],
    "pmctf_error_string": [_I],
    "pmctf_launch_count": [],
    "pmctf_set_conv_mode": [_I],
    "pmctf_get_conv_mode": [],
    "pmctf_tc_error_flag": [],
    "pmctf_tc_clear_error": [],
    "pmctf_tc_inject_timeout": [],
    "pmctf_release_pu_weights": [_P],
    "pmctf_tc_debug_times": [_P],
    "pmctf_tc_mma_probe": [_I, _I, _P, _P],
    "pmctf_pack_pu_weights": [_P] * 10,
    "pmctf_flow_warp": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _f, _I, _P],
    "pmctf_chroma_mv_down": [_P, _P, _I, _I, _I, _P],
    "pmctf_lift_step": [C.POINTER(Step), _P],
    "pmctf_predict_update": [_P, _P, _f, _P, _I, _I, _I, _I, _P],
    "pmctf_temporal_filter": [_P, C.POINTER(Temporal), _I, _P, _I, _I, _I, _P],
    "pmctf_forward_mctf": [C.POINTER(Plane), C.POINTER(Plane), _P, _I, _I, _P, _P, C.POINTER(Temporal),
                           C.POINTER(Plane), C.POINTER(Plane), C.POINTER(Plane), C.POINTER(Plane), _I, _I, _I, _P],
    "pmctf_inverse_mctf": [C.POINTER(Plane), C.POINTER(Plane), _P, _I, _I, _P, _P, C.POINTER(Temporal),
                           C.POINTER(Plane), C.POINTER(Plane), _I, _I, _I, _P],
    "pmctf_iwave1d_forward": [C.POINTER(Plane), C.POINTER(IWave), C.POINTER(Plane), C.POINTER(Plane), _I, _I, _I, _P, _LL, _P],
    "pmctf_iwave1d_backward": [C.POINTER(Plane), C.POINTER(Plane), C.POINTER(IWave), C.POINTER(Plane), _I, _I, _I, _P, _LL, _P],
    "pmctf_lift2d_workspace": [_I, _I, _I],
    "pmctf_lift2d_forward": [_P, C.POINTER(IWave), _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _LL, _P],
    "pmctf_lift2d_backward": [_P, _P, _P, _P, C.POINTER(IWave), _P, _I, _I, _I, _P, _LL, _P],
    "pmctf_lift2d_backward_q": [_P, _P, _P, _P, _f, _f, C.POINTER(IWave), _P, _I, _I, _I, _P, _LL, _P],
    "pmctf_quantize": [_P, _f, _f, _I, _I, _P, _LL, _P],
    "pmctf_dequantize": [_P, _f, _I, _P, _LL, _P],
    "pmctf_quantize_stats": [_P, _f, _f, _I, _P, _I, _LL, _P, _P],
    "pmctf_quantize_code": [_P, _P, _f, _I, _I, _P, _P, _LL, _I, _LL, _P, _P],
    "pmctf_unpack_u8": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pmctf_pp_packed_bytes": [_I],
    "pmctf_pp_pack_conv": [_P, _I, _P, _P],
    "pmctf_pp_conv_in": [_P, _P, _P, _f, _P, _P, _I, _I, _I, _P],
    "pmctf_pp_to_bf16": [_P, _P, _LL, _P],
    "pmctf_pp_conv64": [_P, _P, _P, _I, _P, _f, _P, _P, _P, _f, _f, _P, _I, _I, _I, _P],
    "pmctf_postprocess_workspace": [_I, _I],
    "pmctf_postprocess": [_P, C.POINTER(PostProcessD), _f, _f, _P, _I, _I, _I, _P, _LL, _P],
    "pmctf_ctx_packed_bytes": [_I],
    "pmctf_ctx_pack_conv": [_P, _I, _P, _P],
    "pmctf_ctx_conv_in": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_ctx_conv112": [_P, _P, _I, _P, _P, _P, _f, _P, _P, _I, _I, _I, _P],
    "pmctf_ctx_conv112_head": [_P, _P, _I, _P, _P, _P, _f, _P, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_ctx_lower_subband": [_P, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_ctx_dcb_tail": [_P, _P, C.POINTER(CtxDcb), _P, _P, _I, _I, _I, _P],
    "pmctf_ctx_head": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_lstm_gates": [_P, _P, _P, _P, _P, _LL, _P],
    "pmctf_laplace_bits": [_P, _P, _P, _LL, _P],
    "pmctf_ctx_mask_step": [C.POINTER(CtxStep), _P],
    "pmctf_llar_pack": [_P, _I, _I, _P, _P],
    "pmctf_llar_encode": [C.POINTER(LLar), _P, _P, _P, _P],
    "pmctf_llar_forward": [C.POINTER(LLar), _P, _I, _P, _P, _P, _P, _P, _P],
    "pmctf_llar_decode_step": [C.POINTER(LLar), _I, _P, _P, _P, _P],
    "pmctf_llar_decode_band": [C.POINTER(LLar), _P, C.c_longlong, _P, _P, _I, _I, _P, _P, _P, _P],
    "pmctf_rans_decoder_parts": [_P],
    "pmctf_rans_decoder_peek": [_P, _I, _P, _P, _P, _P],
    "pmctf_rans_decoder_seek": [_P, _I, C.c_ulonglong, C.c_longlong],
    "pmctf_pair_packed_bytes": [_I, _I, _I],
    "pmctf_pair_pack_conv": [_P, _I, _I, _I, _I, _I, _P, _P],
    "pmctf_pair_conv": [_P, _P, _P, _I, _I, _I, _I, _f, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_spynet_prep": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pmctf_avgpool2": [_P, _P, _LL, _I, _I, _P],
    "pmctf_pmf_to_quantized_cdf": [_P, _I, _I, _P],
    "pmctf_rans_encoder_create": [_I, _I, _P],
    "pmctf_rans_encoder_destroy": [_P],
    "pmctf_rans_encoder_reset": [_P],
    "pmctf_rans_encode_with_indexes": [_P, _P, _P, _LL, _P, _I, _I, _P, _P],
    "pmctf_rans_encode_chunked": [_P, _P, _P, _LL, _LL, _P, _I, _I, _P, _P],
    "pmctf_rans_encoder_flush": [_P],
    "pmctf_rans_encoded_size": [_P],
    "pmctf_rans_get_encoded_stream": [_P, _P, _LL],
    "pmctf_rans_decoder_create": [_I, _P],
    "pmctf_rans_decoder_destroy": [_P],
    "pmctf_rans_decoder_set_stream": [_P, _P, _LL],
    "pmctf_rans_decode_stream": [_P, _P, _LL, _P, _I, _I, _P, _P, _P],
    "pmctf_gaussian_symbolize": [_P, _P, _LL, _f, _f, _I, _P, _P, _P],
    "pmctf_frame_sse": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "pmctf_conv3x3": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pmctf_conv3x3_fused": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmctf_conv3x3_wgrad": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pmctf_flow_warp_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _f, _P],
    "pmctf_umma_selftest": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P],
}
_RESTYPES = {"pmctf_error_string": C.c_char_p, "pmctf_lift2d_workspace": C.c_longlong, "pmctf_pp_packed_bytes": C.c_longlong,
             "pmctf_postprocess_workspace": C.c_longlong, "pmctf_rans_encoded_size": C.c_longlong, "pmctf_ctx_packed_bytes": C.c_longlong, "pmctf_pair_packed_bytes": C.c_longlong,
             "pmctf_launch_count": C.c_ulonglong}

NVCC_FLAGS = ["--threads", "8", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compile the CUDA sources for sm_100a into lib/libpmctf_b200.so (in-tree, so the built file
    travels with the repository snapshot to the GPU box).  `defines` / `out`: profiling variants (e.g.
    ("PMCTF_TC_TIMING=1",) -> lib/libpmctf_b200_timing.so, loaded by setting PMCTF_LIB)."""
    out = out or LIB_PATH
    newest = max(os.path.getmtime(p) for p in SOURCES + HEADERS + [os.path.join(INCLUDE, "pmctf_b200.h")])
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", INCLUDE, "-o", out, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU or PyTorch fallback for the pMCTF hot path)")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        if L.pmctf_abi_version() != 2:
            raise RuntimeError("libpmctf_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(code: int, what: str):
    """Raise RuntimeError for a non-zero return of the C ABI (PMCTF_ETIMEOUT included: a tensor-core kernel on this device
    gave up earlier and its outputs are incomplete; ops.clear_tc_error() re-arms the device)."""
    if code != 0:
        msg = lib().pmctf_error_string(code)
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")
