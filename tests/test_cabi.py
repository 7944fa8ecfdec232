"""The C-ABI library loads and exports every symbol include/pmctf_b200.h declares, with the argument counts the
ctypes table binds (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import learned_pmctf_b200 as pkg

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pmctf_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|unsigned long long|const char \*)\s*(pmctf_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_declares_the_whole_path():
    fns = declared_functions()
    for name in ("pmctf_flow_warp", "pmctf_chroma_mv_down", "pmctf_predict_update", "pmctf_temporal_filter",
                 "pmctf_forward_mctf", "pmctf_inverse_mctf", "pmctf_iwave1d_forward", "pmctf_iwave1d_backward",
                 "pmctf_lift2d_forward", "pmctf_lift2d_backward", "pmctf_lift2d_backward_q", "pmctf_quantize",
                 "pmctf_dequantize", "pmctf_quantize_stats", "pmctf_unpack_u8", "pmctf_frame_sse", "pmctf_lift_step"):
        assert name in fns, name


def test_library_exports_every_declared_symbol():
    path = pkg._native.build()
    lib = ctypes.CDLL(path)
    fns = declared_functions()
    assert len(fns) >= 20
    for name, nargs in fns.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in pkg._native.SIGNATURES, f"{name} has no ctypes signature"
        assert len(pkg._native.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding has {len(pkg._native.SIGNATURES[name])}"
    assert set(pkg._native.SIGNATURES) == set(fns), "binding table and header disagree"
    assert lib.pmctf_abi_version() == 2
    lib.pmctf_error_string.restype = ctypes.c_char_p
    assert b"invalid argument" in lib.pmctf_error_string(-1)
    assert b"gave up" in lib.pmctf_error_string(-4)          # PMCTF_ETIMEOUT


def test_struct_layouts_match_header():
    # sizes implied by the header's field lists on LP64
    n = pkg._native
    assert ctypes.sizeof(n.Plane) == 48
    assert ctypes.sizeof(n.Temporal) == 32                   # 2 pointers, 2 floats, lossy, conv_mode
    assert ctypes.sizeof(n.IWave) == 4 * 3 * 4 + 4 * 4 + 8 + 4 * 4 + 8   # + conv_mode (+ tail padding)
    assert n.Step.conv_mode.offset == n.Step.aux_mul.offset + 4 and n.IWave.conv_mode.offset == 88 and n.Temporal.conv_mode.offset == 28
    assert (n.CONV_DEFAULT, n.CONV_FFMA, n.CONV_TENSOR) == (0, 1, 2)   # a zeroed descriptor means "process default"
    assert n.PU_PACKED_FLOATS == 10128


def test_no_cpu_fallback():
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        pkg.flow_warp(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.ops.unpack_u8(torch.zeros(1, 8, 8, dtype=torch.uint8), 8, 8)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.abspath(pkg.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libpmctf_oracle" not in txt, f
