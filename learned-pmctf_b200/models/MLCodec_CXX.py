"""Drop-in for the reference's pybind11 module `pMCTF.models.MLCodec_CXX` (pMCTF/cpp/ops/ops.cpp:84-91)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _native as nat


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """Probability mass function (sequence of floats) -> list of len(pmf) + 1 cumulative frequencies with every interval
    at least one count wide and the last entry exactly 2**precision."""
    p = np.ascontiguousarray(pmf, dtype=np.float32).reshape(-1)
    out = np.empty(p.size + 1, dtype=np.uint32)
    nat.check(nat.lib().pmctf_pmf_to_quantized_cdf(p.ctypes.data_as(C.c_void_p), p.size, int(precision), out.ctypes.data_as(C.c_void_p)),
              "pmf_to_quantized_cdf")
    return out.tolist()
