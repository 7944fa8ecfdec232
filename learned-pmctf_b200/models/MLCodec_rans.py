"""Drop-in for the reference's pybind11 module `pMCTF.models.MLCodec_rans` (pMCTF/cpp/py_rans/py_rans.cpp:229-243): the same two
classes with the same numpy-array methods, bound with ctypes to the C ABI of include/pmctf_b200.h (csrc/pmctf_rans.cu).
A maintainer aliases it with `sys.modules["pMCTF.models.MLCodec_rans"] = learned_pmctf_b200.models.MLCodec_rans`
 -- the reference's own extension cannot be built without network access to ryg_rans."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _native as nat


def _tables(cdfs, cdfs_sizes, offsets):
    cdfs = np.ascontiguousarray(cdfs, dtype=np.int32)
    sizes = np.ascontiguousarray(cdfs_sizes, dtype=np.int32).reshape(-1)
    offs = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
    if cdfs.ndim != 2 or cdfs.shape[0] != sizes.size or offs.size != sizes.size:
        raise RuntimeError("cdfs must be [num, width] with one size and one offset per row")
    return cdfs, sizes, offs


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RansEncoder:
    def __init__(self, multiThread: bool, streamPart: int = 1):
        h = C.c_void_p()
        nat.check(nat.lib().pmctf_rans_encoder_create(int(bool(multiThread)), int(streamPart), C.byref(h)), "rans_encoder_create")
        self._h, self._destroy = h, nat.lib().pmctf_rans_encoder_destroy   # bound now: module globals may be gone at shutdown

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._destroy(h)

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets, chunk=None) -> None:
        """chunk: behave as consecutive calls of `chunk` symbols each (every chunk shared out over the sub-streams on its own)"""
        sym = np.ascontiguousarray(symbols, dtype=np.int16).reshape(-1)
        idx = np.ascontiguousarray(indexes, dtype=np.int16).reshape(-1)
        if sym.size != idx.size:
            raise RuntimeError("one table index per symbol")
        cdfs, sizes, offs = _tables(cdfs, cdfs_sizes, offsets)
        if chunk is not None:
            nat.check(nat.lib().pmctf_rans_encode_chunked(self._h, _ptr(sym), _ptr(idx), sym.size, int(chunk), _ptr(cdfs), cdfs.shape[0],
                                                          cdfs.shape[1], _ptr(sizes), _ptr(offs)), "rans_encode_chunked")
            return
        nat.check(nat.lib().pmctf_rans_encode_with_indexes(self._h, _ptr(sym), _ptr(idx), sym.size, _ptr(cdfs), cdfs.shape[0],
                                                           cdfs.shape[1], _ptr(sizes), _ptr(offs)), "rans_encode_with_indexes")

    def flush(self) -> None:
        nat.check(nat.lib().pmctf_rans_encoder_flush(self._h), "rans_encoder_flush")

    def get_encoded_stream(self) -> np.ndarray:
        n = nat.lib().pmctf_rans_encoded_size(self._h)
        if n < 0:
            nat.check(int(n), "rans_encoded_size")
        out = np.empty(int(n), dtype=np.uint8)
        nat.check(nat.lib().pmctf_rans_get_encoded_stream(self._h, _ptr(out), out.size), "rans_get_encoded_stream")
        return out

    def reset(self) -> None:
        nat.check(nat.lib().pmctf_rans_encoder_reset(self._h), "rans_encoder_reset")


class RansDecoder:
    def __init__(self, streamPart: int = 1):
        h = C.c_void_p()
        nat.check(nat.lib().pmctf_rans_decoder_create(int(streamPart), C.byref(h)), "rans_decoder_create")
        self._h, self._destroy = h, nat.lib().pmctf_rans_decoder_destroy

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._destroy(h)

    def set_stream(self, encoded) -> None:
        buf = np.ascontiguousarray(np.frombuffer(encoded, dtype=np.uint8) if isinstance(encoded, (bytes, bytearray)) else encoded,
                                   dtype=np.uint8)
        nat.check(nat.lib().pmctf_rans_decoder_set_stream(self._h, _ptr(buf), buf.size), "rans_decoder_set_stream")

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets) -> np.ndarray:
        idx = np.ascontiguousarray(indexes, dtype=np.int16).reshape(-1)
        cdfs, sizes, offs = _tables(cdfs, cdfs_sizes, offsets)
        out = np.empty(idx.size, dtype=np.int16)
        nat.check(nat.lib().pmctf_rans_decode_stream(self._h, _ptr(idx), idx.size, _ptr(cdfs), cdfs.shape[0], cdfs.shape[1],
                                                     _ptr(sizes), _ptr(offs), _ptr(out)), "rans_decode_stream")
        return out
