"""Entropy-coder boundary: the native rANS coder, the sub-stream container and the CDF builder
against (i) the pure-Python restatement oracle/rans_oracle.py and (ii) the reference's OWN C++ (rans.cpp / py_rans.cpp /
ops.cpp compiled by `make -C oracle ref` against the restated rans64.h) -- streams must be byte-identical, tables equal."""
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import learned_pmctf_b200 as pkg
from learned_pmctf_b200.entropy_models.entropy_models import EntropyCoder, GaussianEncoder
from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
from learned_pmctf_b200.models import MLCodec_CXX, MLCodec_rans
from oracle import rans_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref(name):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + ".*.so"))
    if not hits:
        pytest.skip("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def tables():
    g = GaussianEncoder("laplace")
    g.update(entropy_coder=None)
    return g.get_cdf_info()


def _symbols(n, seed, tables, wild=True):
    cdf, sizes, offs = tables
    r = np.random.default_rng(seed)
    idx = r.integers(0, cdf.shape[0], n).astype(np.int16)
    scale = np.exp(np.linspace(np.log(0.01), np.log(64.0), 256))[idx]
    sym = np.round(r.laplace(0, scale)).clip(-30000, 30000).astype(np.int16)
    if wild:   # escapes: far outside the table on both sides, exactly at the table edges, many-digit raw values
        sym[:8] = [30000, -30000, 17, -17, 255, -256, 4095, -4096]
        idx[:8] = [0, 0, 0, 0, 255, 255, 128, 128]
        sym[8], idx[8] = -offs[3] + sizes[3] - 2 + offs[3] * 2, 3   # value == escape boundary
        idx[9] = -1                                                 # skipped by the encoder
    return sym, idx


@pytest.mark.parametrize("parts,threaded", [(1, False), (1, True), (2, False), (4, True)])
def test_stream_equals_python_oracle_and_round_trips(tables, parts, threaded, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    cdf, sizes, offs = tables
    n = 1200
    a, ai = _symbols(n, 1, tables)
    b, bi = _symbols(n // 2, 2, tables, wild=False)
    enc = MLCodec_rans.RansEncoder(threaded, parts)
    enc.encode_with_indexes(a, ai, cdf, sizes, offs)
    enc.encode_with_indexes(b, bi, cdf, sizes, offs)
    enc.flush()
    ours = enc.get_encoded_stream().tobytes()
    assert ours == ro.encode([(a, ai, cdf, sizes, offs), (b, bi, cdf, sizes, offs)], parts)
    # decoding: the skipped symbol cannot come back, decode with a valid index there
    keep = ai >= 0
    # the container splits by position, so decode per part with the same split and drop the skipped position
    dec = MLCodec_rans.RansDecoder(parts)
    dec.set_stream(np.frombuffer(ours, dtype=np.uint8))
    if parts == 1:
        out = dec.decode_stream(ai[keep], cdf, sizes, offs)
        assert np.array_equal(out, a[keep])
        assert np.array_equal(dec.decode_stream(bi, cdf, sizes, offs), b)
        od = ro.Decoder(ours, parts)
        assert np.array_equal(od.decode(ai[keep], cdf, sizes, offs), a[keep])
    # reset + reuse gives the same bytes again
    enc.reset()
    enc.encode_with_indexes(a, ai, cdf, sizes, offs)
    enc.encode_with_indexes(b, bi, cdf, sizes, offs)
    enc.flush()
    assert enc.get_encoded_stream().tobytes() == ours


@pytest.mark.parametrize("parts", [1, 2, 4])
def test_multi_part_round_trip(tables, parts, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    cdf, sizes, offs = tables
    a, ai = _symbols(4001, 3, tables)
    ai[9] = 5
    enc = MLCodec_rans.RansEncoder(False, parts)
    enc.encode_with_indexes(a, ai, cdf, sizes, offs)
    enc.flush()
    s = enc.get_encoded_stream()
    assert s[0] == ((parts - 1) << 4) + 1
    dec = MLCodec_rans.RansDecoder(parts)
    dec.set_stream(s)
    assert np.array_equal(dec.decode_stream(ai, cdf, sizes, offs), a)
    assert np.array_equal(ro.Decoder(s.tobytes(), parts).decode(ai, cdf, sizes, offs), a)


def test_empty_and_rejected_arguments(tables, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    cdf, sizes, offs = tables
    enc = MLCodec_rans.RansEncoder(False, 1)
    enc.flush()
    s = enc.get_encoded_stream()
    assert s.size == 1 + 8 and s.tobytes() == ro.encode([], 1)     # flag byte + the 64-bit initial state
    enc.encode_with_indexes(np.zeros(3, np.int16), np.array([0, 300, 0], np.int16), cdf, sizes, offs)   # index beyond the tables
    enc.flush()
    with pytest.raises(RuntimeError):
        enc.get_encoded_stream()
    dec = MLCodec_rans.RansDecoder(2)
    with pytest.raises(RuntimeError):
        dec.set_stream(s)                                          # a 1-part stream in a 2-part decoder
    with pytest.raises(RuntimeError):
        MLCodec_rans.RansDecoder(1).decode_stream(np.zeros(1, np.int16), cdf, sizes, offs)   # no stream set
    d1 = MLCodec_rans.RansDecoder(1)
    d1.set_stream(s)
    with pytest.raises(RuntimeError):                              # reading past the end of the stream is an error, not garbage
        d1.decode_stream(np.full(100000, 255, np.int16), cdf, sizes, offs)
    with pytest.raises(RuntimeError):
        MLCodec_rans.RansEncoder(False, 17)


def test_pmf_to_quantized_cdf_vs_oracle(conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    r = np.random.default_rng(0)
    for n in (1, 2, 7, 64, 103):
        p = r.random(n).astype(np.float32) ** 8          # many near-zero entries: the widening loop runs
        p /= p.sum()
        ours = MLCodec_CXX.pmf_to_quantized_cdf(p.tolist(), 16)
        assert ours == ro.pmf_to_quantized_cdf(p.tolist(), 16)
        assert ours[0] == 0 and ours[-1] == 65536 and all(b > a for a, b in zip(ours, ours[1:]))


def test_against_reference_cpp(tables, conv_mode):
    """Byte-identical streams and cross-decoding against the reference's own rans.cpp / py_rans.cpp; equal CDF tables
    against its ops.cpp."""
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    ref_rans, ref_cxx = _ref("MLCodec_rans"), _ref("MLCodec_CXX")
    cdf, sizes, offs = tables
    r = np.random.default_rng(4)
    for n in (3, 64, 103):
        p = r.random(n).astype(np.float32) ** 6
        p /= p.sum()
        assert MLCodec_CXX.pmf_to_quantized_cdf(p.tolist(), 16) == list(ref_cxx.pmf_to_quantized_cdf(p.tolist(), 16))
    # Only the reference's single-threaded encoder is driven: its RansEncoderLibMultiThread starts the worker thread in the
    # constructor's initialiser list BEFORE the mutexes, condition variables and the task list it uses are constructed
    # (rans.h:106-113 member order), and about one run in three never wakes up again.  Multi-part containers are checked
    # through the reference's DECODER (no threads at construction) and per sub-stream against its single-part encoder.
    for n in (5000, 777):
        a, ai = _symbols(n, 10 + n, tables)
        ai[9] = 7
        b, bi = _symbols(n, 20 + n, tables, wild=False)
        ours, theirs = MLCodec_rans.RansEncoder(True, 1), ref_rans.RansEncoder(False, 1)
        for e in (ours, theirs):
            e.encode_with_indexes(a, ai, cdf, sizes, offs)
            e.encode_with_indexes(b, bi, cdf, sizes, offs)
            e.flush()
        so, st = ours.get_encoded_stream(), theirs.get_encoded_stream()
        assert so.tobytes() == st.tobytes()
        d_ours, d_theirs = MLCodec_rans.RansDecoder(1), ref_rans.RansDecoder(1)
        d_ours.set_stream(st)
        d_theirs.set_stream(so)
        for d in (d_ours, d_theirs):
            assert np.array_equal(d.decode_stream(ai, cdf, sizes, offs), a)
            assert np.array_equal(d.decode_stream(bi, cdf, sizes, offs), b)
    for parts in (2, 4):
        n = 4000
        a, ai = _symbols(n, 30 + parts, tables)
        ai[9] = 7
        ours = MLCodec_rans.RansEncoder(True, parts)
        ours.encode_with_indexes(a, ai, cdf, sizes, offs)
        ours.flush()
        so = ours.get_encoded_stream()
        d = ref_rans.RansDecoder(parts)
        d.set_stream(so)
        assert np.array_equal(d.decode_stream(ai, cdf, sizes, offs), a)
        pos = 1 + 2 * (parts - 1)
        for i in range(parts):       # every sub-stream equals the reference's single-part stream of that slice
            one = ref_rans.RansEncoder(False, 1)
            one.encode_with_indexes(a[i * n // parts:(i + 1) * n // parts], ai[i * n // parts:(i + 1) * n // parts], cdf, sizes, offs)
            one.flush()
            body = one.get_encoded_stream().tobytes()[1:]
            assert so.tobytes()[pos:pos + len(body)] == body
            pos += len(body)
        assert pos == so.size


def test_gaussian_tables_equal_reference(conv_mode):
    """GaussianEncoder.update(): the 256 Laplace tables equal the ones the reference's entropy_models.py builds when its two
    extension modules are aliased to ours (build container only)."""
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    if not os.path.isdir("/root/reference/pMCTF"):
        pytest.skip("reference tree not present")
    sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference"]
    saved = {k: sys.modules.get(k) for k in ("pMCTF.models.MLCodec_rans", "pMCTF.models.MLCodec_CXX")}
    try:
        sys.modules["pMCTF.models.MLCodec_rans"] = MLCodec_rans
        sys.modules["pMCTF.models.MLCodec_CXX"] = MLCodec_CXX
        from pMCTF.entropy_models import entropy_models as rem
        for dist in ("laplace", "gaussian"):
            theirs = rem.GaussianEncoder(dist)
            theirs.update(entropy_coder=rem.EntropyCoder(False, 1))
            ours = GaussianEncoder(dist)
            ours.update(entropy_coder=EntropyCoder(False, 1))
            for x, y in zip(ours.get_cdf_info(), theirs.get_cdf_info()):
                assert np.array_equal(x, y)
            sc = torch.rand(1, 1, 16, 24) * 3
            sym = torch.round(torch.randn(1, 1, 16, 24) * 3)
            for g in (ours, theirs):
                g.entropy_coder.reset()
                g.encode(sym, sc)
                g.entropy_coder.flush()
            assert ours.entropy_coder.get_encoded_stream() == theirs.entropy_coder.get_encoded_stream()
            theirs.entropy_coder.set_stream(ours.entropy_coder.get_encoded_stream())
            assert torch.equal(theirs.decode_stream(sc, torch.float32, "cpu"), sym)
    finally:
        del sys.path[:2]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_compression_model_bits(conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    em = CompressionModel("laplace")
    y, s = torch.tensor([0.0, 1.0, -3.0]), torch.tensor([0.5, 1.0, 2.0])
    d = torch.distributions.laplace.Laplace(torch.zeros(3), s)
    want = torch.clamp_min(-torch.log(d.cdf(y + 0.5) - d.cdf(y - 0.5) + 1e-5) / np.log(2.0), 0)
    assert torch.allclose(em.get_y_laplace_bits(y, s), want)
    r, q, h = em.process(torch.tensor([1.4, -2.6]), torch.tensor([0.25, 0.5]))
    assert torch.equal(q, torch.tensor([1.0, -3.0])) and torch.equal(h, r + torch.tensor([0.25, 0.5]))


@pytest.mark.gpu
def test_device_staging_round_trip(conv_mode):
    """GaussianEncoder on CUDA tensors: one symbolise pass + one synchronisation per coded step; symbols come back exactly,
    indexes equal the reference formula evaluated by torch on the same device."""
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    dev = torch.device("cuda:0")
    em = CompressionModel("laplace", ec_thread=True, stream_part=4)
    em.update()
    g = torch.Generator(device="cpu").manual_seed(5)
    scales = (torch.rand(2, 1, 90, 131, generator=g) * 8).to(dev)
    scales[0, 0, 0, :6] = torch.tensor([0.0, 1e-7, 0.01, 64.0, 1e3, 0.0099], device=dev)
    mask = (torch.arange(131, device=dev) % 2 == 0).float()
    scales = scales * mask                                        # the four-step model zeroes the scales outside its mask
    sym = torch.round(torch.distributions.laplace.Laplace(0.0, scales.cpu() + 0.01).sample()).to(dev) * mask
    sym[1, 0, 5, :4] = torch.tensor([9000.0, -9000.0, 31000.0, -31000.0], device=dev) * mask[:4]
    s16, i16 = em.gaussian_encoder.stage(sym, scales)
    # the kernel evaluates the reference formula with a true fp32 division (what torch does on the CPU); torch's CUDA kernels
    # multiply by the reciprocal of a scalar divisor instead, and the two logf differ by an ulp here and there: an index may
    # differ by one where the scaled logarithm sits on an integer.  Encoder and decoder share the kernel, so streams decode.
    want_idx = em.gaussian_encoder.build_indexes(scales.cpu()).reshape(-1).numpy().astype(np.int16)
    diff = np.abs(i16.astype(np.int32) - want_idx.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, (int(diff.max()), float((diff != 0).mean()))
    assert np.array_equal(s16, sym.clamp(-30000, 30000).to(torch.int16).reshape(-1).cpu().numpy())
    em.entropy_coder.reset()
    em.gaussian_encoder.encode(sym, scales)
    em.entropy_coder.flush()
    stream = em.entropy_coder.get_encoded_stream()
    em.entropy_coder.set_stream(stream)
    back = em.gaussian_encoder.decode_stream(scales, torch.float32, dev)
    assert torch.equal(back, sym.clamp(-30000, 30000))


@pytest.mark.parametrize("parts,chunk", [(1, 1), (4, 1), (4, 3), (3, 7), (2, 1000)])
def test_chunked_encode_equals_consecutive_calls(tables, parts, chunk, conv_mode):
    """encode_with_indexes(..., chunk=c) produces the byte-identical stream of ceil(n / c) consecutive calls (what the LL band's
    one-call encoder must look like to a decoder that asks for one coefficient per call, pWave.py:548-553), and a decoder reading
    `chunk` symbols per call gets the symbols back"""
    if conv_mode != "tensor":
        pytest.skip("independent of the convolution arithmetic")
    cdf, sizes, offs = tables
    n = 211
    sym, idx = _symbols(n, 5, tables, wild=False)
    one = MLCodec_rans.RansEncoder(False, parts)
    one.encode_with_indexes(sym, idx, cdf, sizes, offs, chunk=chunk)
    one.flush()
    many = MLCodec_rans.RansEncoder(False, parts)
    for c0 in range(0, n, chunk):
        many.encode_with_indexes(sym[c0:c0 + chunk], idx[c0:c0 + chunk], cdf, sizes, offs)
    many.flush()
    stream = one.get_encoded_stream()
    assert np.array_equal(stream, many.get_encoded_stream())
    dec = MLCodec_rans.RansDecoder(parts)
    dec.set_stream(stream)
    back = np.concatenate([dec.decode_stream(idx[c0:c0 + chunk], cdf, sizes, offs) for c0 in range(0, n, chunk)])
    assert np.array_equal(back, sym)
    # the reader position survives a peek / seek round trip (the hand-over to the device-side decoder)
    import ctypes as C
    from learned_pmctf_b200 import _native as nat
    lib = nat.lib()
    assert lib.pmctf_rans_decoder_parts(dec._h) == parts
    dec.set_stream(stream)
    first = dec.decode_stream(idx[:chunk], cdf, sizes, offs)
    saved = []
    for p in range(parts):
        x, pos, nw, w = C.c_ulonglong(), C.c_longlong(), C.c_longlong(), C.c_void_p()
        assert lib.pmctf_rans_decoder_peek(dec._h, p, C.byref(x), C.byref(pos), C.byref(nw), C.byref(w)) == 0
        saved.append((x.value, pos.value))
    second = dec.decode_stream(idx[chunk:2 * chunk], cdf, sizes, offs)
    for p, (x, pos) in enumerate(saved):
        assert lib.pmctf_rans_decoder_seek(dec._h, p, C.c_ulonglong(x), pos) == 0
    assert np.array_equal(dec.decode_stream(idx[chunk:2 * chunk], cdf, sizes, offs), second)
    assert np.array_equal(first, sym[:chunk])
    assert lib.pmctf_rans_decoder_seek(dec._h, parts, C.c_ulonglong(0), 2) != 0
