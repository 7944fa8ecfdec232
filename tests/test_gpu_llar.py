"""The LL band's autoregressive model on all coefficients at once (csrc/pmctf_llar.cu, pmctf_llar_forward) against its own
sequential form (bit for bit) and against the module's full-plane torch forward (fp32 tolerance): reference
pMCTF/layers/context_fusion.py:143-204, pMCTF/models/pWave.py:531-584."""
import numpy as np
import pytest
import torch

import learned_pmctf_b200 as pkg  # noqa: F401


def _net(dev, seed=3, std=0.04):
    from learned_pmctf_b200.layers.context_fusion import ContextFusionSubband
    torch.manual_seed(seed)
    net = ContextFusionSubband(num_features=128, num_parameters=2, context=False, in_channels=1)
    with torch.no_grad():
        for p in net.parameters():
            p.normal_(0, std if p.dim() == 4 else 0.05)
        net.convs[2].bias[0] += 1.5
    return net.to(dev).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 9, 14), (1, 72, 120), (3, 5, 33), (1, 1, 1), (2, 36, 60)])
def test_parallel_encoder_equals_sequential(shape):
    """symbols, table indexes and the reconstructed band of the layer-parallel encoder == the coefficient-by-coefficient
    kernel's, for ragged tiles (H*W not a multiple of the 32-coefficient tile), a single coefficient and the 1080p band size"""
    dev = torch.device("cuda:0")
    net = _net(dev)
    B, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(11)
    ll = (torch.randn(B, 1, H, W, generator=g) * 6 + 0.3 * torch.rand(B, 1, H, W, generator=g)).to(dev)   # not yet rounded: the kernels round
    with torch.no_grad():
        a_hat, a_sym, a_idx = net.ar_encode(ll, parallel=True)
        assert net.last_encode_path == "parallel"
        b_hat, b_sym, b_idx = net.ar_encode(ll, parallel=False)
        assert net.last_encode_path == "sequential"
    assert torch.equal(a_hat, b_hat)
    assert np.array_equal(a_sym, b_sym)
    assert np.array_equal(a_idx, b_idx)


@pytest.mark.gpu
def test_forward_equals_bitstream_parameters_and_torch_form():
    """ContextFusionSubband.forward in evaluation runs on the same kernels: its (scale, mean) planes reproduce the table indexes
    and symbols of the encoder exactly, and agree with the stock-torch fp32 full-plane form to rounding"""
    dev = torch.device("cuda:0")
    net = _net(dev)
    from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
    em = CompressionModel("laplace")
    em.update()
    B, H, W = 2, 18, 30
    ll = torch.round(torch.randn(B, 1, H, W, device=dev) * 6)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            params = net(ll)                                      # native path
            ll_hat, sym, idx = net.ar_encode(ll, parallel=True)
            with torch.enable_grad():                              # the stock-torch path (taken whenever autograd records)
                ref = net(ll).detach()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert params.shape == (B, 2, H, W)
    scales, means = params.chunk(2, dim=1)
    assert torch.equal(ll_hat, ll)
    want_sym = torch.round(ll - means).reshape(B, H * W).t().reshape(-1).cpu().numpy().astype(np.int16)
    assert np.array_equal(sym, want_sym)
    want_idx = em.gaussian_encoder.build_indexes(scales.cpu()).reshape(B, H * W).t().reshape(-1).numpy().astype(np.int16)
    assert (idx != want_idx).mean() < 0.01 and np.abs(idx.astype(int) - want_idx).max() <= 1    # logf on the device vs torch.log on the host
    err = (params - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.gpu
def test_speculation_failure_falls_back_to_sequential():
    """means of exactly 0.5 make round(y) - mean a tie for every coefficient: round(symbol + mean) differs from round(y) wherever
    ties-to-even rounds down twice, the parallel pass reports it, and ar_encode returns the sequential kernel's result"""
    dev = torch.device("cuda:0")
    net = _net(dev)
    with torch.no_grad():
        net.convs[2].weight.zero_()
        net.convs[2].bias.copy_(torch.tensor([1.0, 0.5]))
    B, H, W = 1, 6, 11
    ll = torch.round(torch.randn(B, 1, H, W, device=dev) * 4)
    with torch.no_grad():
        a_hat, a_sym, a_idx = net.ar_encode(ll)
        assert net.last_encode_path == "sequential"
        b_hat, b_sym, b_idx = net.ar_encode(ll, parallel=False)
    assert torch.equal(a_hat, b_hat) and np.array_equal(a_sym, b_sym) and np.array_equal(a_idx, b_idx)
    assert not torch.equal(a_hat, ll)      # the reconstruction really differs from round(y) here


@pytest.mark.gpu
def test_llar_forward_rejects_bad_arguments():
    import ctypes as C
    from learned_pmctf_b200 import _native as nat
    dev = torch.device("cuda:0")
    net = _net(dev)
    d, _keep = net._ar_desc(1, 4, 4, dev)
    x = torch.zeros(1, 4, 4, device=dev)
    lib = nat.lib()
    st = torch.cuda.current_stream(dev).cuda_stream
    assert lib.pmctf_llar_forward(C.byref(d), x.data_ptr(), 0, None, None, None, None, None, st) != 0      # nothing to produce
    s16 = torch.zeros(16, dtype=torch.int16, device=dev)
    assert lib.pmctf_llar_forward(C.byref(d), x.data_ptr(), 1, s16.data_ptr(), None, None, None, None, st) != 0   # symbols without indexes
    assert lib.pmctf_llar_forward(C.byref(d), None, 0, None, None, x.data_ptr(), None, None, st) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("stream_part", [1, 4])
@pytest.mark.parametrize("shape", [(1, 9, 14), (2, 7, 5), (1, 3, 1), (1, 72, 120), (3, 1, 6)])
def test_band_decoder_round_trip_and_stream_position(shape, stream_part):
    """encode -> rANS -> the one-launch cluster decoder (device-side rANS): the band equals the encoder's reconstruction, equals
    the per-coefficient decoder's, and the host decoder continues correctly behind the band (symbols coded after it)"""
    from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
    dev = torch.device("cuda:0")
    net = _net(dev)
    em = CompressionModel("laplace", ec_thread=False, stream_part=stream_part)
    em.update()
    cdf, ln, off = em.gaussian_encoder.get_cdf_info()
    B, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(17)
    ll = torch.round(torch.randn(B, 1, H, W, generator=g) * 6).to(dev)
    if H * W > 40:
        ll[0, 0, 1, 2] = 700.0          # far outside every table: the escape path with raw digits
        ll[0, 0, 2, 3] = -900.0
    tail_sym = np.array([3, -2, 0, 41, -300, 1], dtype=np.int16)
    tail_idx = np.array([5, 100, 255, 17, 3, 64], dtype=np.int16)
    with torch.no_grad():
        ll_hat, sym16, idx16 = net.ar_encode(ll)
        coder = em.entropy_coder
        for use_band in (True, False):
            coder.reset()
            coder.encoder.encode_with_indexes(sym16, idx16, cdf, ln, off, chunk=B)     # one coefficient per share-out, as pWave does
            coder.encoder.encode_with_indexes(tail_sym, tail_idx, cdf, ln, off)
            coder.flush()
            coder.set_stream(coder.get_encoded_stream())
            dec = coder.decoder
            if use_band:
                back = net.ar_decode_band([B, 1, H, W], dec, cdf, ln, off, dev)
                assert back is not None
            else:
                back = net.ar_decode([B, 1, H, W], lambda i: dec.decode_stream(i, cdf, ln, off), dev)
            assert torch.equal(back, ll_hat), use_band
            assert np.array_equal(np.asarray(dec.decode_stream(tail_idx, cdf, ln, off)), tail_sym), use_band


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(9, 14), (18, 30)])
def test_kernels_equal_cpu_oracle_bit_for_bit(shape):
    """scales, means, symbols and the reconstructed band of the LL kernels == oracle/pmctf_oracle.c:orc_llar_encode (the same fp32
    contract restated on the CPU from the state_dict-layout weights), every float bit for bit"""
    from oracle import oracle as orc
    dev = torch.device("cuda:0")
    net = _net(dev)
    H, W = shape
    g = torch.Generator(device="cpu").manual_seed(23)
    ll = torch.round(torch.randn(1, 1, H, W, generator=g) * 6)
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    o_scale, o_mean, o_sym, o_rec = orc.llar_encode(sd, ll[0, 0].numpy())
    with torch.no_grad():
        params = net(ll.to(dev))                                # layer-parallel kernels
        hat_p, sym_p, _ = net.ar_encode(ll.to(dev), parallel=True)
        hat_s, sym_s, _ = net.ar_encode(ll.to(dev), parallel=False)   # coefficient-by-coefficient kernel
    assert np.array_equal(o_rec, ll[0, 0].numpy())              # no exact tie in this band: the speculated history is the true one
    assert np.array_equal(params[0, 0].cpu().numpy(), o_scale)
    assert np.array_equal(params[0, 1].cpu().numpy(), o_mean)
    assert np.array_equal(sym_p.reshape(H, W), o_sym.astype(np.int16)) and np.array_equal(sym_s, sym_p)
    assert np.array_equal(hat_p[0, 0].cpu().numpy(), o_rec) and torch.equal(hat_s, hat_p)
