"""The coder around the transform (SURVEY.md section 8f rows 1 and 3 together): module tree / state_dict parity of the full pWave
with the reference's (build container), and on the GPU forward (rate estimate) and compress -> file -> decompress."""
import os
import sys

import numpy as np
import pytest
import torch

import learned_pmctf_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _randomise(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k in ("QP", "QP_ll"):
                p.copy_(torch.tensor([1 / 32, 1 / 2]).view(2, 1, 1, 1))
            elif p.dim() == 4 and p.shape[0] == 112 and p.shape[1] == 112:
                p.copy_(0.022 * torch.randn(p.shape, generator=g))
            elif p.dim() == 4 and p.shape[0] == 128 and p.shape[1] == 128:
                p.copy_(0.03 * torch.randn(p.shape, generator=g))
            elif p.dim() == 4 and p.shape[0] == 2:
                p.copy_(0.03 * torch.randn(p.shape, generator=g))
            elif p.dim() == 4 and "dequantModule" in k:
                p.copy_((1e-4 if p.shape[0] == 1 else 0.04) * torch.randn(p.shape, generator=g))
            elif p.dim() == 4:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
                if p.numel() == 2:
                    p[0] += 1.5
    return m


def test_full_pwave_state_dict_equals_reference(conv_mode):
    """Every key and shape of the reference's pWave (transform, quantiser, long-term context, LL model, 12 four-step models,
    PostProcess) exists in ours and loads strictly (build container only: needs the reference tree)."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    if not os.path.isdir("/root/reference/pMCTF"):
        pytest.skip("reference tree not present")
    sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference"]
    try:
        from pMCTF.models.pWave import pWave as RefPWave
        ref = RefPWave()
        ours = pkg.pWave(entropy_model=True)
        sd_r, sd_o = ref.state_dict(), ours.state_dict()
        assert set(sd_r) == set(sd_o), (sorted(set(sd_r) - set(sd_o))[:5], sorted(set(sd_o) - set(sd_r))[:5])
        assert all(tuple(sd_r[k].shape) == tuple(sd_o[k].shape) for k in sd_r)
        ours.load_state_dict(sd_r, strict=True)
        # the host-side parts (LL model, long-term context) reproduce the reference's outputs exactly on the CPU
        x = torch.round(torch.randn(1, 1, 8, 12) * 5)
        with torch.no_grad():
            top = str(ref.decomp_levels - 1)
            assert torch.equal(ref.context_fusion[top]["ll"](x), ours.context_fusion[top]["ll"](x))
            for o in (ref.context_prediction, ours.context_prediction):
                o.init_sequential(list(x.size()), x.device)
            assert torch.equal(ref.context_prediction.forward_one_subband(x, "ll", 3)["context"],
                               ours.context_prediction.forward_one_subband(x, "ll", 3)["context"])
    finally:
        del sys.path[:2]


@pytest.mark.gpu
def test_forward_rate_and_compress_decompress(tmp_path, conv_mode):
    """pWave(entropy_model=True) on the GPU: forward() fills the reference's return keys with a finite rate; compress() writes a
    stream whose size agrees with that estimate, decompress() of the file reproduces compress()'s x_hat bit for bit (the decoder
    recomputes every parameter with the same kernels from what it has decoded)."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = _randomise(pkg.pWave(entropy_model=True)).to(dev).eval()
    m.update()
    g = np.random.default_rng(3)
    base = g.random((1, 1, 128 + 8, 192 + 8)) * 255
    base = sum(np.roll(base, (dy, dx), (2, 3)) for dy in range(-2, 3) for dx in range(-2, 3)) / 25.0
    x = torch.from_numpy(np.round(base[:, :, 4:132, 4:196]).astype(np.float32)).to(dev)
    with torch.no_grad():
        out = m(x, q_index=12)
    assert set(out) == {"x_hat", "bits", "likelihoods", "subbands", "bpp_total", "bits_total", "mse"}
    assert torch.isfinite(out["bpp_total"]) and out["bpp_total"].item() > 0 and torch.isfinite(out["x_hat"]).all()
    path = str(tmp_path / "plane.bin")
    x_hat = m.compress(x, sideinfo=(1, 1, 128, 192), file_name=path, q_index=12)
    size_bits = 8 * (os.path.getsize(path) - 16)
    est = float(out["bits_total"])
    assert 0.8 * est - 200 < size_bits < 1.25 * est + 400, (size_bits, est)
    back = m.decompress(path, padding=64, q_index=12)["x_hat"]
    assert torch.equal(back, x_hat)
    assert torch.equal(x_hat, out["x_hat"])       # same symbols through the forward pass and the bitstream path
    assert pkg.ops.tc_error_flag() == 0


@pytest.mark.gpu
def test_ll_sequential_kernel_vs_torch_form(conv_mode):
    """The one-kernel sequential LL model against the module's own forward_sequential (the reference's formulation on torch ops):
    same symbols and table indexes except where fp32 summation order moves a value across a rounding / table boundary; encoder and
    decoder of the kernel agree exactly."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
    from learned_pmctf_b200.layers.context_fusion import ContextFusionSubband
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = ContextFusionSubband(num_features=128, num_parameters=2, context=False, in_channels=1)
    with torch.no_grad():
        for p in net.parameters():
            p.normal_(0, 0.04 if p.dim() == 4 else 0.05)
        net.convs[2].bias[0] += 1.5
    net = net.to(dev).eval()
    em = CompressionModel("laplace")
    em.update()
    B, H, W = 2, 9, 14
    ll = torch.round(torch.randn(B, 1, H, W, device=dev) * 6)
    with torch.no_grad():
        ll_hat, sym16, idx16 = net.ar_encode(ll)
        # the reference's loop on torch ops, conditioning on the same reconstruction
        plane = torch.nn.functional.pad(ll_hat, (1, 1, 1, 1))
        want_sym, want_idx = [], []
        for h in range(H):
            for w in range(W):
                scale, mean = net.forward_sequential(plane, h, w).chunk(2, dim=1)
                want_sym.append(torch.round(ll[:, :, h:h + 1, w:w + 1] - mean).reshape(-1))
                want_idx.append(em.gaussian_encoder.build_indexes(scale.cpu()).reshape(-1))
        net.sequential_init = False
    want_sym = torch.stack(want_sym).reshape(-1).cpu().numpy().astype(np.int16)
    want_idx = torch.stack(want_idx).reshape(-1).numpy().astype(np.int16)
    assert (sym16 != want_sym).mean() < 0.02 and np.abs(sym16.astype(int) - want_sym).max() <= 1
    assert (idx16 != want_idx).mean() < 0.05 and np.abs(idx16.astype(int) - want_idx).max() <= 1
    # encode -> rANS -> decode with the kernel on both sides: exact
    cdf, ln, off = em.gaussian_encoder.get_cdf_info()
    em.entropy_coder.reset()
    em.entropy_coder.encoder.encode_with_indexes(sym16, idx16, cdf, ln, off)
    em.entropy_coder.flush()
    em.entropy_coder.set_stream(em.entropy_coder.get_encoded_stream())
    dec = em.entropy_coder.decoder
    back = net.ar_decode([B, 1, H, W], lambda i: dec.decode_stream(i, cdf, ln, off), dev)
    assert torch.equal(back, ll_hat)


@pytest.mark.gpu
def test_fused_glue_kernels_equal_torch_formulas(conv_mode):
    """ConvLSTM gate kernel and Laplace rate kernel against the torch compositions they replace."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
    from learned_pmctf_b200.layers.long_context import LSTM2D
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    cell = LSTM2D(1, 32).to(dev)
    x, h, c = torch.randn(2, 1, 20, 30, device=dev), torch.randn(2, 32, 20, 30, device=dev), torch.randn(2, 32, 20, 30, device=dev)
    with torch.no_grad():
        h1, c1 = cell(x, h, c)
    with torch.enable_grad():
        h2, c2 = cell(x, h, c)
    assert (h1 - h2).abs().max().item() < 1e-5 and (c1 - c2).abs().max().item() < 1e-5
    em = CompressionModel("laplace")
    y = torch.round(torch.randn(2, 1, 40, 50, device=dev) * 5)
    s = torch.rand(2, 1, 40, 50, device=dev) * 4 - 0.2            # includes non-positive scales (clamped to 1e-5)
    with torch.no_grad():
        got = em.get_y_laplace_bits(y, s)
    want = CompressionModel._interval_bits(torch.distributions.laplace.Laplace, y, s)
    assert (got - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
