"""Four-step entropy-parameter network on the GPU: the CTA-pair tcgen05 convolution against a
bf16-operand emulation, the whole module against the fp32 oracle pinned to the reference (parameter errors + final-symbol
mismatch COUNTS), and compress -> rANS -> decompress round trips through the native coder."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ctx_weights  # noqa: E402

import learned_pmctf_b200 as pkg  # noqa: E402
from learned_pmctf_b200 import _native as nat  # noqa: E402
from learned_pmctf_b200.layers.context_fusion_4step import ContextFusionFourStep, _Features  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_bf16_planar(x):   # [N,112,H,W] fp32 -> [N,14,H,W,8] bf16
    N, Cc, H, W = x.shape
    return x.view(N, Cc // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def to_f32_planar(x):    # [N,112,H,W] -> [N,28,H,W,4]
    N, Cc, H, W = x.shape
    return x.view(N, Cc // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()


def from_f32_planar(t):
    N, G, H, W, _ = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(N, G * 4, H, W)


def from_bf16_planar(t):
    N, G, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(N, G * 8, H, W)


@pytest.mark.parametrize("N,H,W,taps,res,slope", [(1, 4, 30, 9, 0, 1.0), (1, 8, 60, 9, 1, 1.0), (2, 11, 47, 9, 2, 0.2), (1, 36, 60, 9, 1, 1.0),
                                                   (3, 5, 31, 1, 0, 0.01), (1, 72, 120, 9, 2, 1.0), (1, 144, 240, 9, 0, 0.2)])
def test_conv112_vs_bf16_emulation(N, H, W, taps, res, slope, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    g = torch.Generator().manual_seed(N * 1000 + H * 10 + W + taps)
    k = 3 if taps == 9 else 1
    x = torch.randn(N, 112, H, W, generator=g)
    w = torch.randn(112, 112, k, k, generator=g) * 0.05
    b = torch.randn(112, generator=g) * 0.1
    r1 = torch.randn(N, 112, H, W, generator=g) if res >= 1 else None
    r2 = torch.randn(N, 112, H, W, generator=g) if res >= 2 else None
    xb, wb = x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double()
    want = F.conv2d(xb, wb, b.double(), padding=k // 2)
    if r1 is not None:
        want = want + r1.double()
    if r2 is not None:
        want = want + r2.double()
    want = torch.where(want >= 0, want, want * slope).float()

    lib = nat.lib()
    xd = to_bf16_planar(x).to(DEV)
    wd, bd = w.to(DEV).contiguous(), b.to(DEV)
    packed = torch.empty(int(lib.pmctf_ctx_packed_bytes(taps)), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    nat.check(lib.pmctf_ctx_pack_conv(wd.data_ptr(), taps, packed.data_ptr(), st), "pack")
    r1d = to_f32_planar(r1).to(DEV) if r1 is not None else None
    r2d = to_f32_planar(r2).to(DEV) if r2 is not None else None
    of = torch.full((N, 28, H, W, 4), float("nan"), device=DEV)
    ob = torch.zeros((N, 14, H, W, 8), dtype=torch.bfloat16, device=DEV)
    nat.check(lib.pmctf_ctx_conv112(xd.data_ptr(), packed.data_ptr(), taps, bd.data_ptr(), r1d.data_ptr() if r1d is not None else None,
                                    r2d.data_ptr() if r2d is not None else None, slope, of.data_ptr(), ob.data_ptr(), N, H, W, st), "conv112")
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    got = from_f32_planar(of).cpu()
    err = (got - want).abs().max().item()
    assert err < 2e-4 * max(1.0, want.abs().max().item()), err
    gotb = from_bf16_planar(ob).cpu()
    assert torch.equal(gotb, got.to(torch.bfloat16).float())


def _load(tag, golden):
    g = golden("ctx4")
    cc = int(g[f"{tag}.ctx_channels"])
    w = ctx_weights.make(int(g[f"{tag}.seed"]), cc)
    m = ContextFusionFourStep(ctx_channels=cc).to(DEV).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    prev = t(g[f"{tag}.prev"]) if cc == 2 else None
    return g, w, m, t(g[f"{tag}.x"]), t(g[f"{tag}.context"]), prev


@pytest.mark.parametrize("tag", ["a", "b"])
def test_four_step_vs_oracle_and_reference(tag, golden, conv_mode, capsys):
    """Parameters within the bf16-operand tolerance of the fp32 oracle; the count of final symbols round(x - mean) that differ from
    the reference's is reported and bounded (a symbol flips only where x - mean lies within the mean's error of a .5 boundary)."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    from oracle import ctx_oracle as co
    g, w, m, x, context, prev = _load(tag, golden)
    with torch.no_grad():
        x_res, x_q, x_hat, s_hat = m(x, context=context, prev_subband=prev)
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    oracle = co.FourStep(w)
    o_res, o_q, o_hat, o_s = oracle.forward(g[f"{tag}.x"], g[f"{tag}.context"], g[f"{tag}.prev"] if prev is not None else None)
    # the oracle itself reproduces the reference's symbols on this fixture
    assert np.array_equal(o_q, g[f"{tag}.x_q"])
    d_mean = np.abs((x.cpu().numpy() - x_res.cpu().numpy()) - (g[f"{tag}.x"] - o_res))      # means, on their masks
    d_scale = np.abs(s_hat.cpu().numpy() - o_s)
    n = x_q.numel()
    flips = int((x_q.cpu().numpy() != g[f"{tag}.x_q"]).sum())
    # the same module on stock torch ops of this GPU, in the two arithmetics the reference itself can run in (cuDNN TF32 is its
    # default on a GPU): how many final symbols THEY flip against the reference's CPU fp32 run
    other = {}
    for name, tf32 in (("cuDNN TF32", True), ("cuDNN fp32", False)):
        keep = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            tq = m._forward_torch(x, context, prev, False)[1]
        torch.backends.cudnn.allow_tf32 = keep
        other[name] = int((tq.cpu().numpy() != g[f"{tag}.x_q"]).sum())
    with capsys.disabled():
        print(f"\n[ctx4 {tag}] bf16 tensor-core path vs fp32 oracle: |d mean| max {d_mean.max():.4f} mean {d_mean.mean():.5f}; "
              f"|d scale| max {d_scale.max():.4f}; final symbols differing from the reference: {flips} of {n} "
              f"(stock torch on this GPU: " + ", ".join(f"{k} {v}" for k, v in other.items()) + ")")
    # bf16 operands through 22 layers with activations up to ~8; a flipped symbol of an early step changes x_hat_so_far by 1 there
    # and with it the later steps' parameters in its neighbourhood, so the bound is on the bulk, the maximum is only reported
    assert d_mean.mean() < 0.01 and np.percentile(d_mean, 99) < 0.06 and d_scale.mean() < 0.01 and np.percentile(d_scale, 99) < 0.06
    assert flips <= 0.02 * n
    # where the symbol agrees the reconstruction differs only by the mean error; x_hat = x_q + mean always
    assert np.abs((x_hat - x_q).cpu().numpy() - (x.cpu().numpy() - x_res.cpu().numpy())).max() < 1e-5
    # the module's own torch formula (autograd path) agrees with the kernels to the same tolerance
    with torch.enable_grad():
        xr = x.clone().requires_grad_(True)
        t_res, t_q, t_hat, t_s = m(xr, context=context, prev_subband=prev)
    assert (t_s.detach() - s_hat).abs().mean().item() < 0.01


def test_compress_decompress_round_trip(golden, conv_mode):
    """compress_staged -> native rANS -> decompress: the decoder re-derives every step's parameters from what it has decoded so far
    with the SAME kernels, so x_hat comes back bit-exactly; the staged int16 planes equal GaussianEncoder's own conversion."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    from learned_pmctf_b200.entropy_models.gaussian_model import CompressionModel
    g, w, m, x, context, prev = _load("a", golden)
    em = CompressionModel("laplace", ec_thread=False, stream_part=1)
    em.update()
    ge = em.gaussian_encoder
    with torch.no_grad():
        x_hat, staged = m.compress_staged(x, context=context, prev_subband=prev)
        outs = m.compress(x, context=context, prev_subband=prev)
    assert torch.equal(outs[8], x_hat)
    em.entropy_coder.reset()
    cdf, ln, off = ge.get_cdf_info()
    for k, (sym16, idx16) in enumerate(staged):
        want_idx = ge.build_indexes(outs[4 + k].cpu()).reshape(-1).to(torch.int16)       # CPU formula = true division, as the kernel
        d = (idx16.cpu().int() - want_idx.int()).abs()
        assert int(d.max()) <= 1 and float((d != 0).float().mean()) < 2e-3                  # an ulp of logf at an integer boundary
        assert torch.equal(sym16, outs[k].reshape(-1).to(torch.int16))
        em.entropy_coder.encoder.encode_with_indexes(sym16.cpu().numpy(), idx16.cpu().numpy(), cdf, ln, off)
    em.entropy_coder.flush()
    stream = em.entropy_coder.get_encoded_stream()
    em.entropy_coder.set_stream(stream)
    with torch.no_grad():
        back = m.decompress(ge, context=context, prev_subband=prev)
    assert torch.equal(back, x_hat)
    bits = 8 * len(stream) / x.numel()
    assert 0.5 < bits < 16


def test_four_step_1080p_subband_runs(conv_mode):
    """One level-0 subband of a 1080p luma plane (576 x 960) through the whole module: finite outputs, symbols consistent with the
    returned parameters (x_q == rint(x_res), x_hat == x_q + (x - x_res)) at the size the bench runs."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    m = ContextFusionFourStep(ctx_channels=2).to(DEV).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ctx_weights.make(5, 2).items()})
    g = torch.Generator().manual_seed(3)
    x = torch.round(torch.randn(1, 1, 576, 960, generator=g) * 4).to(DEV)
    c = torch.tanh(torch.randn(1, 1, 576, 960, generator=g)).to(DEV)
    p = torch.round(torch.randn(1, 1, 288, 480, generator=g) * 2).to(DEV)
    with torch.no_grad():
        x_res, x_q, x_hat, s_hat = m(x, context=c, prev_subband=p)
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    assert torch.isfinite(s_hat).all() and torch.isfinite(x_hat).all()
    assert torch.equal(x_q, torch.round(x_res))
    assert (x_hat - (x_q + (x - x_res))).abs().max().item() < 1e-5


def test_fused_head_equals_separate_projection(conv_mode):
    """pmctf_ctx_conv112_head == pmctf_ctx_conv112 (fp32 map) + pmctf_ctx_head, bit for bit."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    lib = nat.lib()
    g = torch.Generator().manual_seed(9)
    N, H, W = 2, 37, 61
    x = to_bf16_planar(torch.randn(N, 112, H, W, generator=g)).to(DEV)
    res = to_f32_planar(torch.randn(N, 112, H, W, generator=g)).to(DEV)
    w = (torch.randn(112, 112, 3, 3, generator=g) * 0.05).to(DEV)
    b = (torch.randn(112, generator=g) * 0.1).to(DEV)
    hw = (torch.randn(2, 112, 1, 1, generator=g) * 0.1).to(DEV)
    hb = torch.randn(2, generator=g).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    packed = torch.empty(int(lib.pmctf_ctx_packed_bytes(9)), dtype=torch.uint8, device=DEV)
    nat.check(lib.pmctf_ctx_pack_conv(w.data_ptr(), 9, packed.data_ptr(), st), "pack")
    of = torch.empty((N, 28, H, W, 4), device=DEV)
    s1, m1, s2, m2 = (torch.empty((N, 1, H, W), device=DEV) for _ in range(4))
    nat.check(lib.pmctf_ctx_conv112(x.data_ptr(), packed.data_ptr(), 9, b.data_ptr(), res.data_ptr(), None, 1.0, of.data_ptr(), None, N, H, W, st), "conv")
    nat.check(lib.pmctf_ctx_head(of.data_ptr(), hw.data_ptr(), hb.data_ptr(), s1.data_ptr(), m1.data_ptr(), N, H, W, st), "head")
    nat.check(lib.pmctf_ctx_conv112_head(x.data_ptr(), packed.data_ptr(), 9, b.data_ptr(), res.data_ptr(), None, 1.0, hw.data_ptr(), hb.data_ptr(),
                                         s2.data_ptr(), m2.data_ptr(), N, H, W, st), "conv_head")
    torch.cuda.synchronize()
    assert torch.equal(s1, s2) and torch.equal(m1, m2)
