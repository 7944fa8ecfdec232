"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs -- BIT-EXACT, floats included, because both sides implement the same arithmetic
contract -- and against the golden vectors produced by the unmodified reference.

Tolerances vs the reference goldens (0..255 scale) are those of tests/test_oracle_golden.py:
warp / chroma-MV bit-exact with the golden's linspace tables; conv-bearing outputs <= 2e-4 ..1e-3
(MKLDNN's summation order is unspecified); quantised symbols exact at the golden sizes.
"""
import numpy as np
import pytest
import torch

from conftest import sub_sd
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import learned_pmctf_b200 as pkg
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return pkg


@pytest.fixture(scope="module")
def model(P, weights):
    m = P.pMCTF(num_me_stages=4).cuda().eval()
    m.load_reference_state_dict({k: torch.from_numpy(v) for k, v in weights.items()} |
                                {k.replace("lift_h", "lift_v"): torch.from_numpy(v) for k, v in weights.items() if "lift_h" in k})
    return m


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


def rnd(shape, seed, lo=0.0, hi=255.0):
    g = np.random.default_rng(seed)
    return (lo + (hi - lo) * g.random(shape)).astype(np.float32)


def smooth_flow(n, h, w, seed, sigma=5.0):
    g = np.random.default_rng(seed)
    f = g.normal(0, sigma, (n, 2, h, w)).astype(np.float32)
    f[:, :, :2] -= 30
    f[:, :, :, -3:] += 41.5
    return f


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def assert_bitexact(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    if not np.array_equal(got, want):
        bad = got != want
        raise AssertionError(f"{what}: {int(bad.sum())}/{got.size} elements differ, max abs {maxabs(got, want):.3e}")


# ---- a1 / a2 ----------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,fn", [((1, 1, 40, 72), 1), ((2, 1, 33, 65), 2), ((4, 1, 18, 34), 2), ((1, 3, 16, 24), 1)])
def test_flow_warp_vs_oracle(P, shape, fn):
    im, fl = rnd(shape, 1), smooth_flow(fn, shape[2], shape[3], 2)
    for sign in (1.0, -1.0):
        got = npy(P.ops.flow_warp(cu(im), cu(fl), sign))
        if fn in (1, shape[0]):
            want = orc.flow_warp(im, fl, sign)
        else:  # planes share fields pairwise
            want = np.concatenate([orc.flow_warp(im[2 * i:2 * i + 2], fl[i:i + 1], sign) for i in range(fn)])
        assert_bitexact(got, want, f"flow_warp {shape} sign {sign}")


@pytest.mark.parametrize("shape,fn", [((2, 1, 45, 200), 2), ((1, 1, 96, 333), 1), ((4, 1, 32, 128), 2), ((1, 1, 130, 129), 1)])
def test_flow_warp_wide_planes_vs_oracle(P, shape, fn):
    """single-channel planes wider than one 128-thread block and ragged in both directions, with vectors that differ from lane to lane
    (a warp's taps land in different rows), a band of long vectors, vectors out of the frame on all four sides (border clamp) and a row
    of zero vectors (-0.0 under sign = -1)"""
    n, _, h, w = shape
    g = np.random.default_rng(5)
    fl = smooth_flow(fn, h, w, 2)
    fl += g.normal(0, 3.0, fl.shape).astype(np.float32)                 # lanes of a warp land in different rows
    fl[:, :, : h // 3, w // 4: w // 2] *= 4.0                            # a band of long vectors
    fl[:, 0, :, :6] -= 70.0                                              # out of the frame: left / right / top / bottom
    fl[:, 0, :, -6:] += 70.0
    fl[:, 1, :5, :] -= 50.0
    fl[:, 1, -5:, :] += 50.0
    fl[:, :, h // 2, :] = 0.0                                            # zero vectors (and -0.0 under sign = -1)
    im = rnd(shape, 3)
    for sign in (1.0, -1.0):
        got = npy(P.ops.flow_warp(cu(im), cu(fl), sign))
        if fn in (1, n):
            want = orc.flow_warp(im, fl, sign)
        else:
            want = np.concatenate([orc.flow_warp(im[2 * i:2 * i + 2], fl[i:i + 1], sign) for i in range(fn)])
        assert_bitexact(got, want, f"flow_warp wide {shape} sign {sign}")


@pytest.mark.parametrize("tag", ["luma", "chromaN", "tile", "rgb"])
def test_flow_warp_vs_reference_golden(P, golden, tag):
    g = golden("warp")
    for sign, key in ((1.0, "out_pos"), (-1.0, "out_neg")):
        got = npy(P.ops.flow_warp(cu(g[f"{tag}.im"]), cu(g[f"{tag}.flow"]), sign, cu(g[f"{tag}.lin_x"]), cu(g[f"{tag}.lin_y"])))
        assert_bitexact(got, g[f"{tag}.{key}"], f"{tag}.{key}")
        got = npy(P.flow_warp(cu(g[f"{tag}.im"]), cu(g[f"{tag}.flow"]) * sign))  # public API, own linspace
        assert maxabs(got, g[f"{tag}.{key}"]) <= 2e-4 * 255


def test_chroma_mv_down(P, golden):
    g = golden("warp")
    assert_bitexact(npy(P.ops.chroma_mv_down(cu(g["down.mv"]))), g["down.out"], "chroma mv vs reference")
    mv = smooth_flow(3, 20, 36, 5)
    assert_bitexact(npy(P.ops.chroma_mv_down(cu(mv))), orc.chroma_mv_down(mv), "chroma mv vs oracle")
    assert_bitexact(npy(P.bilineardownsacling(cu(mv))), orc.chroma_mv_down(mv) * 2, "bilineardownsacling")


# ---- a3 / a4 ----------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 32, 32), (2, 1, 24, 40), (1, 1, 33, 31), (3, 1, 70, 45), (1, 1, 5, 7), (1, 1, 64, 96)])
def test_predict_update_vs_oracle(P, model, weights, shape):
    x = rnd(shape, 7)
    pu_t = orc.PU(sub_sd(weights, "temporal_filtering.0.P_t."))
    assert_bitexact(npy(model.temporal_filtering[0].P_t(cu(x))), orc.predict_update(x, pu_t), f"P_t {shape}")
    pu_s = orc.PU(sub_sd(weights, "hp_coder.wavelet_transform.lift_h.U_2."))
    got = npy(P.ops.predict_update(cu(x), model.hp_coder.wavelet_transform.lift_h.U_2.packed(), 1 / 256.0))
    assert_bitexact(got, orc.predict_update(x, pu_s, 1 / 256.0), f"U_2 {shape}")


def test_predict_update_vs_reference_golden(model, golden):
    g = golden("pu")
    x = cu(g["x"])
    assert maxabs(npy(model.temporal_filtering[0].P_t(x)), g["P_t0"]) <= 2e-5 * float(np.abs(g["P_t0"]).max())
    assert maxabs(npy(model.temporal_filtering[0].predict_filter(x)), g["predict0"]) <= 1e-4
    assert maxabs(npy(model.temporal_filtering[3].update_filter(x - 100.0)), g["update3"]) <= 1e-4


@pytest.mark.parametrize("stage", [0, 3])
def test_temporal_filters_vs_oracle(model, weights, stage):
    x = rnd((2, 1, 37, 50), 9)
    tl = model.temporal_filtering[stage]
    assert_bitexact(npy(tl.predict_filter(cu(x))),
                    orc.temporal_filter(x, orc.PU(sub_sd(weights, f"temporal_filtering.{stage}.P_t.")), orc.SCALE_P), "predict_filter")
    assert_bitexact(npy(tl.update_filter(cu(x))),
                    orc.temporal_filter(x, orc.PU(sub_sd(weights, f"temporal_filtering.{stage}.U_t.")), orc.SCALE_U), "update_filter")


# ---- a5 / a6 ----------------------------------------------------------------------------------
@pytest.mark.parametrize("stage,shape", [(0, (1, 1, 64, 96)), (3, (2, 1, 34, 46)), (1, (1, 1, 33, 70))])
def test_mctf_vs_oracle(model, weights, stage, shape):
    ref, cur = rnd(shape, 11), rnd(shape, 12)
    mv = smooth_flow(1, shape[2], shape[3], 13)
    Pt = orc.PU(sub_sd(weights, f"temporal_filtering.{stage}.P_t."))
    Ut = orc.PU(sub_sd(weights, f"temporal_filtering.{stage}.U_t."))
    got = model.forward_MCTF(cu(ref), cu(cur), cu(mv), stage_idx=stage)
    want = orc.forward_mctf(ref, cur, mv, Pt, Ut)
    for g_, w_, name in zip(got, want, ("L_t", "H_t", "pred", "inv")):
        assert_bitexact(npy(g_), w_, f"forward_MCTF {name}")
    r, c = model.inverse_MCTF(got[0], got[1], cu(mv), stage_idx=stage)
    wr, wc = orc.inverse_mctf(want[0], want[1], mv, Pt, Ut)
    assert_bitexact(npy(r), wr, "inverse_MCTF ref")
    assert_bitexact(npy(c), wc, "inverse_MCTF cur")
    assert maxabs(npy(r), ref) <= 2e-4 and maxabs(npy(c), cur) <= 2e-4  # perfect reconstruction


def test_mctf_chroma_downscale_vs_oracle(model, weights):
    H, W = 48, 80
    ref, cur = rnd((2, 1, H // 2, W // 2), 21), rnd((2, 1, H // 2, W // 2), 22)
    mv = smooth_flow(1, H, W, 23)
    Pt = orc.PU(sub_sd(weights, "temporal_filtering.2.P_t."))
    Ut = orc.PU(sub_sd(weights, "temporal_filtering.2.U_t."))
    mvc = orc.chroma_mv_down(mv)
    want = orc.forward_mctf(ref, cur, mvc, Pt, Ut)
    got = model.forward_MCTF(cu(ref), cu(cur), cu(mv), stage_idx=2, mv_down=True)          # fused 2x2-mean/2
    got2 = model.forward_MCTF(cu(ref), cu(cur), cu(mvc), stage_idx=2)                       # pre-scaled field
    for a, b, w_ in zip(got, got2, want):
        assert_bitexact(npy(a), w_, "chroma forward (fused mv_down)")
        assert_bitexact(npy(b), w_, "chroma forward (pre-scaled mv)")
    r, c = model.inverse_MCTF(got[0], got[1], cu(mv), downscale=True, stage_idx=2)
    wr, wc = orc.inverse_mctf(want[0], want[1], mv, Pt, Ut, downscale=True)
    assert_bitexact(npy(r), wr, "chroma inverse ref")
    assert_bitexact(npy(c), wc, "chroma inverse cur")


def test_mctf_batched_pairs_equal_single(model):
    """Batching independent pairs along N (n/mv_n planes share a field) must not change a bit."""
    H, W = 40, 56
    ref, cur = rnd((4, 1, H, W), 31), rnd((4, 1, H, W), 32)
    mv = smooth_flow(2, H, W, 33)
    got = model.forward_MCTF(cu(ref), cu(cur), cu(mv), stage_idx=1)
    for i in range(2):
        one = model.forward_MCTF(cu(ref[2 * i:2 * i + 2]), cu(cur[2 * i:2 * i + 2]), cu(mv[i:i + 1]), stage_idx=1)
        for a, b in zip(got, one):
            assert_bitexact(npy(a[2 * i:2 * i + 2]), npy(b), "batched vs single")


@pytest.mark.parametrize("s", [0, 3])
def test_mctf_vs_reference_golden(model, golden, s):
    g = golden("mctf")
    L, H, pred, inv = model.forward_MCTF(cu(g["ref"]), cu(g["cur"]), cu(g["mv"]), stage_idx=s)
    for t, k in ((L, "L"), (H, "H"), (pred, "pred"), (inv, "inv")):
        assert maxabs(npy(t), g[f"s{s}.{k}"]) <= 2e-4 * 255, k
    r, c = model.inverse_MCTF(cu(g[f"s{s}.L"]), cu(g[f"s{s}.H"]), cu(g["mv"]), stage_idx=s)
    assert maxabs(npy(r), g[f"s{s}.ref_rec"]) <= 2e-4 * 255 and maxabs(npy(c), g[f"s{s}.cur_rec"]) <= 2e-4 * 255
    Lc, Hc, _, _ = model.forward_MCTF(cu(g["ref_c"]), cu(g["cur_c"]), cu(g["mv"]), stage_idx=s, mv_down=True)
    assert maxabs(npy(Lc), g[f"s{s}.Lc"]) <= 2e-4 * 255 and maxabs(npy(Hc), g[f"s{s}.Hc"]) <= 2e-4 * 255
    rc, cc = model.inverse_MCTF(cu(g[f"s{s}.Lc"]), cu(g[f"s{s}.Hc"]), cu(g["mv"]), downscale=True, stage_idx=s)
    assert maxabs(npy(rc), g[f"s{s}.ref_c_rec"]) <= 2e-4 * 255 and maxabs(npy(cc), g[f"s{s}.cur_c_rec"]) <= 2e-4 * 255


# ---- a7 - a10 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 1, 32, 48), (1, 1, 4, 6), (1, 1, 66, 35), (3, 1, 10, 130)])
def test_iwave1d_vs_oracle(model, weights, shape):
    w = orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h."))
    lift = model.lp_coder.wavelet_transform.lift_h
    x = rnd(shape, 41, -120, 130)
    l, h = lift.forward_lift(cu(x))
    wl, wh = orc.iwave1d_forward(x, w)
    assert_bitexact(npy(l), wl, "forward_lift l")
    assert_bitexact(npy(h), wh, "forward_lift h")
    assert_bitexact(npy(lift.backward_lift(l, h)), orc.iwave1d_backward(wl, wh, w), "backward_lift")
    # transposed view in, transposed views out (the column pass of wavelet_transform.py:32-40)
    xt = np.ascontiguousarray(x.transpose(0, 1, 3, 2))
    if xt.shape[2] % 2 == 0 and xt.shape[2] >= 4:
        lt, ht = lift.forward_lift(cu(x).permute(0, 1, 3, 2))
        wlt, wht = orc.iwave1d_forward(xt, w)
        assert_bitexact(npy(lt), wlt, "forward_lift on a permuted view (l)")
        assert_bitexact(npy(ht), wht, "forward_lift on a permuted view (h)")
        assert_bitexact(npy(lift.backward_lift(lt, ht)), orc.iwave1d_backward(wlt, wht, w), "backward_lift on permuted views")


@pytest.mark.parametrize("shape", [(2, 1, 32, 48), (1, 1, 4, 4), (1, 1, 70, 36), (1, 1, 64, 96)])
def test_lift2d_vs_oracle(model, weights, shape):
    w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    lift = model.hp_coder.wavelet_transform
    x = rnd(shape, 43, -100, 150)
    d = lift.forward_lift_2d(cu(x))
    wd = orc.lift2d_forward(x, w)
    for k in ("ll", "lh", "hl", "hh"):
        assert_bitexact(npy(d[k]), wd[k], f"forward_lift_2d {k}")
    assert_bitexact(npy(d["l"].permute(0, 1, 3, 2)), wd["l"], "row-pass l")
    assert_bitexact(npy(d["h"].permute(0, 1, 3, 2)), wd["h"], "row-pass h")
    assert_bitexact(npy(lift.backward_lift_2d(d)), orc.lift2d_backward(wd, w), "backward_lift_2d")


def test_pwave_encode_decode(model, weights, golden):
    g = golden("pwave")
    w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    coder = model.hp_coder
    y = coder.encode(cu(g["x"]))
    wy = orc.pwave_encode(g["x"], w)
    for lvl in range(4):
        for k in ("ll", "lh", "hl", "hh"):
            assert_bitexact(npy(y[lvl][k]), wy[lvl][k], f"encode {lvl}.{k} vs oracle")
            assert maxabs(npy(y[lvl][k]), g[f"enc.{lvl}.{k}"]) <= 1e-3, f"encode {lvl}.{k} vs reference"
    dec = coder.decode({lvl: dict(y[lvl]) for lvl in range(4)})
    assert_bitexact(npy(dec), orc.pwave_decode({lvl: dict(wy[lvl]) for lvl in range(4)}, w), "decode vs oracle")
    assert maxabs(npy(dec), g["x"]) <= 1e-3  # perfect reconstruction (reference: <= 2.5e-4 .. 1e-3)


# ---- a11 / a12 --------------------------------------------------------------------------------
@pytest.mark.parametrize("qi", [0, 4, 8, 12, 16, 20])
@pytest.mark.parametrize("tag", ["lp", "hp1"])
def test_symbols_bitexact_vs_reference(model, weights, golden, qi, tag):
    g = golden("pwave")
    w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    p = f"q{qi}.{tag}."
    q, qll = float(g[p + "q"].reshape(())), float(g[p + "qll"].reshape(()))
    x_hat, hat = model.hp_coder.spatial_wavelet_dec(cu(g["x"]), q, qll, return_symbols=True)
    ox, ohat = orc.spatial_wavelet_dec(g["x"], w, q, qll)
    assert_bitexact(npy(x_hat), ox, "x_hat vs oracle")
    for lvl in hat:
        for b, v in hat[lvl].items():
            assert_bitexact(npy(v), ohat[lvl][b], f"symbols {lvl}.{b} vs oracle")
            assert_bitexact(npy(v).astype(np.int16), g[p + f"sym.{lvl}.{b}"], f"symbols {lvl}.{b} vs REFERENCE (q_index {qi}, {tag})")
    assert maxabs(npy(x_hat), g[p + "x_hat"]) <= 1e-3 * 255


def test_q_from_parameters_matches_reference(model, golden):
    g = golden("pwave")
    for qi in (0, 4, 8, 12, 16, 20):
        q, qll = model.hp_coder.q_pair(qi, model.hp_qp_scale(1, qi))
        assert abs(float(q) - float(g[f"q{qi}.hp1.q"].reshape(()))) <= 2e-7 * float(q)
        assert abs(float(qll) - float(g[f"q{qi}.hp1.qll"].reshape(()))) <= 2e-7 * float(qll)


def test_quantize_dequantize_vs_oracle(P):
    s = rnd((1, 1, 37, 53), 51, -40000, 40000)
    s.ravel()[:8] = [0.5, 1.5, 2.5, -0.5, -1.5, 8191.5, -8192.5, 3.4999998]
    for q in (1.0, 0.0625, 0.37):
        for do_round in (True, False):
            assert_bitexact(npy(P.ops.quantize(cu(s), q, 8192.0, True, do_round)), orc.quantize(s, q, 8192.0, True, do_round), "quantize")
        assert_bitexact(npy(P.ops.dequantize(cu(s), q)), orc.dequantize(s, q), "dequantize")
    assert npy(P.ops.quantize(cu(np.zeros((0,), np.float32)), 1.0)).size == 0  # empty input
    # the 128-bit fast paths need 16-byte aligned pointers: odd offsets and odd sizes take the scalar path / the tail loop
    big = cu(rnd((1, 1, 41, 67), 52, -9000, 9000))
    for off, n in ((0, 2747), (1, 2744), (3, 2001), (4, 5), (2, 3)):
        v = big.view(-1)[off:off + n]
        assert_bitexact(npy(P.ops.quantize(v, 0.37)), orc.quantize(npy(v), 0.37), f"quantize off {off} n {n}")
        assert_bitexact(npy(P.ops.dequantize(v, 0.37)), orc.dequantize(npy(v), 0.37), f"dequantize off {off} n {n}")
    for shape, off in (((3, 1, 9, 7), 0), ((2, 1, 8, 8), 1), ((1, 1, 5, 3), 2)):   # plane sizes not divisible by 4, unaligned base
        n = int(np.prod(shape))
        v = big.view(-1)[off:off + n].view(shape)
        st = torch.zeros((shape[0], 2), dtype=torch.int64, device="cuda")
        want = orc.quantize(npy(v), 0.21)
        assert_bitexact(npy(P.ops.quantize_stats(v, 0.21, st)), want, f"quantize_stats {shape} off {off}")
        a = np.abs(want.reshape(shape[0], -1)).astype(np.int64)
        assert st.cpu().tolist() == np.stack([a.sum(1), (a != 0).sum(1)], 1).tolist()


@pytest.mark.parametrize("bitdepth,levels", [(8, 2), (10, 3), (8, 1)])
def test_pwave_other_depths_and_levels(P, bitdepth, levels):
    """Constructor arguments the reference exposes but its scripts never vary: decomp_levels (pWave.py:36,139-157) and bitdepth
    (dynamic_range = 2 ** bitdepth of the lifting steps, lifting_1d.py:62,108)."""
    torch.manual_seed(7)
    coder = P.pWave(bitdepth=bitdepth, decomp_levels=levels, lossy=True).cuda().eval()
    with torch.no_grad():
        for p in coder.wavelet_transform.parameters():
            if p.dim() == 4 and p.shape[-1] == 3:
                p.normal_(0, 0.08)
            elif p.dim() == 1:
                p.normal_(0, 0.05)
    sd = {k: v.detach().cpu().numpy() for k, v in coder.state_dict().items()}
    w = orc.IWave(sub_sd(sd, "wavelet_transform.lift_h."), dynamic_range=float(2 ** bitdepth))
    x = rnd((2, 1, 48, 80), 71, 0, 2 ** bitdepth - 1)
    y = orc.pwave_encode(x, w, levels=levels)
    enc = coder.encode_bands(cu(x))
    assert sorted(enc) == list(range(levels))
    for lvl in range(levels):
        for b in ("ll", "lh", "hl", "hh"):
            assert_bitexact(npy(enc[lvl][b]), y[lvl][b], f"bitdepth {bitdepth} levels {levels}: {lvl}.{b}")
    dec = coder.decode({lvl: dict(enc[lvl]) for lvl in range(levels)})
    assert_bitexact(npy(dec), orc.pwave_decode({lvl: dict(y[lvl]) for lvl in range(levels)}, w, levels=levels), "decode")


# ---- error behaviour ----------------------------------------------------------------------------
def test_errors(P, model):
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        P.flow_warp(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="does not match"):
        P.flow_warp(torch.zeros(1, 1, 8, 8).cuda(), torch.zeros(1, 2, 8, 9).cuda())
    with pytest.raises(RuntimeError, match="float32"):
        P.flow_warp(torch.zeros(1, 1, 8, 8).cuda().half(), torch.zeros(1, 2, 8, 8).cuda())
    with pytest.raises(RuntimeError):
        model.hp_coder.wavelet_transform.forward_lift_2d(torch.zeros(1, 1, 9, 8).cuda())  # odd height
    # the fused kernels have no backward: the raw ops refuse tensors that require grad, while the modules switch to the
    # differentiable training kernels (tests/test_gpu_train.py)
    xg = torch.zeros(1, 1, 8, 8, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError, match="backward"):
        with torch.enable_grad():
            P.ops.temporal_filter(xg, model.temporal_filtering[0].descriptor(), 0)
    with torch.enable_grad():
        assert model.temporal_filtering[0].predict_filter(xg).grad_fn is not None
    s = P._native.Step()
    assert P._native.lib().pmctf_lift_step(s, None) == -1  # EINVAL, nothing launched
    # tensor mode takes the small fp32 parameters from the host registry pmctf_pack_pu_weights() fills: a device-side COPY of a
    # packed block is not a registered block and is rejected instead of running with somebody else's weights
    if P.ops.get_conv_mode() == "tensor":
        pu = model.temporal_filtering[0].P_t
        x = torch.zeros(1, 1, 16, 32, device="cuda")
        P.ops.predict_update(x, pu.packed())
        with pytest.raises(RuntimeError, match="predict_update"):
            P.ops.predict_update(x, pu.packed().clone())


# ---- BASELINE.json full sizes: size-independent properties --------------------------------------
def test_full_size_1080p_properties(model):
    torch.manual_seed(5)
    H, W = 1152, 1920
    ref = (torch.rand(1, 1, H, W, device="cuda") * 255).round()
    cur = (ref.roll((2, -3), (2, 3)) + 4 * torch.randn(1, 1, H, W, device="cuda")).clamp(0, 255).round()
    mv = torch.nn.functional.avg_pool2d(torch.randn(1, 2, H, W, device="cuda") * 12, 9, 1, 4)
    L, Hh, _, _ = model.forward_MCTF(ref, cur, mv, stage_idx=0)
    r, c = model.inverse_MCTF(L, Hh, mv, stage_idx=0)
    assert float((r - ref).abs().max()) <= 3e-4 and float((c - cur).abs().max()) <= 3e-4
    # chroma planes + fused MV down-scaling
    refc, curc = ref[:, :, ::2, ::2].repeat(2, 1, 1, 1).contiguous(), cur[:, :, ::2, ::2].repeat(2, 1, 1, 1).contiguous()
    Lc, Hc, _, _ = model.forward_MCTF(refc, curc, mv, stage_idx=0, mv_down=True)
    rc, cc = model.inverse_MCTF(Lc, Hc, mv, downscale=True, stage_idx=0)
    assert float((rc - refc).abs().max()) <= 3e-4 and float((cc - curc).abs().max()) <= 3e-4
    # spatial transform: perfect reconstruction and linear-ish energy sanity at 1080p
    y = model.hp_coder.encode_bands(Hh)
    assert y[3]["ll"].shape == (1, 1, 72, 120)
    dec = model.hp_coder.decode({lvl: dict(y[lvl]) for lvl in range(4)})
    assert float((dec - Hh).abs().max()) <= 2e-3
    # quantise -> dequantise -> decode stays within half a step of the unquantised path per band
    x_hat, hat = model.hp_coder.spatial_wavelet_dec(Hh, 0.25, 0.5, return_symbols=True)
    for lvl in hat:
        for b, v in hat[lvl].items():
            assert bool((v == v.round()).all()) and float(v.abs().max()) <= 8192
            q = 0.5 if b == "ll" else 0.25
            assert float((v - (y[lvl][b] * q).clamp(-8192, 8192)).abs().max()) <= 0.5 + 1e-3
    assert torch.isfinite(x_hat).all()


# ---- randomised shape sweep: tile-boundary and ragged cases of the fused step (both conv modes) ------------------
@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_vs_oracle(P, model, weights, seed):
    g = np.random.default_rng(100 + seed)
    n = int(g.integers(1, 4))
    h, w = int(g.integers(2, 75)), int(g.integers(2, 99))
    x = rnd((n, 1, h, w), seed, -300, 300)
    pu = orc.PU(sub_sd(weights, "temporal_filtering.1.U_t."))
    assert_bitexact(npy(model.temporal_filtering[1].U_t(cu(x))), orc.predict_update(x, pu), f"PU {x.shape}")
    if h >= 2 and w >= 2:
        mv = smooth_flow(1, h, w, seed)
        ref, cur = rnd((n, 1, h, w), seed + 50), rnd((n, 1, h, w), seed + 60)
        Pt = orc.PU(sub_sd(weights, "temporal_filtering.1.P_t."))
        got = model.forward_MCTF(cu(ref), cu(cur), cu(mv), stage_idx=1)
        want = orc.forward_mctf(ref, cur, mv, Pt, pu)
        for a, b, name in zip(got, want, ("L", "H", "pred", "inv")):
            assert_bitexact(npy(a), b, f"forward_MCTF {name} {ref.shape}")
    h2, w2 = 2 * int(g.integers(2, 40)), 2 * int(g.integers(2, 50))
    xs = rnd((n, 1, h2, w2), seed + 7, -120, 130)
    iw = orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h."))
    d = model.lp_coder.wavelet_transform.forward_lift_2d(cu(xs))
    wd = orc.lift2d_forward(xs, iw)
    for k in ("ll", "lh", "hl", "hh"):
        assert_bitexact(npy(d[k]), wd[k], f"lift2d {k} {xs.shape}")
    assert_bitexact(npy(model.lp_coder.wavelet_transform.backward_lift_2d(d)), orc.lift2d_backward(wd, iw), f"lift2d inverse {xs.shape}")
