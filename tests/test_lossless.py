"""The LOSSLESS variant of the path (pMCTF(lossy=False): rounded warp, rounded 0.1 * PU updates, no subband scaling --
lifting_1d.py:110-148, wavelet_transform_temporal_mctf.py:30-43, pMCTF_L.py:302-326).

  * CPU: the oracle against vectors of the unmodified reference (tests/golden/lossless.npz, oracle/make_golden.py lossless).
    Everything is integer valued, so agreement is EXACT except where fp32 round-off of a pre-round value sits on a rounding
    boundary; such flips (a +-1 on a sample) are counted and bounded.
  * GPU (-m gpu): the CUDA kernels against the oracle, bit-exact, and the domain's own size-independent property -- perfect
    reconstruction of integer frames through analysis + synthesis -- at 1080p.
"""
import numpy as np
import pytest
import torch

from conftest import sub_sd
from oracle import oracle as orc


@pytest.fixture(scope="module")
def g(golden):
    return golden("lossless")


@pytest.fixture(scope="module")
def w(g):
    return {k[2:]: g[k] for k in g.files if k.startswith("w.")}


def flips(a, b):
    """(#samples that differ, max abs difference)"""
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    return int(np.sum(d != 0)), float(d.max())


def pus(w, i):
    return orc.PU(sub_sd(w, f"temporal_filtering.{i}.P_t.")), orc.PU(sub_sd(w, f"temporal_filtering.{i}.U_t."))


@pytest.mark.parametrize("s", [0, 2])
def test_oracle_lossless_mctf_vs_reference(g, w, s):
    Pt, Ut = pus(w, s)
    L, Hh, pred, inv = orc.forward_mctf(g["ref"], g["cur"], g["mv"], Pt, Ut, lossy=False, lin_x=g["lin_x"], lin_y=g["lin_y"])
    for got, key in ((L, "L"), (Hh, "H"), (pred, "pred"), (inv, "inv")):
        assert np.all(got == np.rint(got)), "lossless outputs are integers"
        n, m = flips(got, g[f"s{s}.{key}"])
        assert n <= 4 and m <= 1.0, f"stage {s} {key}: {n} samples differ (max {m})"
    r, c = orc.inverse_mctf(g[f"s{s}.L"], g[f"s{s}.H"], g["mv"], Pt, Ut, lossy=False, lin_x=g["lin_x"], lin_y=g["lin_y"])
    assert flips(r, g[f"s{s}.ref_rec"])[0] <= 4 and flips(c, g[f"s{s}.cur_rec"])[0] <= 4
    # the oracle's own round trip is exact (integer lifting)
    r2, c2 = orc.inverse_mctf(L, Hh, g["mv"], Pt, Ut, lossy=False, lin_x=g["lin_x"], lin_y=g["lin_y"])
    assert np.array_equal(r2, g["ref"]) and np.array_equal(c2, g["cur"])
    # chroma planes with the down-scaled, tiled motion field
    Lc, Hc, _, _ = orc.forward_mctf(g["ref_c"], g["cur_c"], orc.chroma_mv_down(g["mv"]), Pt, Ut, lossy=False, lin_x=g["lin_xc"], lin_y=g["lin_yc"])
    assert flips(Lc, g[f"s{s}.Lc"])[0] <= 4 and flips(Hc, g[f"s{s}.Hc"])[0] <= 4


def test_oracle_lossless_spatial_vs_reference(g, w):
    iw = orc.IWave(sub_sd(w, "lp_coder.wavelet_transform.lift_h."), lossy=False)
    l, h = orc.iwave1d_forward(g["x1d"], iw)
    assert flips(l, g["l1d"])[0] <= 4 and flips(h, g["h1d"])[0] <= 4 and np.all(l == np.rint(l)) and np.all(h == np.rint(h))
    assert np.array_equal(orc.iwave1d_backward(l, h, iw), g["x1d"])
    y = orc.pwave_encode(g["x"], iw)
    total = bad = 0
    for lvl in range(4):
        for b in ("ll", "lh", "hl", "hh"):
            n, m = flips(y[lvl][b], g[f"enc.{lvl}.{b}"])
            assert m <= 2.0, (lvl, b, m)
            total, bad = total + y[lvl][b].size, bad + n
    assert bad <= max(4, total // 500), f"{bad} of {total} lossless coefficients differ from the reference"
    assert np.array_equal(orc.pwave_decode({lvl: dict(y[lvl]) for lvl in range(4)}, iw), g["x"])   # perfect reconstruction
    assert np.array_equal(g["dec"], g["x"])                                                          # ... as in the reference


# ---- CUDA ------------------------------------------------------------------------------------------------------------
def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def model(w):
    import learned_pmctf_b200 as P
    m = P.pMCTF(lossy=False, num_me_stages=4).cuda().eval()
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    m.load_reference_state_dict(sd | {k.replace("lift_h", "lift_v"): v for k, v in sd.items() if "lift_h" in k})
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("s", [0, 2])
def test_cuda_lossless_mctf_vs_oracle(g, w, model, s):
    Pt, Ut = pus(w, s)
    oL, oH, _, _ = orc.forward_mctf(g["ref"], g["cur"], g["mv"], Pt, Ut, lossy=False)
    L, Hh, _, _ = model.forward_MCTF(cu(g["ref"]), cu(g["cur"]), cu(g["mv"]), stage_idx=s)
    assert np.array_equal(L.cpu().numpy(), oL) and np.array_equal(Hh.cpu().numpy(), oH)
    r, c = model.inverse_MCTF(L, Hh, cu(g["mv"]), stage_idx=s)
    assert np.array_equal(r.cpu().numpy(), g["ref"]) and np.array_equal(c.cpu().numpy(), g["cur"])
    # chroma: fused motion-vector down-scaling
    oLc, oHc, _, _ = orc.forward_mctf(g["ref_c"], g["cur_c"], orc.chroma_mv_down(g["mv"]), Pt, Ut, lossy=False)
    Lc, Hc, _, _ = model.forward_MCTF(cu(g["ref_c"]), cu(g["cur_c"]), cu(g["mv"]), stage_idx=s, mv_down=True)
    assert np.array_equal(Lc.cpu().numpy(), oLc) and np.array_equal(Hc.cpu().numpy(), oHc)


@pytest.mark.gpu
def test_cuda_lossless_spatial_vs_oracle(g, w, model):
    iw = orc.IWave(sub_sd(w, "lp_coder.wavelet_transform.lift_h."), lossy=False)
    y = orc.pwave_encode(g["x"], iw)
    enc = model.lp_coder.encode_bands(cu(g["x"]))
    for lvl in range(4):
        for b in ("ll", "lh", "hl", "hh"):
            assert np.array_equal(enc[lvl][b].cpu().numpy(), y[lvl][b]), (lvl, b)
    dec = model.lp_coder.decode({lvl: dict(enc[lvl]) for lvl in range(4)})
    assert np.array_equal(dec.cpu().numpy(), g["x"])


@pytest.mark.gpu
def test_cuda_lossless_perfect_reconstruction_1080p(model):
    """Size-independent property at BASELINE's full size: integer 1080p frames survive temporal analysis -> 4-level spatial
    analysis -> spatial synthesis -> temporal synthesis exactly."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    ref = torch.randint(0, 256, (1, 1, 1152, 1920), device="cuda", generator=gen).float()
    cur = (ref.roll((2, -3), (2, 3)) + torch.randint(-4, 5, ref.shape, device="cuda", generator=gen)).clamp(0, 255)
    mv = torch.nn.functional.avg_pool2d(torch.randn(1, 2, 1152, 1920, device="cuda", generator=gen) * 12, 9, 1, 4)
    L, Hh, _, _ = model.forward_MCTF(ref, cur, mv, stage_idx=1)
    assert bool((L == L.round()).all()) and bool((Hh == Hh.round()).all())
    for coder, x in ((model.lp_coder, L), (model.hp_coder, Hh)):
        y = coder.encode_bands(x)
        assert all(bool((v == v.round()).all()) for lvl in y for v in y[lvl].values())
        assert torch.equal(coder.decode({lvl: dict(y[lvl]) for lvl in range(4)}), x)
    r, c = model.inverse_MCTF(L, Hh, mv, stage_idx=1)
    assert torch.equal(r, ref) and torch.equal(c, cur)


@pytest.mark.gpu
def test_cuda_lossless_gop_is_exact(model):
    """A whole GOP through the batched pipeline in lossless mode: no quantisation step is applied (pWave.py:184-202 with
    lossy=False), the symbols ARE the integer subbands, and the reconstruction equals the input frames exactly."""
    import learned_pmctf_b200 as P
    from learned_pmctf_b200 import gop as Gm
    gop, h0, w0 = 8, 128, 192
    gen = torch.Generator(device="cuda").manual_seed(9)
    y_u8 = torch.randint(0, 256, (gop, h0, w0), device="cuda", generator=gen, dtype=torch.uint8)
    c_u8 = torch.randint(0, 256, (gop, 2, h0 // 2, w0 // 2), device="cuda", generator=gen, dtype=torch.uint8)
    Y = P.ops.unpack_u8(y_u8, h0, w0)
    C = P.ops.unpack_u8(c_u8.view(-1, h0 // 2, w0 // 2), h0 // 2, w0 // 2).view(gop, 2, 1, h0 // 2, w0 // 2)
    mvs, n = [], gop
    while n > 1:
        n //= 2
        mvs.append(torch.randn(n, 2, h0, w0, device="cuda", generator=gen) * 2.5)
    codec = Gm.GopCodec(model, gop, q_index=12)
    rec_y, rec_c, st = codec.code_gop(Y, C, mvs, y_u8, c_u8)
    assert torch.equal(rec_y, Y) and torch.equal(rec_c, C)
    assert float(st[:, 3:6].abs().max()) == 0.0          # zero squared error on every plane
    s = P.ops.quantize(Y, 0.37, lossy=False)              # lossless quantise / dequantise leave integers alone
    assert torch.equal(s, Y) and torch.equal(P.ops.dequantize(s, 0.37, lossy=False), Y)
