"""GPU parity of the batched GOP pipeline and the host-format kernels against the oracle's pair-by-pair
restatement of the reference loop (test_pMCTF_flex.py:131-310).  Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import sub_sd
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import learned_pmctf_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


@pytest.fixture(scope="module")
def model(P, weights):
    m = P.pMCTF(num_me_stages=4).cuda().eval()
    m.load_reference_state_dict({k: torch.from_numpy(v) for k, v in weights.items()} |
                                {k.replace("lift_h", "lift_v"): torch.from_numpy(v) for k, v in weights.items() if "lift_h" in k})
    with torch.no_grad():  # distinct steps per temporal level / q_index (random init leaves them degenerate)
        for c in (m.lp_coder, m.hp_coder):
            c.QP.copy_(torch.tensor([1 / 32, 1 / 2]).view(2, 1, 1, 1))
            c.QP_ll.copy_(torch.tensor([1 / 16, 1 / 2]).view(2, 1, 1, 1))  # LL x q_ll must stay below clip_value 8192
        for i, p in enumerate(m.hp_q_scale):
            p.copy_(torch.tensor([1.0, 0.7 - 0.1 * i]).view(2, 1, 1, 1))
    return m


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_unpack_u8_and_sse(P):
    g = np.random.default_rng(1)
    for (n, h0, w0, hp, wp) in ((3, 10, 12, 16, 16), (2, 30, 50, 32, 64), (1, 7, 9, 8, 12)):
        u = g.integers(0, 256, (n, h0, w0), dtype=np.uint8)
        got = P.ops.unpack_u8(cu(u), hp, wp)
        assert np.array_equal(got.cpu().numpy()[:, 0], orc.unpack_u8(u, hp, wp))
        rec = got + torch.randn_like(got) * 3
        sse = P.ops.frame_sse(rec, cu(u))
        assert sse.cpu().tolist() == orc.frame_sse(rec.cpu().numpy()[:, 0], u).tolist()
    with pytest.raises(RuntimeError):
        P.ops.unpack_u8(cu(np.zeros((1, 8, 8), np.uint8)), 8, 6)  # padded width must be a multiple of 4 and >= w0


def test_quantize_stats(P):
    g = np.random.default_rng(2)
    s = (g.normal(0, 300, (5, 1, 18, 22))).astype(np.float32)
    s[0, 0, 0, :4] = [0.5, 1.5, -2.5, 1e6]
    st = torch.zeros((5, 2), dtype=torch.int64, device="cuda")
    out = P.ops.quantize_stats(cu(s), 0.37, st)
    want = orc.quantize(s, 0.37)
    assert np.array_equal(out.cpu().numpy(), want)
    a = np.abs(want.reshape(5, -1)).astype(np.int64)
    assert st.cpu().tolist() == np.stack([a.sum(1), (a != 0).sum(1)], 1).tolist()


def _inputs(G, H, W, seed):
    g = np.random.default_rng(seed)
    base = g.random((H + 64, W + 64)) * 255
    k = np.ones((5, 5)) / 25
    from numpy.lib.stride_tricks import sliding_window_view
    base = (sliding_window_view(np.pad(base, 2, mode="edge"), (5, 5)) * k).sum((-1, -2))
    y = np.stack([np.clip(base[8 + f:8 + f + H, 8 + 2 * f:8 + 2 * f + W] + g.normal(0, 2, (H, W)), 0, 255) for f in range(G)])
    y = np.rint(y).astype(np.uint8)
    c = np.rint(np.clip(g.random((G, 2, H // 2, W // 2)) * 64 + 96, 0, 255)).astype(np.uint8)
    mvs, n = [], G
    while n > 1:
        n //= 2
        mvs.append((g.normal(0, 2.5, (n, 2, H, W))).astype(np.float32))
    return y, c, mvs


@pytest.mark.parametrize("gop,h0,w0", [(4, 120, 190), (2, 128, 128), (8, 64, 64)])
def test_code_gop_vs_oracle(P, model, weights, gop, h0, w0):
    from learned_pmctf_b200 import gop as Gm
    q_index = 12
    codec = Gm.GopCodec(model, gop, q_index=q_index)
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    y, c, mvs = _inputs(gop, h0, w0, 7)
    mvs = [np.ascontiguousarray(np.pad(m, ((0, 0), (0, 0), (0, hp - h0), (0, wp - w0)))) for m in mvs]
    yd, cd = cu(y), cu(c)
    Y = P.ops.unpack_u8(yd, hp, wp)
    C = P.ops.unpack_u8(cd.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(gop, 2, 1, hp // 2, wp // 2)
    rec_y, rec_c, st = codec.code_gop(Y, C, [cu(m) for m in mvs], yd, cd)
    # oracle: pair by pair in the reference's order
    temporal = [(orc.PU(sub_sd(weights, f"temporal_filtering.{i}.P_t.")), orc.PU(sub_sd(weights, f"temporal_filtering.{i}.U_t.")))
                for i in range(4)]
    hp_w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    lp_w = orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h."))
    S = Gm.num_stages(gop)
    q_hp = [codec.q_pair("hp", s) for s in range(S)]
    assert len(set(q_hp)) == min(S, 4), "temporal-layer-adaptive steps must differ per stage"
    oy, oc, osym = orc.code_gop(orc.unpack_u8(y, hp, wp)[:, None], orc.unpack_u8(c, hp // 2, wp // 2)[:, :, None], mvs, temporal,
                                hp_w, lp_w, q_hp, codec.q_pair("lp", 0))
    assert np.array_equal(rec_y.cpu().numpy(), oy), f"luma: max diff {np.abs(rec_y.cpu().numpy() - oy).max()}"
    assert np.array_equal(rec_c.cpu().numpy(), oc)
    st = st.cpu().numpy()
    assert np.array_equal(st[:, 1:3].astype(np.int64), osym)
    assert np.array_equal(st[:, 3].astype(np.int64), orc.frame_sse(oy[:, 0], y))
    assert np.array_equal(st[:, 4:6].astype(np.int64), orc.frame_sse(oc[:, :, 0], c))
    assert st[0, 0] == 0 and np.all(st[1:, 0] == 1) and np.all(st[:, 7] == h0 * w0)
    px = h0 * w0
    mse = np.stack([st[:, 3] / px, st[:, 4] / (px // 4), st[:, 5] / (px // 4)], 1)
    psnr = 10 * np.log10(255.0 ** 2 / mse)
    assert np.allclose(st[:, 6], (6 * psnr[:, 0] + psnr[:, 1] + psnr[:, 2]) / 8, rtol=1e-12)
    # the host-buffer entry point gives the same statistics
    host = codec.code_sequence_host(torch.from_numpy(y).pin_memory(), torch.from_numpy(c).pin_memory(),
                                    [[torch.from_numpy(m).pin_memory() for m in mvs]])
    assert np.array_equal(host.numpy(), st)


def test_gop16_1080p_properties(P, model):
    """BASELINE.json full size (C3): one 1080p GOP-16, PROPERTY checks only (the parity test proper is
    test_config2_gop16_1080p_vs_oracle below).  Size-independent properties: with a very fine quantiser the
    whole analysis -> code -> synthesis chain is near-lossless, statistics are consistent, and coarser steps
    give fewer symbols and more distortion."""
    from learned_pmctf_b200 import gop as Gm
    y, c = Gm.synthetic_sequence(0, 16, 1080, 1920, "cuda")
    Y = P.ops.unpack_u8(y, 1152, 1920)
    C = P.ops.unpack_u8(c.view(-1, 540, 960), 576, 960).view(16, 2, 1, 576, 960)
    mvs = Gm.synthetic_motion(0, 0, 16, 1152, 1920, "cuda")
    assert [m.shape[0] for m in mvs] == [8, 4, 2, 1]
    res = {}
    for qi in (0, 20):
        codec = Gm.GopCodec(model, 16, q_index=qi)
        ry, rc, st = codec.code_gop(Y, C, mvs, y, c)
        assert ry.shape == Y.shape and rc.shape == C.shape and torch.isfinite(ry).all() and torch.isfinite(rc).all()
        res[qi] = st.cpu().numpy()
    assert res[20][:, 2].sum() > res[0][:, 2].sum()          # finer step -> more nonzero symbols
    assert res[20][:, 3].sum() < res[0][:, 3].sum()          # ... and less distortion
    assert np.all(res[20][:, 6] > res[0][:, 6])
    # transform chain alone (no quantiser) reconstructs: analysis -> synthesis
    codec = Gm.GopCodec(model, 16, q_index=20)
    Ly, Lc, Hs = codec.analysis(Y, C, mvs)
    ry, rc = codec.synthesis(Ly, Lc, Hs, mvs)
    assert float((ry - Y).abs().max()) <= 2e-3 and float((rc - C).abs().max()) <= 2e-3


def test_quantize_code_per_plane_steps(P):
    """pmctf_quantize_code: per-plane steps, int16 symbols at a column offset of a wider row, statistics, fused dequantise."""
    g = np.random.default_rng(5)
    s = g.normal(0, 400, (5, 1, 18, 22)).astype(np.float32)
    s[0, 0, 0, :4] = [0.5, 1.5, -2.5, 1e6]
    qs = np.array([0.37, 0.05, 0.5, 0.11, 0.25], np.float32)
    st = torch.zeros((5, 2), dtype=torch.int64, device="cuda")
    sym16 = torch.full((5, 1000), -7, dtype=torch.int16, device="cuda")
    out = P.ops.quantize_code(cu(s), cu(qs), st, sym16=sym16, sym16_offset=100)
    raw = P.ops.quantize_code(cu(s), cu(qs), None, dequant=False)
    for p in range(5):
        want = orc.quantize(s[p], float(qs[p]))
        assert np.array_equal(raw[p].cpu().numpy(), want)
        assert np.array_equal(out[p].cpu().numpy(), orc.dequantize(want, float(qs[p])))
        assert np.array_equal(sym16[p, 100:100 + 396].cpu().numpy(), want.reshape(-1).astype(np.int16))
        a = np.abs(want.reshape(-1)).astype(np.int64)
        assert st[p].cpu().tolist() == [int(a.sum()), int((a != 0).sum())]
    assert bool((sym16[:, :100] == -7).all()) and bool((sym16[:, 496:] == -7).all())
    with pytest.raises(RuntimeError):
        P.ops.quantize_code(cu(s), cu(qs[:3]), st)


def test_host_symbols_equal_oracle(P, model, weights):
    """The path's product on the HOST: code_sequence_host(return_symbols=True) brings the int16 symbols of every coded plane
    (entropy_models.py:37-40 hands exactly these to rANS) to pinned host memory; they equal the oracle's symbols, frame by
    frame, band by band, for both GOPs of a 2-GOP sequence."""
    from learned_pmctf_b200 import gop as Gm
    gop, h0, w0, n_gops = 4, 120, 190, 2
    codec = Gm.GopCodec(model, gop, q_index=12)
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    ys, cs, mvl = [], [], []
    for g in range(n_gops):
        y, c, mvs = _inputs(gop, h0, w0, 41 + g)
        ys.append(y), cs.append(c)
        mvl.append([np.ascontiguousarray(np.pad(m, ((0, 0), (0, 0), (0, hp - h0), (0, wp - w0)))) for m in mvs])
    y, c = np.concatenate(ys), np.concatenate(cs)
    stats, sym = codec.code_sequence_host(torch.from_numpy(y).pin_memory(), torch.from_numpy(c).pin_memory(),
                                          [[torch.from_numpy(m).pin_memory() for m in g] for g in mvl], return_symbols=True)
    assert sym["y"].dtype == torch.int16 and sym["y"].shape == (n_gops * gop, hp * wp) and sym["c"].shape == (n_gops * gop, 2, hp * wp // 4)
    assert sym["y"].is_pinned() and sym["coding_order"] == [1, 3, 2, 0]
    temporal, hp_w, lp_w = _oracle_weights(weights)
    q_hp = [codec.q_pair("hp", s) for s in range(2)]
    lay_y, lay_c = model.hp_coder.band_layout(hp, wp), model.hp_coder.band_layout(hp // 2, wp // 2)
    for g in range(n_gops):
        trace = {}
        _, _, osym = orc.code_gop(orc.unpack_u8(ys[g], hp, wp)[:, None], orc.unpack_u8(cs[g], hp // 2, wp // 2)[:, :, None], mvl[g],
                                  temporal, hp_w, lp_w, q_hp, codec.q_pair("lp", 0), trace=trace)
        for row, frame in enumerate(sym["coding_order"]):
            hat_y, hat_c = trace["sym"][frame]
            for lvl, b, off, n in lay_y:
                assert np.array_equal(sym["y"][g * gop + row, off:off + n].numpy(), hat_y[lvl][b].reshape(-1).astype(np.int16)), (g, frame, lvl, b)
            for lvl, b, off, n in lay_c:
                for ch in range(2):
                    assert np.array_equal(sym["c"][g * gop + row, ch, off:off + n].numpy(), hat_c[lvl][b][ch].reshape(-1).astype(np.int16))
        assert np.array_equal(stats.numpy()[g * gop:(g + 1) * gop, 1:3].astype(np.int64), osym)
    assert sum(n for *_, n in lay_y) == hp * wp


def _adversarial_motion(mvs):
    """SURVEY.md section 8d (ii) on top of the N(0, 4^2) fields: saturated +-32 px blocks, and 64 px vectors pointing out of the
    frame along all four borders (border clamping of the warp, video_net.py:47-50)."""
    mvs = [m.clone() for m in mvs]
    f = mvs[0]
    f[0, 0, :, :96] = -64.0          # left border, pointing left
    f[0, 0, :, -96:] = 64.0          # right border, pointing right
    f[1, 1, :96, :] = -64.0          # top border, pointing up
    f[1, 1, -96:, :] = 64.0          # bottom border (inside the zero padding rows), pointing down
    f[2, :, 300:600, 500:900] = 32.0
    f[3, :, 300:600, 500:900] = -32.0
    mvs[-1][0, 0, 200:400, 100:1800] = 32.0       # coarsest stage (temporal module set 3)
    mvs[-1][0, 1, 700:1100, 100:1800] = -64.0
    return mvs


def test_config2_gop16_1080p_vs_oracle(P, model, weights, conv_mode):
    """configs[2] at its own size: ONE 1080p GOP-16 (1152x1920 luma, [16,2,1,576,960] chroma, 8/4/2/1 motion fields incl. the
    +-32 px and out-of-frame cases, all four temporal module sets, q_index 12) through GopCodec -- every temporal subband frame
    (H of each stage, final L), the reconstruction and the symbol / distortion statistics BIT-EXACT against the oracle's
    pair-by-pair restatement of test_pMCTF_flex.py:131-291 / pMCTF_L.py:297-330."""
    if conv_mode == "ffma":
        pytest.skip("full-size oracle run once (tensor mode = the default contract); ffma is covered at the smaller sizes")
    from learned_pmctf_b200 import gop as Gm
    G, h0, w0, hp, wp = 16, 1080, 1920, 1152, 1920
    y, c = Gm.synthetic_sequence(0, G, h0, w0, "cuda")
    mvs = _adversarial_motion(Gm.synthetic_motion(0, 0, G, hp, wp, "cuda"))
    assert [m.shape[0] for m in mvs] == [8, 4, 2, 1] and float(mvs[0].abs().max()) == 64.0
    Y = P.ops.unpack_u8(y, hp, wp)
    C = P.ops.unpack_u8(c.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(G, 2, 1, hp // 2, wp // 2)
    codec = Gm.GopCodec(model, G, q_index=12)
    assert model.num_me_stages == 4 and codec.stages == 4
    Ly, Lc, Hs = codec.analysis(Y, C, mvs)                       # the temporal subbands as coded
    rec_y, rec_c, st = codec.code_gop(Y, C, mvs, y, c)
    torch.cuda.synchronize()
    temporal, hp_w, lp_w = _oracle_weights(weights)
    q_hp = [codec.q_pair("hp", s) for s in range(4)]
    assert len(set(q_hp)) == 4
    trace = {}
    y_np, c_np = y.cpu().numpy(), c.cpu().numpy()
    oy, oc, osym = orc.code_gop(orc.unpack_u8(y_np, hp, wp)[:, None], orc.unpack_u8(c_np, hp // 2, wp // 2)[:, :, None],
                                [m.cpu().numpy() for m in mvs], temporal, hp_w, lp_w, q_hp, codec.q_pair("lp", 0), trace=trace)
    # temporal analysis: every H frame of every stage, and the final low-pass frame
    for s, pairs in enumerate(Gm.dyadic_schedule(G)):
        Hy, Hc = Hs[s]
        for g, (_, cur) in enumerate(pairs):
            assert np.array_equal(Hy[g:g + 1].cpu().numpy(), trace["H"][cur][0]), f"H luma, stage {s} pair {g}"
            assert np.array_equal(Hc[g].cpu().numpy(), trace["H"][cur][1]), f"H chroma, stage {s} pair {g}"
    assert np.array_equal(Ly.cpu().numpy(), trace["L"][0]) and np.array_equal(Lc[0].cpu().numpy(), trace["L"][1])
    # reconstruction + statistics
    assert np.array_equal(rec_y.cpu().numpy(), oy), f"luma: max diff {np.abs(rec_y.cpu().numpy() - oy).max()}"
    assert np.array_equal(rec_c.cpu().numpy(), oc), f"chroma: max diff {np.abs(rec_c.cpu().numpy() - oc).max()}"
    st = st.cpu().numpy()
    assert np.array_equal(st[:, 1:3].astype(np.int64), osym)
    assert np.array_equal(st[:, 3].astype(np.int64), orc.frame_sse(oy[:, 0], y_np))
    assert np.array_equal(st[:, 4:6].astype(np.int64), orc.frame_sse(oc[:, :, 0], c_np))


def test_mctf_1080p_standalone_vs_oracle(P, model, weights, conv_mode):
    """forward_MCTF / inverse_MCTF on one 1152x1920 luma pair and one [2,1,576,960] chroma pair (mv_down: the fused 2x2-mean/2 of
    the luma field) with out-of-frame vectors, incl. the pred / inv outputs -- bit-exact vs the oracle (pMCTF_L.py:297-330)."""
    if conv_mode == "ffma":
        pytest.skip("full-size oracle run once (tensor mode)")
    from learned_pmctf_b200 import gop as Gm
    hp, wp = 1152, 1920
    y, c = Gm.synthetic_sequence(1, 2, 1080, 1920, "cuda")
    Y = P.ops.unpack_u8(y, hp, wp)
    C = P.ops.unpack_u8(c.view(-1, 540, 960), hp // 2, wp // 2).view(2, 2, 1, hp // 2, wp // 2)
    mv = _adversarial_motion(Gm.synthetic_motion(1, 0, 16, hp, wp, "cuda"))[0][:1].contiguous()
    Pt, Ut = _oracle_weights(weights)[0][2]
    L, H, pred, inv = model.forward_MCTF(Y[0:1], Y[1:2], mv, stage_idx=2)
    oL, oH, opred, oinv = orc.forward_mctf(Y[0:1].cpu().numpy(), Y[1:2].cpu().numpy(), mv.cpu().numpy(), Pt, Ut)
    for a, b, n in ((L, oL, "L"), (H, oH, "H"), (pred, opred, "pred"), (inv, oinv, "inv")):
        assert np.array_equal(a.cpu().numpy(), b), n
    r, cu_ = model.inverse_MCTF(L, H, mv, stage_idx=2)
    orr, occ = orc.inverse_mctf(oL, oH, mv.cpu().numpy(), Pt, Ut)
    assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(cu_.cpu().numpy(), occ)
    Lc, Hc, _, _ = model.forward_MCTF(C[0], C[1], mv, stage_idx=2, mv_down=True)
    omv = orc.chroma_mv_down(mv.cpu().numpy())
    oLc, oHc, _, _ = orc.forward_mctf(C[0].cpu().numpy(), C[1].cpu().numpy(), omv, Pt, Ut)
    assert np.array_equal(Lc.cpu().numpy(), oLc) and np.array_equal(Hc.cpu().numpy(), oHc)
    rc, cc = model.inverse_MCTF(Lc, Hc, mv, downscale=True, stage_idx=2)
    orc_, occ_ = orc.inverse_mctf(oLc, oHc, mv.cpu().numpy(), Pt, Ut, downscale=True)
    assert np.array_equal(rc.cpu().numpy(), orc_) and np.array_equal(cc.cpu().numpy(), occ_)


# ---- BASELINE.json configs as parity cases at their own sizes ---------------------------------------------------
def _oracle_weights(weights):
    temporal = [(orc.PU(sub_sd(weights, f"temporal_filtering.{i}.P_t.")), orc.PU(sub_sd(weights, f"temporal_filtering.{i}.U_t.")))
                for i in range(4)]
    return temporal, orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h.")), orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h."))


def test_config0_gop2_256x448_vs_oracle(P, model, weights):
    """configs[0]: pMCTF-L GOP-2 forward (flow warp + one temporal lifting level + pWave++ on the subbands) on a synthetic
    2-frame 256x448 clip (padded to 256x512 as test_pMCTF_flex.py does), bit-exact against the oracle."""
    from learned_pmctf_b200 import gop as Gm
    h0, w0, gop = 256, 448, 2
    codec = Gm.GopCodec(model, gop, q_index=8)
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    assert (hp, wp) == (256, 512)
    y, c, mvs = _inputs(gop, h0, w0, 11)
    mvs = [np.ascontiguousarray(np.pad(m, ((0, 0), (0, 0), (0, hp - h0), (0, wp - w0)))) for m in mvs]
    yd, cd = cu(y), cu(c)
    Y = P.ops.unpack_u8(yd, hp, wp)
    C = P.ops.unpack_u8(cd.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(gop, 2, 1, hp // 2, wp // 2)
    rec_y, rec_c, st = codec.code_gop(Y, C, [cu(m) for m in mvs], yd, cd)
    temporal, hp_w, lp_w = _oracle_weights(weights)
    oy, oc, osym = orc.code_gop(orc.unpack_u8(y, hp, wp)[:, None], orc.unpack_u8(c, hp // 2, wp // 2)[:, :, None], mvs, temporal,
                                hp_w, lp_w, [codec.q_pair("hp", 0)], codec.q_pair("lp", 0))
    assert np.array_equal(rec_y.cpu().numpy(), oy) and np.array_equal(rec_c.cpu().numpy(), oc)
    assert np.array_equal(st.cpu().numpy()[:, 1:3].astype(np.int64), osym)


def test_config1_pwave_1080p_six_q_points_vs_oracle(P, model, weights, conv_mode):
    """configs[1]: pWave++ analysis + quantise + dequantise + synthesis of one 1920x1080 frame (padded 1152x1920) at the
    6 q_index points of test_pMCTF_flex.py:436-443 -- every quantised symbol and the reconstruction bit-exact vs the oracle."""
    if conv_mode == "ffma":
        pytest.skip("full-size oracle run once (tensor mode); the ffma mode is covered at the smaller sizes")
    from learned_pmctf_b200 import gop as Gm
    y, _ = Gm.synthetic_sequence(3, 1, 1080, 1920, "cuda")
    X = P.ops.unpack_u8(y, 1152, 1920)
    x_np = X.cpu().numpy()
    _, _, lp_w = _oracle_weights(weights)
    y_or = orc.pwave_encode(x_np, lp_w)          # analysis once; quantisation per q point
    coder = model.lp_coder
    enc = coder.encode_bands(X)
    for lvl in range(4):
        for b in ("ll", "lh", "hl", "hh"):
            assert np.array_equal(enc[lvl][b].cpu().numpy(), y_or[lvl][b]), (lvl, b)
    nsym = 0
    for qi in (0, 4, 8, 12, 16, 20):
        q, qll = coder.q_pair(qi)
        q, qll = float(q.detach().reshape(-1)[0].cpu()), float(qll.detach().reshape(-1)[0].cpu())
        x_hat, hat = coder.spatial_wavelet_dec(X, q, qll, post_process=False, return_symbols=True)
        for lvl in hat:
            for b, v in hat[lvl].items():
                want = orc.quantize(y_or[lvl][b], qll if b == "ll" else q)
                assert np.array_equal(v.cpu().numpy(), want), f"q_index {qi} level {lvl} band {b}"
                nsym += want.size
        if qi in (0, 20):  # full synthesis vs the oracle at the two end points
            rec = {lvl: {b: orc.dequantize(v.cpu().numpy(), qll if b == "ll" else q) for b, v in hat[lvl].items()} for lvl in hat}
            assert np.array_equal(x_hat.cpu().numpy(), orc.pwave_decode(rec, lp_w))
    assert nsym == 6 * 1152 * 1920


@pytest.mark.parametrize("gop,h0,w0", [(4, 120, 190), (8, 256, 320), (16, 64, 96)])
def test_concurrent_chroma_equals_single_stream(P, model, gop, h0, w0):
    """GopCodec runs the chroma chain on a side stream (luma and chroma never meet on the path).  The two-stream run must be
    bit-identical to the single-stream one, repeatedly (a cross-stream race would show up as a flaky difference), and the
    public analysis()/code()/synthesis() trio must agree with both."""
    from learned_pmctf_b200 import gop as Gm
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    y, c, mvs = _inputs(gop, h0, w0, 23)
    mvs = [cu(np.ascontiguousarray(np.pad(m, ((0, 0), (0, 0), (0, hp - h0), (0, wp - w0))))) for m in mvs]
    yd, cd = cu(y), cu(c)
    Y = P.ops.unpack_u8(yd, hp, wp)
    C = P.ops.unpack_u8(cd.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(gop, 2, 1, hp // 2, wp // 2)
    serial = Gm.GopCodec(model, gop, q_index=12, concurrent_chroma=False)
    ry, rc, st = serial.code_gop(Y, C, mvs, yd, cd)
    torch.cuda.synchronize()
    Ly, Lc, Hs = serial.analysis(Y, C, mvs)
    Lyh, Lch, Hh, _, _ = serial.code(Ly, Lc, Hs)
    ty, tc = serial.synthesis(Lyh, Lch, Hh, mvs)
    assert torch.equal(ty, ry) and torch.equal(tc, rc)
    both = Gm.GopCodec(model, gop, q_index=12)
    assert both.concurrent_chroma
    for _ in range(4):
        by, bc, bst = both.code_gop(Y, C, mvs, yd, cd)
        by2, bc2, _ = both.code_gop(Y, C, mvs, yd, cd)     # back to back: the second GOP forks while the first still runs
        torch.cuda.synchronize()
        assert torch.equal(by, ry) and torch.equal(bc, rc) and torch.equal(bst, st)
        assert torch.equal(by2, ry) and torch.equal(bc2, rc)


def test_code_sequence_lanes_equal_gop_by_gop(P, model):
    """code_sequence() alternates GOPs between stream-pair lanes; its statistics must equal the GOP-by-GOP single-stream run."""
    from learned_pmctf_b200 import gop as Gm
    gop, h0, w0, n_gops = 4, 120, 190, 5
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    ys, cs, mvl = [], [], []
    for g in range(n_gops):
        y, c, mvs = _inputs(gop, h0, w0, 31 + g)
        ys.append(y), cs.append(c)
        mvl.append([cu(np.ascontiguousarray(np.pad(m, ((0, 0), (0, 0), (0, hp - h0), (0, wp - w0))))) for m in mvs])
    yd, cd = cu(np.concatenate(ys)), cu(np.concatenate(cs))
    Y = P.ops.unpack_u8(yd, hp, wp)
    C = P.ops.unpack_u8(cd.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(gop * n_gops, 2, 1, hp // 2, wp // 2)
    serial = Gm.GopCodec(model, gop, q_index=12, concurrent_chroma=False)
    want = torch.stack([serial.code_gop(Y[g * gop:(g + 1) * gop], C[g * gop:(g + 1) * gop], mvl[g], yd[g * gop:(g + 1) * gop],
                                        cd[g * gop:(g + 1) * gop])[2] for g in range(n_gops)])
    assert torch.equal(serial.code_sequence(Y, C, mvl, yd, cd), want)
    lanes = Gm.GopCodec(model, gop, q_index=12)
    lanes.GOP_LANES = 2
    for _ in range(3):
        assert torch.equal(lanes.code_sequence(Y, C, mvl, yd, cd), want)
    host = lanes.code_sequence_host(yd.cpu().pin_memory(), cd.cpu().pin_memory(), [[m.cpu().pin_memory() for m in g] for g in mvl])
    assert torch.equal(host, want.reshape(-1, want.size(-1)).cpu())
