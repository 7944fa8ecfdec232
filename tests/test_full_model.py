"""The whole model surface (SURVEY.md section 8a row a14 + 8f rows 1-4 together): pMCTF(motion=True, entropy_model=True) against the
reference's pMCTF -- strict state_dict parity (3 224 entries), the MV codec bit-identical on the CPU, the motion bitstream byte-
identical -- and on the GPU forward_one_stage without a given field plus compress / decompress of motion and frames."""
import os
import sys

import numpy as np
import pytest
import torch

import learned_pmctf_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _perturb(m, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.startswith(("mv_", "optic_flow")) and "q_scale" not in k:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            if "q_scale" in k:
                p.copy_(torch.tensor([0.8, 1.3]).view(2, 1, 1, 1))


@pytest.fixture(scope="module")
def ref_and_ours():
    if not os.path.isdir("/root/reference/pMCTF"):
        pytest.skip("reference tree not present")
    sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference"]
    saved = {k: sys.modules.get(k) for k in ("pMCTF.models.MLCodec_rans", "pMCTF.models.MLCodec_CXX")}
    try:
        sys.modules["pMCTF.models.MLCodec_rans"] = P.models.MLCodec_rans
        sys.modules["pMCTF.models.MLCodec_CXX"] = P.models.MLCodec_CXX
        from pMCTF.models.video.pMCTF_L import pMCTF as Ref
        torch.manual_seed(0)
        ref = Ref(num_me_stages=4).eval()
        _perturb(ref)
        ours = P.pMCTF(num_me_stages=4, entropy_model=True, motion=True).eval()
        yield ref, ours
    finally:
        del sys.path[:2]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_state_dict_and_mv_codec_equal_reference(ref_and_ours, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    ref, ours = ref_and_ours
    sr, so = ref.state_dict(), ours.state_dict()
    assert len(sr) == 3224 and set(sr) == set(so)
    assert all(tuple(sr[k].shape) == tuple(so[k].shape) for k in sr)
    ours.load_state_dict(sr, strict=True)
    ours.optic_flow.forward = ref.optic_flow.forward       # CPU: both MV codecs are fed the same estimated flow
    try:
        g = torch.Generator().manual_seed(2)
        cur = torch.rand(1, 1, 128, 192, generator=g) * 255
        prev = torch.roll(cur, (1, -2), (2, 3))
        dpb = {"mv_feature": None, "ref_mv_y": None}
        with torch.no_grad():
            a = ref.compute_and_code_motion(prev, cur, 12, dpb, stage_idx=1)
            b = ours.compute_and_code_motion(prev, cur, 12, dpb, stage_idx=1)
            assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
            dpb2 = {"mv_feature": a[1]["mv_feature"], "ref_mv_y": a[1]["mv_y_hat"]}
            a2 = ref.compute_and_code_motion(cur, prev, 12, dpb2, stage_idx=0)
            b2 = ours.compute_and_code_motion(cur, prev, 12, dpb2, stage_idx=0)
            assert torch.equal(a2[0], b2[0]) and torch.equal(a2[1]["mv_y_hat"], b2[1]["mv_y_hat"])
            # the motion bitstream: same bytes, and each side decodes the other's
            ref.update(force=True)
            ours.update(force=True)
            ca = ref.compress_mv(prev, cur, dpb, stage_idx=1, q_index=12)
            cb = ours.compress_mv(prev, cur, dpb, stage_idx=1, q_index=12)
            assert ca["bit_stream"] == cb["bit_stream"] and torch.equal(ca["mv_hat"], cb["mv_hat"])
            da = ref.decompress_mv(cb["bit_stream"], torch.float32, 128, 192, dpb, stage_idx=1, q_index=12)
            db = ours.decompress_mv(ca["bit_stream"], torch.float32, 128, 192, dpb, stage_idx=1, q_index=12)
            assert torch.equal(da["mv_hat"], cb["mv_hat"]) and torch.equal(db["mv_hat"], ca["mv_hat"])
    finally:
        del ours.optic_flow.forward


@pytest.mark.gpu
def test_full_stage_on_gpu(tmp_path, conv_mode):
    """forward_one_stage WITHOUT a given motion field (SpyNet + MV codec + lifting + both coders), the motion bitstream and the
    frame bitstreams of one stage, on the GPU."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_pwave_coder import _randomise
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = P.pMCTF(num_me_stages=4, entropy_model=True, motion=True)
    _randomise(m.lp_coder, 1), _randomise(m.hp_coder, 2)
    _perturb(m)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.startswith("temporal_filtering") and p.dim() == 4:
                p.normal_(0, 0.08)
    m = m.to(dev).eval()
    m.update(force=True)
    g = np.random.default_rng(5)
    base = g.random((1, 1, 136, 200)) * 255
    base = sum(np.roll(base, (dy, dx), (2, 3)) for dy in range(-2, 3) for dx in range(-2, 3)) / 25.0
    cur = torch.from_numpy(np.round(base[:, :, 4:132, 4:196]).astype(np.float32)).to(dev)
    ref = torch.roll(cur, (1, -2), (2, 3)).contiguous()
    dpb = {"mv_feature": None, "ref_mv_y": None}
    with torch.no_grad():
        out = m.forward_one_stage(ref, cur, 12, True, dpb, stage_idx=0)
    for k in ("bpp_mv_y", "bpp_mv_z", "bpp_me", "bpp", "bpp_H", "bpp_L", "bit", "me_mse", "mse_H", "mse_L"):
        assert torch.isfinite(out[k]).all(), k
    assert tuple(out["mv_hat"].shape) == (1, 2, 128, 192) and out["dpb"]["mv_feature"] is not None
    # motion bitstream round trip
    c = m.compress_mv(ref, cur, dpb, stage_idx=0, q_index=12)
    d = m.decompress_mv(c["bit_stream"], torch.float32, 128, 192, dpb, stage_idx=0, q_index=12)
    assert torch.equal(c["mv_hat"], d["mv_hat"]) and len(c["bit_stream"]) > 8
    # frame bitstreams of the stage
    name = str(tmp_path / "1_main.bin")
    enc = m.compress_one_stage(ref, cur, True, c["mv_hat"], False, sideinfo=(1, 1, 128, 192), file_name=name, stage_idx=0, q_index=12)
    dec = m.decompress_one_stage(name, True, False, psize=64, q_index=12, stage_idx=0)
    # like the reference, decompress_one_stage hands back pWave.decompress's dicts (pMCTF_L.py:429-439)
    assert torch.equal(dec["H_t"]["x_hat"], enc["H_t_hat"]) and torch.equal(dec["L_t"]["x_hat"], enc["L_t_hat"])
    r, cc = m.inverse_MCTF(dec["L_t"]["x_hat"], dec["H_t"]["x_hat"], c["mv_hat"], stage_idx=0)
    assert torch.isfinite(r).all() and float(((cc - cur) ** 2).mean()) < 5000.0     # random-weight PostProcess: a sanity bound, not a quality claim
    assert P.ops.tc_error_flag() == 0


@pytest.mark.gpu
def test_gop4_forward_and_bitstream_paths(tmp_path, conv_mode):
    """The reference's GOP loop on the full model at a small size: the forward (rate-estimate) path and the bitstream path with
    the decoder reading the files back reconstruct the same frames where the same symbols are coded (the LL band's
    autoregressive path and the skip_decoding shortcut are both exercised)."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_pwave_coder import _randomise
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = P.pMCTF(num_me_stages=4, entropy_model=True, motion=True)
    _randomise(m.lp_coder, 1), _randomise(m.hp_coder, 2)
    _perturb(m)
    m = m.to(dev).eval()
    m.update(force=True)
    g = np.random.default_rng(6)
    base = g.random((1, 1, 136, 272)) * 255          # 128 x 256: multiples of the 128-pixel padding unit the decoder assumes (psize)
    base = sum(np.roll(base, (dy, dx), (2, 3)) for dy in range(-2, 3) for dx in range(-2, 3)) / 25.0
    ys, cs = [], []
    for t in range(4):
        f = np.round(base[:, :, 4:132, 4 + 2 * t:260 + 2 * t]).astype(np.float32)
        ys.append(torch.from_numpy(f).to(dev))
        cs.append(torch.from_numpy(np.stack([f[0, :, ::2, ::2], 255 - f[0, :, ::2, ::2]])).to(dev))
    ry, rc, bits = m.code_gop_forward(ys, cs, q_index=12)
    assert len(ry) == 4 and all(torch.isfinite(t).all() for t in ry + rc) and all(b > 0 for b in bits)
    # H frames of a stage as one batch per coder call: same reconstruction and rate (every frame is coded independently)
    ry3, rc3, bits3 = m.code_gop_forward(ys, cs, q_index=12, batched=True)
    # (the stock-torch parts -- ConvLSTM context, LL model -- may pick another cuDNN algorithm for another batch size, so "same" is to
    # rounding: isolated symbols may flip, the bulk and the rate agree)
    for a, b in zip(ry + rc, ry3 + rc3):
        assert float((a - b).abs().mean()) < 0.05 * max(1.0, float(a.abs().mean()) / 100)
    assert abs(sum(bits) - sum(bits3)) < 0.01 * sum(bits)
    folder = str(tmp_path)
    ry2, rc2, bits2 = m.code_gop_forward(ys, cs, q_index=12, bin_folder=folder, skip_decoding=False)
    assert all(torch.isfinite(t).all() for t in ry2 + rc2)
    assert os.path.exists(os.path.join(folder, "1.bin")) and os.path.exists(os.path.join(folder, "1_mv.bin")) and os.path.exists(os.path.join(folder, "0_main.bin"))
    # the estimate and the real stream sizes agree to within the coder's overheads
    assert 0.6 * sum(bits) < sum(bits2) < 1.6 * sum(bits) + 4000, (sum(bits), sum(bits2))
    assert P.ops.tc_error_flag() == 0
