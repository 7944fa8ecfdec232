"""Row a14 of SURVEY.md section 8 on a GPU: the reference's ENTRY POINTS over the B200 kernels.

(1) `accelerate()` on a reference-shaped model object (tests/ref_double.py: the reference's module tree, parameter names and
    the call sequence of pWave.forward_one_channel / pMCTF.forward_one_stage, pWave.py:231-312 / pMCTF_L.py:332-379) -- the
    rebinding that INTEGRATION.md describes for the real objects, executed where a GPU exists, compared bit-exactly with the
    oracle.
(2) The stand-alone classes' own `forward_one_stage` / `pWave.forward` (same signatures, same return keys).
(3) `GopCodec` on the accelerated object equals `GopCodec` on the stand-alone model.
(4) On a machine that has BOTH a GPU and /root/reference (not the driver's box, not the build container): the real object.
"""
import os
import sys

import numpy as np
import pytest
import torch

import ref_double
from conftest import sub_sd
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import learned_pmctf_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def _state(weights):
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    sd |= {k.replace("lift_h", "lift_v"): v for k, v in sd.items() if "lift_h" in k}
    return sd


@pytest.fixture(scope="module")
def double(P, weights):
    """The reference-shaped object with the golden weights, accelerated."""
    m = ref_double.pMCTF(num_me_stages=4).eval()
    sd = _state(weights)
    own = m.state_dict()
    keep = {k: own[k] for k in own if k not in sd}                     # QP / QP_ll / hp_q_scale endpoints set by the double
    missing = m.load_state_dict({k: v for k, v in sd.items() if k in own} | keep, strict=True)
    assert not missing.missing_keys
    m = m.cuda()
    before = {k: v.data_ptr() for k, v in m.state_dict().items()}
    m = P.accelerate(m)
    assert {k: v.data_ptr() for k, v in m.state_dict().items()} == before, "accelerate() must not touch the state_dict"
    assert ref_double.flow_warp is P.flow_warp                          # the names the model file bound at import
    return m


@pytest.fixture(scope="module")
def standalone(P, weights, double):
    m = P.pMCTF(num_me_stages=4).cuda().eval()
    m.load_reference_state_dict(_state(weights))
    with torch.no_grad():
        for a, b in ((m.lp_coder, double.lp_coder), (m.hp_coder, double.hp_coder)):
            a.QP.copy_(b.QP), a.QP_ll.copy_(b.QP_ll)
        for a, b in zip(m.hp_q_scale, double.hp_q_scale):
            a.copy_(b)
    return m


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _frame(shape, seed):
    g = np.random.default_rng(seed)
    n, c, h, w = shape
    base = g.random((h + 8, w + 8)) * 255
    base = sum(np.roll(base, (dy, dx), (0, 1)) for dy in range(-2, 3) for dx in range(-2, 3)) / 25.0
    return np.stack([np.rint(np.clip(base[4 + i:4 + i + h, 4:4 + w] + g.normal(0, 2, (h, w)), 0, 255)) for i in range(n)])[:, None].astype(np.float32)


def _q(t):
    return float(t.detach().reshape(-1)[0].cpu())


def _check_coder_output(res, x_np, w, q, qll):
    ox, ohat = orc.spatial_wavelet_dec(x_np, w, q, qll)
    for lvl in ohat:
        for b, v in ohat[lvl].items():
            assert np.array_equal(res["subbands"][lvl][b].cpu().numpy(), v), (lvl, b)
    assert np.array_equal(res["x_hat"].cpu().numpy(), ox)
    return ox


@pytest.mark.parametrize("q_index", [4, 16])
def test_accelerated_forward_one_channel_call_sequence(P, double, weights, q_index):
    """pWave.forward -> forward_one_channel of the reference-shaped object: encode -> quantize_subband (13x) ->
    dequantize_subbands -> decode run on the kernels, in the reference's order, with ONE device->host read per step tensor."""
    coder = double.hp_coder
    coder.calls.clear()
    x = _frame((2, 1, 128, 192), 5)
    reads = []
    orig = torch.Tensor.to

    def counting_to(self, *a, **k):
        if self.is_cuda and self.numel() == 1 and a and a[0] == "cpu":
            reads.append(1)
        return orig(self, *a, **k)

    torch.Tensor.to = counting_to
    try:
        with torch.no_grad():
            res = coder.forward(cu(x), q_index, qp_scale=None)
    finally:
        torch.Tensor.to = orig
    assert coder.calls == ["encode"] + ["quantize_subband"] * 13 + ["dequantize_subbands", "decode"]
    assert len(reads) <= 2, f"{len(reads)} host reads of step scalars in one forward_one_channel (one per step tensor expected)"
    w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    with torch.no_grad():
        q, qll = _q(coder.get_curr_q(coder.QP, q_index)), _q(coder.get_curr_q(coder.QP_ll, q_index))
    _check_coder_output(res, x, w, q, qll)


@pytest.mark.parametrize("which", ["accelerated", "standalone"])
def test_forward_one_stage_vs_oracle(P, double, standalone, weights, which):
    """forward_one_stage(ref, cur, q_index, code_lt=True, dpb, mv_hat=..., stage_idx=...) (pMCTF_L.py:332-379): chroma-shaped
    planes with the luma motion field, as the reference calls it with a given field (:333-336)."""
    m = double if which == "accelerated" else standalone
    stage, q_index, h, w = 2, 12, 64, 96
    ref, cur = _frame((2, 1, h, w), 11), _frame((2, 1, h, w), 12)
    g = np.random.default_rng(3)
    mv = g.normal(0, 3, (1, 2, 2 * h, 2 * w)).astype(np.float32)
    with torch.no_grad():
        out = m.forward_one_stage(cu(ref), cu(cur), q_index, True, None, mv_hat=cu(mv), stage_idx=stage)
    Pt, Ut = (orc.PU(sub_sd(weights, f"temporal_filtering.{stage}.{n}.")) for n in ("P_t", "U_t"))
    omv = orc.chroma_mv_down(mv)
    assert np.array_equal(out["mv_hat"].cpu().numpy(), omv)
    oL, oH, opred, _ = orc.forward_mctf(ref, cur, omv, Pt, Ut)
    hp_w = orc.IWave(sub_sd(weights, "hp_coder.wavelet_transform.lift_h."))
    lp_w = orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h."))
    with torch.no_grad():
        sc = m.get_curr_q(m.hp_q_scale[stage], q_index)
        q, qll = _q(m.hp_coder.get_curr_q(m.hp_coder.QP, q_index) * sc), _q(m.hp_coder.get_curr_q(m.hp_coder.QP_ll, q_index) * sc)
        ql, qlll = _q(m.lp_coder.get_curr_q(m.lp_coder.QP, q_index)), _q(m.lp_coder.get_curr_q(m.lp_coder.QP_ll, q_index))
    oHx, _ = orc.spatial_wavelet_dec(oH, hp_w, q, qll)
    oLx, _ = orc.spatial_wavelet_dec(oL, lp_w, ql, qlll)
    assert np.array_equal(out["H_t"].cpu().numpy(), oHx) and np.array_equal(out["L_t"].cpu().numpy(), oLx)
    assert abs(float(out["me_mse"]) - float(np.mean((opred.astype(np.float64) - cur) ** 2))) <= 1e-3 * max(1.0, float(out["me_mse"]))
    if which == "standalone":       # the reference's return keys (pMCTF_L.py:353-378)
        for k in ("bpp_mv_y", "bpp_mv_z", "bpp_me", "me_mse", "bpp", "bpp_H", "bit_H", "bit_ME", "mse_H", "mv_hat", "dpb", "H_t", "L_t",
                  "bpp_L", "bit_L", "mse_L", "me_mse_inv", "bit"):
            assert k in out, k
        assert torch.isnan(out["bpp"]) and out["bpp_mv_y"] is None     # entropy model out of scope: never an invented number


def test_standalone_pwave_forward_keys(P, standalone, weights):
    x = _frame((1, 1, 64, 64), 9)
    with torch.no_grad():
        res = standalone.lp_coder(cu(x), 8)
        q, qll = _q(standalone.lp_coder.q_pair(8)[0]), _q(standalone.lp_coder.q_pair(8)[1])
    assert set(res) == {"x_hat", "bits", "likelihoods", "subbands", "bpp_total", "bits_total", "mse"}     # pWave.py:302-310
    ox = _check_coder_output(res, x, orc.IWave(sub_sd(weights, "lp_coder.wavelet_transform.lift_h.")), q, qll)
    assert abs(float(res["mse"]) - float(np.mean((ox.astype(np.float64) - x) ** 2))) <= 1e-3 * max(1.0, float(res["mse"]))


def test_gopcodec_on_accelerated_object_equals_standalone(P, double, standalone):
    """GopCodec(accelerate(model)) -- the q_pair / hp_qp_scale helpers are grafted too (round-1 advisor finding)."""
    from learned_pmctf_b200 import gop as Gm
    G, h0, w0 = 4, 120, 190
    _, pr, _, pb = Gm.get_padding_size(h0, w0, 128)
    hp, wp = h0 + pb, w0 + pr
    g = torch.Generator(device="cuda").manual_seed(4)
    y = torch.randint(0, 256, (G, h0, w0), dtype=torch.uint8, device="cuda", generator=g)
    c = torch.randint(64, 192, (G, 2, h0 // 2, w0 // 2), dtype=torch.uint8, device="cuda", generator=g)
    Y = P.ops.unpack_u8(y, hp, wp)
    C = P.ops.unpack_u8(c.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(G, 2, 1, hp // 2, wp // 2)
    mvs = Gm.synthetic_motion(0, 0, G, hp, wp, "cuda")
    a = Gm.GopCodec(double, G, q_index=12).code_gop(Y, C, mvs, y, c)
    b = Gm.GopCodec(standalone, G, q_index=12).code_gop(Y, C, mvs, y, c)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree AND a GPU on the same machine")
def test_accelerated_real_reference_object_runs_on_gpu(P, weights):
    stubs = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_stubs")
    sys.path[:0] = [stubs, REF]
    try:
        from pMCTF.models.video.pMCTF_L import pMCTF
        torch.manual_seed(0)
        m = P.accelerate(pMCTF(num_me_stages=4).eval().cuda())
    finally:
        sys.path.remove(stubs), sys.path.remove(REF)
    x = cu(_frame((2, 1, 128, 192), 5))
    with torch.no_grad():
        res = m.hp_coder.forward(x, 8)                                   # the reference's own forward_one_channel body
        assert torch.isfinite(res["x_hat"]).all() and float(res["bpp_total"]) > 0
        out = m.forward_one_stage(x, x.flip(0), 8, True, None, mv_hat=torch.zeros(1, 2, 256, 384, device="cuda"), stage_idx=1)
        assert torch.isfinite(out["H_t"]).all() and torch.isfinite(out["L_t"]).all()
