"""Boundary hardening (round-1 verdict item 8 / advisor findings), on the GPU through the C ABI:
per-call convolution mode, tensor-core watchdog surfaced on the host, packed-weight registry eviction, cache invalidation,
pickling / deep copies, device guard."""
import copy
import ctypes as C
import io

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


def _pu(P, weights, prefix="temporal_filtering.0.P_t."):
    pu = P.PredictUpdate(1).cuda().eval()
    pu.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(weights, prefix).items()})
    return pu


def _x(seed=0, shape=(2, 1, 40, 56)):
    return (np.random.default_rng(seed).random(shape) * 255).astype(np.float32)


def test_conv_mode_is_per_call(P, weights, conv_mode):
    """The process default (set by the fixture) is what `conv_mode=None` uses; a call / a module can name the other
    arithmetic without touching the default -- each result bit-exact against the oracle run in that mode."""
    other = "ffma" if conv_mode == "tensor" else "tensor"
    pu, x = _pu(P, weights), _x()
    o = orc.PU(sub_sd(weights, "temporal_filtering.0.P_t."))
    want = {}
    for mode in (conv_mode, other):
        orc.set_conv_mode(mode)
        want[mode] = orc.predict_update(x, o)
    orc.set_conv_mode(conv_mode)
    assert not np.array_equal(want[conv_mode], want[other]), "the two arithmetics must differ somewhere for this test to mean anything"
    xd = torch.from_numpy(x).cuda()
    with torch.no_grad():
        assert np.array_equal(pu(xd).cpu().numpy(), want[conv_mode])
        assert np.array_equal(P.ops.predict_update(xd, pu.packed(), conv_mode=other).cpu().numpy(), want[other])
        pu.conv_mode = other                                            # pinned on the module
        assert np.array_equal(pu(xd).cpu().numpy(), want[other])
        pu.conv_mode = None
        assert np.array_equal(pu(xd).cpu().numpy(), want[conv_mode])
    assert P.ops.get_conv_mode() == conv_mode                           # the default was never touched
    # descriptors: TemporalLifting / iWave1D carry the field
    tl = P.TemporalLifting().cuda().eval()
    tl.conv_mode = other
    assert tl.descriptor().conv_mode == P.ops.conv_mode_code(other)
    iw = P.iWave1D().cuda().eval()
    d0 = iw.descriptor().conv_mode
    iw.conv_mode = other
    assert d0 == 0 and iw.descriptor().conv_mode == P.ops.conv_mode_code(other)
    # an unknown mode value in a raw descriptor is rejected, nothing is launched
    s = P._native.Step()
    s.conv_mode = 7
    assert P._native.lib().pmctf_lift_step(C.byref(s), None) == -1


def test_watchdog_flag_reaches_the_host(P, weights, conv_mode):
    """A tensor-core kernel that gave up waiting sets the device's watchdog word; from then on launches are REFUSED with a
    RuntimeError (PMCTF_ETIMEOUT) instead of returning partially written outputs with rc 0.  The word is injected here
    (pmctf_tc_inject_timeout) -- the kernel's own path to it is the same mapped host word."""
    if conv_mode != "tensor":
        pytest.skip("watchdog belongs to the tensor-core kernel")
    from learned_pmctf_b200 import gop as Gm
    lib = P._native.lib()
    pu, xd = _pu(P, weights), torch.from_numpy(_x()).cuda()
    with torch.no_grad():
        ok = pu(xd)
        torch.cuda.synchronize()
        assert P.ops.tc_error_flag() == 0
        assert lib.pmctf_tc_inject_timeout() == 0
        try:
            with pytest.raises(RuntimeError, match="gave up"):
                pu(xd)
            with pytest.raises(RuntimeError, match="gave up"):                      # repacking is refused as well
                P.ops.pack_pu(pu.ordered_params(), torch.empty(P._native.PU_PACKED_FLOATS, device="cuda"))
            with pytest.raises(RuntimeError, match="tensor-core kernel gave up"):
                P.ops.check_tc_error(xd.device)
            m = P.pMCTF(num_me_stages=1).cuda().eval()
            y = torch.zeros((2, 16, 16), dtype=torch.uint8)
            c = torch.zeros((2, 2, 8, 8), dtype=torch.uint8)
            with pytest.raises(RuntimeError, match="gave up"):                      # the public end-to-end call raises, not rc 0
                Gm.GopCodec(m, 2, q_index=8).code_sequence_host(y, c, [[torch.zeros(1, 2, 128, 128)]])
            # the FFMA arithmetic does not depend on the tensor cores and keeps working
            assert torch.isfinite(P.ops.predict_update(xd, pu.packed(), conv_mode="ffma")).all()
        finally:
            P.ops.clear_tc_error()
        assert torch.equal(pu(xd), ok)


def test_repack_and_release_of_packed_weights(P, weights, conv_mode):
    """The registry of per-block parameters follows the block: a repack at the same address gives the NEW result; a released
    block is rejected (tensor mode) instead of running with stale conv1 / conv4 / biases."""
    pu, xd = _pu(P, weights), torch.from_numpy(_x(1)).cuda()
    o0 = orc.PU(sub_sd(weights, "temporal_filtering.0.P_t."))
    o1 = orc.PU(sub_sd(weights, "temporal_filtering.1.U_t."))
    with torch.no_grad():
        a = pu(xd)
        addr = pu.packed().data_ptr()
        pu.load_state_dict({k: torch.from_numpy(v) for k, v in sub_sd(weights, "temporal_filtering.1.U_t.").items()})
        b = pu(xd)                                                      # load_state_dict invalidated the pack
        assert pu.packed().data_ptr() == addr, "the repack reuses the block (same device address)"
        assert np.array_equal(a.cpu().numpy(), orc.predict_update(xd.cpu().numpy(), o0))
        assert np.array_equal(b.cpu().numpy(), orc.predict_update(xd.cpu().numpy(), o1))
        # a write that bypasses the version counter needs invalidate_packed()
        pu.conv4.bias.data.add_(1.0)
        assert torch.equal(pu(xd), b)                                   # stale by design until told
        pu.invalidate_packed()
        c = pu(xd)
        assert float((c - b).abs().max()) > 0.5
        # release: the address is forgotten
        packed = pu.packed()
        P.ops.release_pu(packed)
        if conv_mode == "tensor":
            with pytest.raises(RuntimeError, match="invalid argument"):
                P.ops.predict_update(xd, packed)
        pu.invalidate_packed()
        assert torch.equal(pu(xd), c)                                   # packed again, registered again


def test_deepcopy_and_pickle_after_descriptors_exist(P, weights):
    m = P.pMCTF(num_me_stages=2).cuda().eval()
    xd = torch.from_numpy(_x(2, (1, 1, 32, 48))).cuda()
    mv = torch.zeros(1, 2, 32, 48, device="cuda")
    with torch.no_grad():
        a = m.forward_MCTF(xd, xd.flip(-1), mv)[1]
        e = m.hp_coder.encode_bands(xd)[0]["hh"]                         # builds the ctypes descriptor (device pointers inside)
        m2 = copy.deepcopy(m)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        m3 = torch.load(buf, weights_only=False)
        for mm in (m2, m3):
            assert torch.equal(mm.forward_MCTF(xd, xd.flip(-1), mv)[1], a)
            assert torch.equal(mm.hp_coder.encode_bands(xd)[0]["hh"], e)
        with torch.no_grad():
            m2.temporal_filtering[0].P_t.conv1.weight.mul_(2.0)           # the copy owns its weights and its packed block
        assert torch.equal(m.forward_MCTF(xd, xd.flip(-1), mv)[1], a)
        assert not torch.equal(m2.forward_MCTF(xd, xd.flip(-1), mv)[1], a)


def test_qcache_follows_parameter_versions(P):
    from learned_pmctf_b200 import gop as Gm
    m = P.pMCTF(num_me_stages=2).cuda().eval()
    codec = Gm.GopCodec(m, 4, q_index=10)
    q0 = codec.q_pair("hp", 1)
    with torch.no_grad():
        m.hp_coder.QP.mul_(0.5)
    assert codec.q_pair("hp", 1)[0] == pytest.approx(q0[0] * 0.5) and codec.q_pair("hp", 1)[1] == q0[1]
    with torch.no_grad():
        m.hp_q_scale[1].mul_(2.0)
    assert codec.q_pair("hp", 1)[1] == pytest.approx(q0[1] * 2.0)


def test_operands_on_different_devices_are_rejected(P):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    a = torch.zeros(1, 1, 16, 16, device="cuda:0")
    f = torch.zeros(1, 2, 16, 16, device="cuda:1")
    with pytest.raises(RuntimeError, match="different devices"):
        P.flow_warp(a, f)


def test_device_guard_other_than_current_device(P, weights):
    """Tensors on cuda:1 while cuda:0 is current: the launch goes to cuda:1's stream and context (ATen ops guard the same way)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    assert torch.cuda.current_device() == 0
    x = _x(3)
    pu0, pu1 = _pu(P, weights), _pu(P, weights).to("cuda:1")
    with torch.no_grad():
        a = pu0(torch.from_numpy(x).cuda())
        b = pu1(torch.from_numpy(x).to("cuda:1"))
        im = torch.from_numpy(x).to("cuda:1")
        w = P.flow_warp(im, torch.ones(2, 2, 40, 56, device="cuda:1"))
    assert b.device.index == 1 and torch.equal(a.cpu(), b.cpu()) and w.device.index == 1
    assert torch.cuda.current_device() == 0
