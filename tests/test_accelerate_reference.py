"""Drop-in contract on the REFERENCE model object (only where /root/reference exists, i.e. the build container):
accelerate() keeps the state_dict (keys, shapes, tensor identity) of the reference pMCTF and swaps the hot-path
modules / methods for the B200 ones.  No kernel is launched."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
STUBS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_stubs")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def ref_model():
    sys.path[:0] = [STUBS, REF]
    try:
        from pMCTF.models.video.pMCTF_L import pMCTF
        torch.manual_seed(0)
        return pMCTF(num_me_stages=4).eval()
    finally:
        sys.path.remove(STUBS)
        sys.path.remove(REF)


def test_accelerate_keeps_state_dict_and_swaps_hot_path(ref_model):
    import learned_pmctf_b200 as P
    before = {k: (v.data_ptr(), tuple(v.shape)) for k, v in ref_model.state_dict().items()}
    assert len(before) == 3224
    m = P.accelerate(ref_model)
    after = {k: (v.data_ptr(), tuple(v.shape)) for k, v in m.state_dict().items()}
    assert after == before, "state_dict keys / shapes / storage must be untouched"
    # hot-path modules are ours, everything else is still the reference's
    assert all(type(t).__module__.startswith("learned_pmctf_b200") for t in m.temporal_filtering)
    for coder in (m.lp_coder, m.hp_coder):
        assert type(coder.wavelet_transform).__module__.startswith("learned_pmctf_b200")
        assert coder.wavelet_transform.lift_v is coder.wavelet_transform.lift_h
        assert type(coder.dequantModule).__module__.startswith("learned_pmctf_b200")      # PostProcess (section 8f row 2) is ours now
        for lvl in coder.context_fusion:                                                   # the four-step models are ours (section 8f row 1) ...
            for band in ("lh", "hl", "hh"):
                assert type(coder.context_fusion[lvl][band]).__module__.startswith("learned_pmctf_b200")
        assert type(coder.context_fusion["3"]["ll"]).__module__.startswith("learned_pmctf_b200")   # ... and so is the LL model; the ConvLSTM context stays
        assert type(coder.context_prediction).__module__.startswith("pMCTF.")
        assert coder.encode.__func__ is sys.modules["learned_pmctf_b200.models.pWave"].pWaveTransform.encode
    assert type(m.optic_flow).__module__.startswith("learned_pmctf_b200")                 # SpyNet (section 8f row 4) is ours now
    assert all(type(x).__module__.startswith("pMCTF.") for x in m.mv_encoder)             # the MV codec is not
    assert m.forward_MCTF.__func__ is sys.modules["learned_pmctf_b200.models.video.pMCTF_L"].MCTFMixin.forward_MCTF
    # a strict load of the reference's own checkpoint format still works
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()}, strict=True)
    # and the hot path refuses to run on the CPU instead of silently falling back
    with pytest.raises(RuntimeError, match="CUDA"):
        m.forward_MCTF(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16), torch.zeros(1, 2, 16, 16))


def test_standalone_model_loads_reference_checkpoint(ref_model):
    import learned_pmctf_b200 as P
    ours = P.pMCTF(num_me_stages=4)
    ours.load_reference_state_dict(ref_model.state_dict())
    sd = ref_model.state_dict()
    for k, v in ours.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert len(ours.state_dict()) == 4 * 16 + 4 + 2 * (2 * 40 + 2)   # temporal nets, hp_q_scale, both coders' lifting + QP
