"""SURVEY.md section 8c, third symbol check (build container only -- needs /root/reference): the reference's OWN context /
entropy-parameter nets (ContextFusionFourStep, ContextFusionSubband, SubbandContext: out of scope, stock torch) are fed
(a) the reference's subbands and (b) the subbands of this repo's arithmetic contract -- the oracle in tensor mode, which the
CUDA kernels reproduce bit for bit (tests/test_gpu_parity.py).  The symbols the context model finally rounds,
round((s - mu) ...) per subband (context_fusion_4step.py:127-137) and ll_hat (pWave.py:257), are compared and the mismatch
COUNT is reported and bounded: a flipped symbol needs (s - mu) within ~1e-5 of a rounding boundary."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


@pytest.mark.parametrize("q_index", [4, 16])
def test_symbols_after_reference_context_model(q_index, conv_mode):
    # both arithmetic contracts of the library: the tensor-core one (default) and the fp32 FMA-chain one
    from oracle import make_golden as MG      # imports the unmodified reference through oracle/ref_stubs
    from oracle import oracle as orc
    from conftest import sub_sd
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    m = MG.build_model()
    coder = m.lp_coder.eval()
    x = MG.frames(1, 128, 192, 77)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    w = orc.IWave(sub_sd(sd, "lp_coder.wavelet_transform.lift_h."))
    q = coder.get_curr_q(coder.QP, q_index)
    qll = coder.get_curr_q(coder.QP_ll, q_index)

    def run(encode=None):
        got = {}
        hooks = []
        for lvl, bands in coder.context_fusion.items():
            for band, mod in bands.items():
                if band != "ll":   # (s_res, s_q, s_hat, scales): s_q are the rounded symbols
                    hooks.append(mod.register_forward_hook(lambda _m, _i, out, key=(int(lvl), band): got.__setitem__(key, out[1].numpy().copy())))
        orig = coder.encode
        if encode is not None:
            coder.encode = encode
        try:
            with torch.no_grad():
                out = coder.forward_one_channel(x, q, qll)
        finally:
            coder.encode = orig
            for h in hooks:
                h.remove()
        got[(coder.decomp_levels - 1, "ll")] = out["subbands"][coder.decomp_levels - 1]["ll"].numpy().copy()
        return got

    ref = run()
    y = orc.pwave_encode(x.numpy(), w)
    ours = run(lambda _x: {lvl: {b: torch.from_numpy(np.ascontiguousarray(v)) for b, v in y[lvl].items()} for lvl in y})
    assert set(ref) == set(ours) and len(ref) == 3 * coder.decomp_levels + 1
    total = sum(v.size for v in ref.values())
    bad = {k: int(np.sum(ref[k] != ours[k])) for k in ref}
    n_bad = sum(bad.values())
    print(f"conv mode {conv_mode}, q_index {q_index}: {n_bad} of {total} context-model symbols differ ({ {k: v for k, v in bad.items() if v} })")
    assert total == 128 * 192
    assert all(np.all(v == np.rint(v)) for v in ours.values()), "symbols are integers"
    assert n_bad <= max(2, total // 5000), "more symbol flips than fp32 round-off near rounding boundaries explains"
