"""The fp32 SpyNet oracle against the vectors the reference's ME_Spynet produced (tests/golden/spynet.npz)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import spynet_weights  # noqa: E402


def test_spynet_oracle_reproduces_reference(golden, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    from oracle import spynet_oracle as so
    g = golden("spynet")
    sd = spynet_weights.make(int(g["seed"]))
    flow = so.spynet(sd, np.repeat(g["cur"], 3, axis=1), np.repeat(g["ref"], 3, axis=1))
    assert np.abs(flow - g["flow"]).max() < 5e-5          # fp32 summation order of the 7x7 layers (MKLDNN vs one fma chain)


def test_ctx_oracle_reproduces_reference(golden, conv_mode):
    """oracle/ctx_oracle.py against the reference's ContextFusionFourStep: identical final symbols, parameters to 1e-4."""
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    import ctx_weights
    from oracle import ctx_oracle as co
    g = golden("ctx4")
    for tag in "ab":
        cc = int(g[f"{tag}.ctx_channels"])
        m = co.FourStep(ctx_weights.make(int(g[f"{tag}.seed"]), cc))
        x_res, x_q, x_hat, s_hat = m.forward(g[f"{tag}.x"], g[f"{tag}.context"], g[f"{tag}.prev"] if cc == 2 else None)
        assert np.array_equal(x_q, g[f"{tag}.x_q"])
        for got, name in ((x_res, "x_res"), (x_hat, "x_hat"), (s_hat, "s_hat")):
            assert np.abs(got - g[f"{tag}.{name}"]).max() < 1e-4
