"""Seeded weights of a ContextFusionFourStep (reference key names and shapes, context_fusion_4step.py:23-98 with
num_features = 112, num_parameters = 2), shared by oracle/make_golden.py (which loads them into the reference's module) and the
tests, so the 10 MB of weights per module are regenerated instead of committed.  numpy's PCG64 stream is stable across versions.

Scaling: 112 -> 112 layers N(0, 0.022) (gain just below 1 through conv + LeakyReLU + skip, so activations stay O(1) through the 22
layers), 1|2 -> 112 N(0, 0.25), the depthwise layer N(0, 0.3), the 112 -> 2 projections N(0, 0.03), biases N(0, 0.05); the `scales` output channel gets
a bias of 1.5 so the predicted scales spread over the useful part of the table instead of clamping at its lower end."""
import numpy as np

F = 112


def shapes(ctx_channels: int):
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    def res(name):
        conv(name + ".conv1", F, F, 3)
        conv(name + ".conv2", F, F, 3)

    res("y_hierarchical_prior_enc.0")
    res("y_hierarchical_prior_enc.1")
    conv("conv1_context", F, ctx_channels, 3)
    if ctx_channels > 1:
        conv("lower_level_subband.1", 1, 1, 3)
    p = "y_hierarchical_prior_out.block."
    conv(p + "0.conv1.0", F, F, 1)
    conv(p + "0.depth_conv", F, 1, 3)
    conv(p + "0.conv2", 2, F, 1)
    conv(p + "0.adaptor", 2, F, 1)
    conv(p + "1.conv.0", 8, 2, 1)
    conv(p + "1.conv.2", 2, 8, 1)
    for k in (1, 2, 3):
        conv(f"y_spatial_prior_{k}.0", F, 1, 3)
        res(f"y_spatial_prior_{k}.1")
        res(f"y_spatial_prior_{k}_out.0")
        res(f"y_spatial_prior_{k}_out.1")
        conv(f"y_spatial_prior_{k}_out.2", 2, F, 1)
    return s


def make(seed: int, ctx_channels: int):
    g = np.random.default_rng(seed)
    out = {}
    for name, shp in shapes(ctx_channels).items():
        if name.endswith(".bias"):
            v = 0.05 * g.standard_normal(shp)
            if shp == (2,) and ("_out.2" in name or "conv2" in name):
                v[0] += 1.5
        elif shp[0] == F and shp[1] == F:
            v = (0.022 if shp[2] == 3 else 0.08) * g.standard_normal(shp)
        elif shp[1] in (1, 2) and shp[0] == F and "depth_conv" not in name:
            v = 0.25 * g.standard_normal(shp)
        elif "depth_conv" in name:
            v = 0.3 * g.standard_normal(shp)
        elif shp[0] == 2 and shp[1] == F:
            v = 0.03 * g.standard_normal(shp)
        else:
            v = 0.3 * g.standard_normal(shp)
        out[name] = v.astype(np.float32)
    return out
