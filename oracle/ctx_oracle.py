"""TEST INFRASTRUCTURE -- fp32 CPU restatement of the four-step entropy-parameter network of pWave++.

Follows pMCTF/layers/context_fusion_4step.py (ContextResidual :9-20, ContextFusionFourStep.forward :127-194, process_with_mask
:115-125, decompress :196-249) and the DepthConvBlock head of pMCTF/layers/video/layers.py:113-172, on plain fp32 convolutions
(oracle/pmctf_oracle.c: orc_conv2d, one fma chain per output in (ci, ky, kx) order).  The CUDA path evaluates the 112 -> 112
layers with bf16 operands on the tensor cores, so this is a TOLERANCE oracle for the parameters (scales, means) and the reference
point for the final-symbol mismatch COUNT; it is pinned to outputs of the reference's own module (tests/golden/ctx4.npz,
oracle/make_golden.py ctx4).  Only tests/, smoke() and bench.py's checker legs may import this file."""
import ctypes as C

import numpy as np

from . import oracle as orc


def conv2d(x, w, b=None, groups=1):
    """x [ci,H,W], w [co,ci/groups,k,k] -> [co,H,W]"""
    x, w = orc._a(x), orc._a(w)
    b = orc._a(b) if b is not None else None
    ci, H, W = x.shape
    co, _, k, _ = w.shape
    out = np.empty((co, H, W), np.float32)
    orc.lib().orc_conv2d(orc._p(x), ci, orc._p(w), orc._p(b) if b is not None else None, co, k, groups, orc._p(out), H, W)
    return out


def lrelu(x, s):
    return np.where(x >= 0, x, x * np.float32(s)).astype(np.float32)


class FourStep:
    """Weights of one ContextFusionFourStep from a state_dict-like mapping (reference key names)."""

    def __init__(self, sd, prefix="", lossy=True):
        self.sd = {k[len(prefix):]: orc._a(v) for k, v in sd.items() if k.startswith(prefix)}
        self.lossy = lossy
        self.has_lower = "lower_level_subband.1.weight" in self.sd

    def conv(self, name, x, groups=1):
        return conv2d(x, self.sd[name + ".weight"], self.sd[name + ".bias"], groups)

    def resblock(self, name, x):
        return self.conv(name + ".conv2", lrelu(self.conv(name + ".conv1", x), 0.2)) + x

    def context(self, context, prev_subband=None):
        """[1,H,W] (+ [1,H/2,W/2]) -> the 112-channel context features (:128-133)"""
        if prev_subband is not None:
            up = np.repeat(np.repeat(prev_subband, 2, axis=1), 2, axis=2)
            context = np.concatenate([context, self.conv("lower_level_subband.1", up)], axis=0)
        c = self.conv("conv1_context", context)
        c = self.resblock("y_hierarchical_prior_enc.0", c)
        return self.resblock("y_hierarchical_prior_enc.1", c)

    def hierarchical(self, c):
        """DepthConvBlock(112, 2): DepthConv then ConvFFN -> (scales, means), each [1,H,W]"""
        p = "y_hierarchical_prior_out.block."
        t = lrelu(self.conv(p + "0.conv1.0", c), 0.01)
        t = self.conv(p + "0.depth_conv", t, groups=t.shape[0])
        y = self.conv(p + "0.conv2", t) + self.conv(p + "0.adaptor", c)
        f = lrelu(self.conv(p + "1.conv.2", lrelu(self.conv(p + "1.conv.0", y), 0.1)), 0.1)
        y = y + f
        return y[0:1], y[1:2]

    def spatial(self, k, x_hat_so_far, c):
        """step k = 1..3: parameters of the next mask from what has been coded so far (:147-151)"""
        t = self.conv(f"y_spatial_prior_{k}.0", x_hat_so_far)
        t = self.resblock(f"y_spatial_prior_{k}.1", t) + c
        t = self.resblock(f"y_spatial_prior_{k}_out.0", t)
        t = self.resblock(f"y_spatial_prior_{k}_out.1", t)
        y = self.conv(f"y_spatial_prior_{k}_out.2", t)
        return y[0:1], y[1:2]

    @staticmethod
    def mask(k, H, W):
        yy, xx = np.mgrid[0:H, 0:W]
        return ((2 * (yy & 1) + (xx & 1)) == k).astype(np.float32)[None]

    def forward_one(self, x, context, prev_subband=None, trace=None):
        """x, context: [1,H,W].  Returns x_res, x_q, x_hat, s_hat plus the per-step symbol / scale planes (compress, :188-194)."""
        _, H, W = x.shape
        c = self.context(context, prev_subband)
        scales, means = self.hierarchical(c)
        x_hat = np.zeros_like(x)
        x_res, x_q, s_hat = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        steps = []
        for k in range(4):
            if k > 0:
                scales, means = self.spatial(k, x_hat, c)
            if trace is not None:
                trace.append((scales.copy(), means.copy()))
            m = self.mask(k, H, W)
            mu = means if self.lossy else np.rint(means)
            mu_hat = mu * m
            res = (x - mu_hat) * m
            q = np.rint(res).astype(np.float32)
            x_hat = x_hat + (q + mu_hat)
            x_res, x_q, s_hat = x_res + res, x_q + q, s_hat + scales * m
            steps.append((q, scales * m))
        return x_res, x_q, x_hat, s_hat, steps

    def forward(self, x, context, prev_subband=None):
        outs = [self.forward_one(x[n], context[n], prev_subband[n] if prev_subband is not None else None) for n in range(x.shape[0])]
        stack = lambda i: np.stack([o[i] for o in outs])  # noqa: E731
        return stack(0), stack(1), stack(2), stack(3)
