"""Host mirror of pMCTF/entropy_models/gaussian_model.py:14-72 (CompressionModel: the rate estimate of the training /
evaluation forward pass and the owner of the entropy coder)."""
from __future__ import annotations

import math

import torch
from torch import nn

from ..layers.layers import RoundNoGradient
from .entropy_models import EntropyCoder, GaussianEncoder


class CompressionModel(nn.Module):
    def __init__(self, y_distribution, ec_thread: bool = False, stream_part: int = 1):
        super().__init__()
        self.y_distribution = y_distribution
        self.entropy_coder = None
        self.gaussian_encoder = GaussianEncoder(distribution=y_distribution)
        self.ec_thread = ec_thread
        self.stream_part = stream_part
        self.masks = {}

    def quant(self, x):
        return RoundNoGradient.apply(x) if self.training else torch.round(x)

    def get_curr_q(self, q_scale, q_basic, q_index=None):
        return q_basic * q_scale[q_index]

    @staticmethod
    def probs_to_bits(probs):
        return torch.clamp_min(-1.0 * torch.log(probs + 1e-5) / math.log(2.0), 0)

    @staticmethod
    def _interval_bits(dist, y, sigma):
        d = dist(torch.zeros_like(sigma), sigma.clamp(1e-5, 1e10))
        return CompressionModel.probs_to_bits(d.cdf(y + 0.5) - d.cdf(y - 0.5))

    def get_y_gaussian_bits(self, y, sigma):
        return self._interval_bits(torch.distributions.normal.Normal, y, sigma)

    def get_y_laplace_bits(self, y, sigma):
        if y.is_cuda and not torch.is_grad_enabled() and y.dtype == torch.float32 and y.shape == sigma.shape:
            # one kernel instead of ~25 ATen launches and the two host synchronisations of torch.distributions' argument checks
            from .. import _native as nat
            from .. import ops
            yc, sc = y.contiguous(), sigma.contiguous().float()
            out = torch.empty_like(yc)
            ops._launch(y.device, "laplace_bits", nat.lib().pmctf_laplace_bits, yc.data_ptr(), sc.data_ptr(), out.data_ptr(), yc.numel())
            return out
        return self._interval_bits(torch.distributions.laplace.Laplace, y, sigma)

    def update(self, force: bool = False):
        self.entropy_coder = EntropyCoder(self.ec_thread, self.stream_part)
        self.gaussian_encoder.update(force=force, entropy_coder=self.entropy_coder)

    def process(self, y, means):
        y_q = self.quant(y)
        y_res = y_q - means
        return y_res, y_q, y_res + means

    def get_z_bits(self, z, bit_estimator):
        return CompressionModel.probs_to_bits(bit_estimator.get_cdf(z + 0.5) - bit_estimator.get_cdf(z - 0.5))

    def add_noise(self, x):
        return x + torch.empty_like(x).uniform_(-0.5, 0.5)
