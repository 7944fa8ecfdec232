"""Four-part (quad-tree) spatial prior of the motion-vector latent (reference: pMCTF/layers/video/four_part_prior.py:11-279).

The latent's channels are split into four quarters and its pixels into the four positions of a 2 x 2 pattern; in step s quarter q
is coded at pattern position ORDER[s][q], so after four steps every (quarter, position) pair has been coded once, each step
conditioned on everything coded before it."""
from __future__ import annotations

import torch
from torch import nn

from .mv_codec import LowerBound

ORDER = ((0, 1, 2, 3), (3, 2, 1, 0), (2, 3, 0, 1), (1, 0, 3, 2))     # four_part_prior.py:124-171: mask index per channel quarter


class MVCoderQuad(nn.Module):
    def __init__(self, enc_dec_quant=False):
        super().__init__()
        self.enc_dec_quant = enc_dec_quant
        self.masks = {}

    def quant(self, x, force_detach=False):
        if self.training or force_detach:
            return x + (torch.round(x) - x).clone().detach()
        return torch.round(x)

    @staticmethod
    def separate_prior(params):
        return params.chunk(3, 1)

    @staticmethod
    def separate_prior_enc_dec(params):
        quant_step, scales, means = params.chunk(3, 1)
        quant_step = LowerBound.apply(quant_step, 0.5)
        return 1.0 / quant_step, quant_step, scales, means

    def process_with_mask(self, y, scales, means, mask):
        means_hat = means * mask
        y_res = (y - means_hat) * mask
        y_q = self.quant(y_res)
        return y_res, y_q, y_q + means_hat, scales * mask

    def get_mask_four_parts(self, height, width, dtype, device):
        key = f"{width}x{height}"
        if key not in self.masks:
            yy, xx = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
            code = 2 * (yy & 1) + (xx & 1)
            self.masks[key] = [(code == k).to(dtype)[None, None] for k in range(4)]
        return self.masks[key]

    def forward_four_part_prior(self, y, common_params, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3,
                                y_spatial_prior, write=False):
        if self.enc_dec_quant:
            q_enc, q_dec, scales, means = self.separate_prior_enc_dec(common_params)
            y = y * q_enc
        else:
            quant_step, scales, means = self.separate_prior(common_params)
            quant_step = torch.clamp_min(quant_step, 0.5)
            y = y / quant_step
        masks = self.get_mask_four_parts(y.size(2), y.size(3), y.dtype, y.device)
        parts = y.chunk(4, 1)
        adaptors = (None, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3)
        sc, mu = scales.chunk(4, 1), means.chunk(4, 1)
        so_far = None
        res = [[None] * 4 for _ in range(4)]      # res[quarter][mask index] = (y_res, y_q, y_hat, s_hat)
        for step in range(4):
            if step > 0:
                p = y_spatial_prior(adaptors[step](torch.cat((so_far, common_params), dim=1))).chunk(8, 1)
                sc, mu = p[:4], p[4:]
            outs = [self.process_with_mask(parts[q], sc[q], mu[q], masks[ORDER[step][q]]) for q in range(4)]
            for q in range(4):
                res[q][ORDER[step][q]] = outs[q]
            cur = torch.cat([o[2] for o in outs], dim=1)
            so_far = cur if so_far is None else so_far + cur

        def combine(i):                            # per quarter the sum over its four mask positions, in mask order (:84-93)
            return torch.cat([res[q][0][i] + res[q][1][i] + res[q][2][i] + res[q][3][i] for q in range(4)], dim=1)
        y_hat = combine(2) * (q_dec if self.enc_dec_quant else quant_step)
        if write:                                  # per step the sum over the quarters, in quarter order (:190-198)
            def per_step(i, step):
                t = [res[q][ORDER[step][q]][i] for q in range(4)]
                return t[0] + t[1] + t[2] + t[3]
            return (*[per_step(1, s) for s in range(4)], *[per_step(3, s) for s in range(4)], y_hat)
        return combine(0), combine(1), y_hat, combine(3)

    def compress_four_part_prior(self, y, common_params, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3,
                                 y_spatial_prior):
        return self.forward_four_part_prior(y, common_params, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2,
                                            y_spatial_prior_adaptor_3, y_spatial_prior, write=True)

    def decompress_four_part_prior(self, common_params, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3,
                                   y_spatial_prior, gaussian_encoder):
        if self.enc_dec_quant:
            _, quant_step, scales, means = self.separate_prior_enc_dec(common_params)
        else:
            quant_step, scales, means = self.separate_prior(common_params)
            quant_step = torch.clamp_min(quant_step, 0.5)
        dtype, device = means.dtype, means.device
        masks = self.get_mask_four_parts(means.size(2), means.size(3), dtype, device)
        adaptors = (None, y_spatial_prior_adaptor_1, y_spatial_prior_adaptor_2, y_spatial_prior_adaptor_3)
        sc, mu = scales.chunk(4, 1), means.chunk(4, 1)
        so_far = None
        for step in range(4):
            if step > 0:
                p = y_spatial_prior(adaptors[step](torch.cat((so_far, common_params), dim=1))).chunk(8, 1)
                sc, mu = p[:4], p[4:]
            m = [masks[ORDER[step][q]] for q in range(4)]
            scales_r = sc[0] * m[0] + sc[1] * m[1] + sc[2] * m[2] + sc[3] * m[3]
            y_q_r = gaussian_encoder.decode_stream(scales_r, dtype, device)
            cur = torch.cat([(y_q_r + mu[q]) * m[q] for q in range(4)], dim=1)
            so_far = cur if so_far is None else so_far + cur
        return so_far * quant_step
