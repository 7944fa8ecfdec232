"""Motion-vector codec networks (reference: pMCTF/layers/video/video_net.py:124-193): analysis / synthesis of the decoded flow's
latent and its hyper-prior, host-side torch modules with the reference's module tree."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from .layers import DepthConvBlock, DepthConvBlock4, ResidualBlockUpsample, ResidualBlockWithStride, subpel_conv1x1


class LowerBound(Function):
    """max(x, bound) whose gradient also passes where it pulls x back up to the bound (video_net.py:14-30)"""

    @staticmethod
    def forward(ctx, inputs, bound):
        b = torch.ones_like(inputs) * bound
        ctx.save_for_backward(inputs, b)
        return torch.max(inputs, b)

    @staticmethod
    def backward(ctx, grad_output):
        inputs, b = ctx.saved_tensors
        keep = (inputs >= b) | (grad_output < 0)
        return keep.type(grad_output.dtype) * grad_output, None


class MvEnc(nn.Module):
    def __init__(self, input_channel, channel, inplace=False):
        super().__init__()
        self.enc_1 = nn.Sequential(ResidualBlockWithStride(input_channel, channel, stride=2, inplace=inplace),
                                   DepthConvBlock(channel, channel, inplace=inplace))
        self.enc_2 = ResidualBlockWithStride(channel, channel, stride=2, inplace=inplace)
        self.adaptor_0 = DepthConvBlock(channel, channel, inplace=inplace)
        self.adaptor_1 = DepthConvBlock(channel * 2, channel, inplace=inplace)
        self.enc_3 = nn.Sequential(ResidualBlockWithStride(channel, channel, stride=2, inplace=inplace),
                                   DepthConvBlock(channel, channel, inplace=inplace), nn.Conv2d(channel, channel, 3, stride=2, padding=1))

    def forward(self, x, context, quant_step):
        out = self.enc_2(self.enc_1(x) * quant_step)
        out = self.adaptor_0(out) if context is None else self.adaptor_1(torch.cat((out, context), dim=1))
        return self.enc_3(out)


class MvDec(nn.Module):
    def __init__(self, output_channel, channel, inplace=False):
        super().__init__()
        self.dec_1 = nn.Sequential(DepthConvBlock(channel, channel, inplace=inplace), ResidualBlockUpsample(channel, channel, 2, inplace=inplace),
                                   DepthConvBlock(channel, channel, inplace=inplace), ResidualBlockUpsample(channel, channel, 2, inplace=inplace),
                                   DepthConvBlock(channel, channel, inplace=inplace))
        self.dec_2 = ResidualBlockUpsample(channel, channel, 2, inplace=inplace)
        self.dec_3 = nn.Sequential(DepthConvBlock(channel, channel, inplace=inplace), subpel_conv1x1(channel, output_channel, 2))

    def forward(self, x, quant_step):
        feature = self.dec_1(x)
        return self.dec_3(self.dec_2(feature) * quant_step), feature


def get_hyper_enc_model(channel_N, channel_mv):
    return nn.Sequential(DepthConvBlock4(channel_mv, channel_N), nn.Conv2d(channel_N, channel_N, 3, stride=2, padding=1), nn.LeakyReLU(),
                         nn.Conv2d(channel_N, channel_N, 3, stride=2, padding=1))


def get_hyper_dec_model(channel_N, channel_mv):
    return nn.Sequential(ResidualBlockUpsample(channel_N, channel_N, 2), ResidualBlockUpsample(channel_N, channel_N, 2),
                         DepthConvBlock4(channel_N, channel_mv))
