"""Building blocks of the motion-vector codec (reference: pMCTF/layers/video/layers.py:22-192, the DCVC-DC block family): host-side
torch modules with the reference's parameter names.  The MV codec works on the 1/16-resolution latent of the motion field (64-192
channels, < 1 % of the codec's FLOPs); it runs on stock convolutions."""
from __future__ import annotations

import torch.nn as nn


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def subpel_conv1x1(in_ch, out_ch, r=1):
    """1x1 convolution to out_ch * r^2 channels + pixel shuffle (layers.py:35-39)"""
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=1, padding=0), nn.PixelShuffle(r))


class ResidualBlockWithStride(nn.Module):   # layers.py:47-80
    def __init__(self, in_ch, out_ch, stride=2, inplace=False):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=inplace)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.leaky_relu2 = nn.LeakyReLU(negative_slope=0.1, inplace=inplace)
        self.downsample = conv1x1(in_ch, out_ch, stride=stride) if stride != 1 else None

    def forward(self, x):
        out = self.leaky_relu2(self.conv2(self.leaky_relu(self.conv1(x))))
        return out + (self.downsample(x) if self.downsample is not None else x)


class ResidualBlockUpsample(nn.Module):     # layers.py:83-110
    def __init__(self, in_ch, out_ch, upsample=2, inplace=False):
        super().__init__()
        self.subpel_conv = subpel_conv1x1(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=inplace)
        self.conv = conv3x3(out_ch, out_ch)
        self.leaky_relu2 = nn.LeakyReLU(negative_slope=0.1, inplace=inplace)
        self.upsample = subpel_conv1x1(in_ch, out_ch, upsample)

    def forward(self, x):
        out = self.leaky_relu2(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return out + self.upsample(x)


class DepthConv(nn.Module):                 # layers.py:113-141
    def __init__(self, in_ch, out_ch, depth_kernel=3, stride=1, slope=0.01, inplace=False):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(in_ch, in_ch, 1, stride=stride), nn.LeakyReLU(negative_slope=slope, inplace=inplace))
        self.depth_conv = nn.Conv2d(in_ch, in_ch, depth_kernel, padding=depth_kernel // 2, groups=in_ch)
        self.conv2 = nn.Conv2d(in_ch, out_ch, 1)
        self.adaptor = None
        if stride != 1:
            assert stride == 2
            self.adaptor = nn.Conv2d(in_ch, out_ch, 2, stride=2)
        elif in_ch != out_ch:
            self.adaptor = nn.Conv2d(in_ch, out_ch, 1)

    def forward(self, x):
        identity = self.adaptor(x) if self.adaptor is not None else x
        return self.conv2(self.depth_conv(self.conv1(x))) + identity


class ConvFFN(nn.Module):                   # layers.py:144-157
    def __init__(self, in_ch, slope=0.1, inplace=False):
        super().__init__()
        mid = max(min(in_ch * 4, 1024), in_ch * 2)
        self.conv = nn.Sequential(nn.Conv2d(in_ch, mid, 1), nn.LeakyReLU(negative_slope=slope, inplace=inplace), nn.Conv2d(mid, in_ch, 1),
                                  nn.LeakyReLU(negative_slope=slope, inplace=inplace))

    def forward(self, x):
        return x + self.conv(x)


class ConvFFN3(nn.Module):                  # layers.py:159-172
    def __init__(self, in_ch, inplace=False):
        super().__init__()
        mid = in_ch * 2
        self.conv = nn.Conv2d(in_ch, mid * 2, 1)
        self.conv_out = nn.Conv2d(mid, in_ch, 1)
        self.relu1 = nn.LeakyReLU(negative_slope=0.1, inplace=inplace)
        self.relu2 = nn.LeakyReLU(negative_slope=0.01, inplace=inplace)

    def forward(self, x):
        a, b = self.conv(x).chunk(2, 1)
        return x + self.conv_out(self.relu1(a) + self.relu2(b))


class DepthConvBlock(nn.Module):            # layers.py:175-185
    def __init__(self, in_ch, out_ch, depth_kernel=3, stride=1, slope_depth_conv=0.01, slope_ffn=0.1, inplace=False):
        super().__init__()
        self.block = nn.Sequential(DepthConv(in_ch, out_ch, depth_kernel, stride, slope=slope_depth_conv, inplace=inplace),
                                   ConvFFN(out_ch, slope=slope_ffn, inplace=inplace))

    def forward(self, x):
        return self.block(x)


class DepthConvBlock4(nn.Module):           # layers.py:188-192
    def __init__(self, in_ch, out_ch, slope_depth_conv=0.01, inplace=False):
        super().__init__()
        self.block = nn.Sequential(DepthConv(in_ch, out_ch, slope=slope_depth_conv, inplace=inplace), ConvFFN3(out_ch, inplace=inplace))

    def forward(self, x):
        return self.block(x)
