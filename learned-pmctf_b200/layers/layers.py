"""Straight-through round / clamp and the 3x3 conv factory (reference: pMCTF/layers/layers.py:54-56,71-92).

These stay plain torch: they are parameter containers / autograd glue, not part of the GPU hot path
(the kernels apply rint / clamp themselves in their epilogues)."""
import torch
import torch.nn as nn


class RoundNoGradient(torch.autograd.Function):
    """y = round(x) (half to even), dy/dx = 1.  layers.py:71-80"""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, grad):
        return grad


class ClampNoGradient(torch.autograd.Function):
    """y = clamp(x, lo, hi), dy/dx = 1.  layers.py:83-92"""

    @staticmethod
    def forward(ctx, x, lo, hi):
        return torch.clamp(x, min=lo, max=hi)

    @staticmethod
    def backward(ctx, grad):
        return grad.clone(), None, None


def conv3x3(in_ch: int, out_ch: int, stride: int = 1, padding_mode: str = "zeros") -> nn.Conv2d:
    """Holds the weights of one 3x3 'same' convolution (layers.py:54-56)."""
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1, padding_mode=padding_mode)
