"""PostProcess, the de-quantisation filter of pWave++ (reference: pMCTF/layers/postprocessing.py:7-44), on the B200 tensor cores.

Same module tree and parameter names as the reference (conv1, resBlocks.{0..5}.conv1/conv2, conv2, conv3), so the
`dequantModule.*` entries of its state_dicts load unchanged.  forward(x) = x + conv3(conv2(ResBlocks(conv1(x))) + conv1(x)); the
thirteen 64 -> 64 convolutions (958 kFLOP per pixel) and the last layer run as tcgen05 implicit GEMMs with bf16 operands and fp32
accumulation (csrc/pmctf_pp.cu), everything else in fp32.  CUDA tensors only; under autograd the module falls back to composing
the same formula from torch ops (training of this filter is not part of the hot path)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from .. import ops


class ResBlock(nn.Module):
    def __init__(self, intermediate_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(intermediate_channels, intermediate_channels, 3, padding=1)
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv2d(intermediate_channels, intermediate_channels, 3, padding=1)


class PostProcess(nn.Module):
    def __init__(self, num_res=6, intermediate_channels=64, in_channels=1, out_channels=1):
        super().__init__()
        if (num_res, intermediate_channels, in_channels, out_channels) != (6, 64, 1, 1):
            raise NotImplementedError("the B200 PostProcess kernel is built for the configuration pWave++ uses (6 ResBlocks, 64 channels, "
                                      "1 -> 1: pWave.py:62)")
        self.num_res = num_res
        self.resBlocks = nn.ModuleList(ResBlock(intermediate_channels) for _ in range(num_res))
        self.conv1 = nn.Conv2d(in_channels, intermediate_channels, 3, padding=1)
        self.conv2 = nn.Conv2d(intermediate_channels, intermediate_channels, 3, padding=1)
        self.conv3 = nn.Conv2d(intermediate_channels, out_channels, 3, padding=1)
        self._key = None
        self._packed = None
        self._desc = None

    # caches hold device pointers
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_key"], st["_packed"], st["_desc"] = None, None, None
        return st

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._key = None

    def invalidate_packed(self):
        self._key = None

    def _layers(self):
        """the 14 tensor-core layers in execution order: 12 ResBlock convs, conv2, conv3"""
        return [c for b in self.resBlocks for c in (b.conv1, b.conv2)] + [self.conv2, self.conv3]

    def descriptor(self) -> nat.PostProcessD:
        params = [self.conv1.weight, self.conv1.bias] + [p for c in self._layers() for p in (c.weight, c.bias)]
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        if key != self._key:
            dev = self.conv1.weight.device
            lib = nat.lib()
            n64, n1 = int(lib.pmctf_pp_packed_bytes(64)), int(lib.pmctf_pp_packed_bytes(1))
            packed = torch.empty(13 * n64 + n1, dtype=torch.uint8, device=dev)
            keep = [self.conv1.weight.detach().contiguous(), self.conv1.bias.detach().contiguous()]
            d = nat.PostProcessD()
            d.conv1_w, d.conv1_b = keep[0].data_ptr(), keep[1].data_ptr()
            off = 0
            for i, c in enumerate(self._layers()):
                w, b = ops._chk(c.weight.detach().contiguous(), "conv weight"), ops._chk(c.bias.detach().contiguous(), "conv bias")
                co = w.size(0)
                if tuple(w.shape) != (co, 64, 3, 3) or co not in (64, 1):
                    raise RuntimeError(f"PostProcess layer {i}: weight {tuple(w.shape)} (postprocessing.py:10-33 expects [64|1, 64, 3, 3])")
                ops._launch(dev, "pp_pack_conv", lib.pmctf_pp_pack_conv, w.data_ptr(), co, packed.data_ptr() + off)
                keep += [w, b]
                if i < 12:
                    d.res_w[i], d.res_b[i] = packed.data_ptr() + off, b.data_ptr()
                elif i == 12:
                    d.conv2_w, d.conv2_b = packed.data_ptr() + off, b.data_ptr()
                else:
                    d.conv3_w, d.conv3_b = packed.data_ptr() + off, b.data_ptr()
                off += n64 if co == 64 else n1
            self._key, self._packed, self._desc = key, (packed, keep), d
        return self._desc

    def forward(self, x, in_mul: float = 1.0, out_mul: float = 1.0):
        """x [N,1,H,W] -> x + correction (postprocessing.py:35-44).  in_mul / out_mul fuse the scaling pWave wraps around the
        call (dequantModule(x_hat / 256) * 256, pWave.py:300): forward(x_hat, 1/256, 256)."""
        from .. import train
        if train.needs_grad(x, self):
            xs = x * in_mul
            c1 = self.conv1(xs)
            t = c1
            for b in self.resBlocks:
                t = b.conv2(F.leaky_relu(b.conv1(t), 0.2)) + t
            return (xs + self.conv3(self.conv2(t) + c1)) * out_mul
        return ops.postprocess(x, self.descriptor(), in_mul, out_mul)
