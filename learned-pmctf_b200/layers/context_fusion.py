"""The autoregressive entropy-parameter network of the LL subband (reference: pMCTF/layers/context_fusion.py:9-204 with
MaskedConv2d of pMCTF/layers/layers.py:23-52), as pWave++ instantiates it (pWave.py:80-82: 128 features, 2 parameters, no
context input).

The LL band of a 1080p plane is 72 x 120 coefficients (1/256 of the plane): this module is host-side torch code on stock
convolutions, with the reference's module tree, parameter AND buffer names (`mask` buffers are part of its state_dicts).  The full-
plane `forward` is what the training / evaluation pass uses; `forward_sequential` is the pixel-by-pixel form the bitstream path
needs (every coefficient's parameters depend on the coefficients decoded before it); on CUDA tensors that form runs as ONE kernel
per coefficient (decoder) or per band (encoder) instead of the reference's ~25 ATen calls per coefficient: `ar_encode` / `ar_decode`
(csrc/pmctf_llar.cu)."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from .. import ops


class MaskedConv2d(nn.Conv2d):
    """PixelCNN-style causal convolution: type 'A' hides the centre tap and everything after it in raster order, 'B' keeps the
    centre.  The mask is applied to the weights in place at every call, as the reference does (layers.py:49-51)."""

    def __init__(self, *args, mask_type: str = "A", **kwargs):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        mask = torch.ones_like(self.weight.data)
        kh, kw = mask.shape[-2:]
        mask[:, :, kh // 2, kw // 2 + (1 if mask_type == "B" else 0):] = 0
        mask[:, :, kh // 2 + 1:] = 0
        self.register_buffer("mask", mask)

    def forward(self, x, onehot=None):
        self.weight.data *= self.mask
        return super().forward(x)


def _masked3x3(cin, cout, mask_type):
    return MaskedConv2d(cin, cout, kernel_size=(3, 3), stride=(1, 1), mask_type=mask_type, padding=(1, 1))


class MaskResidual(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.conv1 = _masked3x3(num_features, num_features, "B")
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = _masked3x3(num_features, num_features, "B")
        self.conv1_input = self.conv2_input = None
        self.maskedWeight1 = self.maskedWeight2 = None

    def forward(self, x):
        return self.conv2(self.lrelu(self.conv1(x))) + x

    def init_sequential(self, like):
        self.conv1_input, self.conv2_input = torch.zeros_like(like), torch.zeros_like(like)
        self.maskedWeight1 = self.conv1.weight * self.conv1.mask
        self.maskedWeight2 = self.conv2.weight * self.conv2.mask

    def forward_sequential(self, x, h, w, kernel_size, padding):
        """one pixel: x [N,F,1,1] is written into the padded history of conv1's input at (h, w), the 3x3 window around it gives
        conv1's output there, and the same again for conv2 (context_fusion.py:30-41)"""
        self.conv1_input[:, :, h + padding:h + padding + 1, w + padding:w + padding + 1] = x
        t = F.conv2d(self.conv1_input[:, :, h:h + kernel_size, w:w + kernel_size], self.maskedWeight1, bias=self.conv1.bias)
        t = self.lrelu(t)
        self.conv2_input[:, :, h + padding:h + padding + 1, w + padding:w + padding + 1] = t
        t = F.conv2d(self.conv2_input[:, :, h:h + kernel_size, w:w + kernel_size], self.maskedWeight2, bias=self.conv2.bias)
        return t + x


class ContextFusionSubband(nn.Module):
    def __init__(self, in_channels=1, ctx_channels=1, num_features=128, context=False, num_parameters=9, prev_ctxs=False,
                 pretrained=None, ll_ctx=False, lower_subband=True, adaptive_quant=False):
        super().__init__()
        if context:
            raise NotImplementedError("pWave++ builds the LL model without a context input (pWave.py:80-82)")
        self.sequential_init = False
        self.residual_blocks = 2
        self.num_features, self.context, self.num_parameters, self.ctx_channels = num_features, context, num_parameters, ctx_channels
        self.maskedConv1 = _masked3x3(in_channels, num_features, "A")
        self.residualBlocks = nn.ModuleList(MaskResidual(num_features) for _ in range(self.residual_blocks))
        self.maskedConv2 = _masked3x3(num_features, num_features, "B")
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)
        self.convs = nn.ModuleList([nn.Conv2d(num_features, num_features, 1), nn.Conv2d(num_features, num_features, 1),
                                    nn.Conv2d(num_features, num_parameters, 1)])
        self.maskedWeight1 = self.maskedWeight2 = self.maskedConv2_input = self.mask1 = None

    def _tail(self, x):
        for i, c in enumerate(self.convs):
            x = c(x)
            if i < len(self.convs) - 1:
                x = self.lrelu(x)
        return x

    def forward(self, x, context=None, prev_subband=None, channel_idx=0):
        if x.is_cuda and not torch.is_grad_enabled() and x.size(1) == 1 and self.num_features == 128 and self.num_parameters == 2:
            return self.ar_forward(x)     # evaluation: the whole band on the kernels of csrc/pmctf_llar.cu, bit-equal to the bitstream path
        first = self.maskedConv1(x)
        x = first
        for blk in self.residualBlocks:
            x = blk(x)
        x = self.lrelu(self.maskedConv2(x + first))
        return self._tail(x)

    def init_sequential(self, y_hat):
        self.maskedWeight1 = self.maskedConv1.weight * self.maskedConv1.mask
        size = list(y_hat.size())
        size[1] = self.num_features
        self.mask1 = self.maskedConv1.mask[0].unsqueeze(0).repeat(size[0], 1, 1, 1)
        self.maskedConv2_input = torch.zeros(size, dtype=y_hat.dtype, device=y_hat.device)
        self.maskedWeight2 = self.maskedConv2.weight * self.maskedConv2.mask
        for blk in self.residualBlocks:
            blk.init_sequential(self.maskedConv2_input)
        self.sequential_init = True

    def forward_sequential(self, y_hat, h, w, context_list=None, channel_idx=0):
        """parameters of coefficient (h, w) from the zero-padded plane of what has been coded so far (context_fusion.py:160-204)"""
        if not self.sequential_init:
            self.init_sequential(y_hat)
        k, pad = 3, 1
        crop = y_hat[:, :, h:h + k, w:w + k].detach().clone()
        crop[self.mask1 == 0] = 0       # non-causal positions must not leak in even multiplied by a zero weight (-0, NaN)
        t = F.conv2d(crop, self.maskedWeight1, bias=self.maskedConv1.bias)
        first = t
        for blk in self.residualBlocks:
            t = blk.forward_sequential(t, h, w, k, pad)
        t = t + first
        self.maskedConv2_input[:, :, h + pad:h + pad + 1, w + pad:w + pad + 1] = t
        t = F.conv2d(self.maskedConv2_input[:, :, h:h + k, w:w + k], self.maskedWeight2, bias=self.maskedConv2.bias)
        return self._tail(self.lrelu(t))


    # ---- the sequential form on the GPU (csrc/pmctf_llar.cu) -------------------------------------------------------------------
    def _ar_weights(self):
        convs = [self.maskedConv1, self.residualBlocks[0].conv1, self.residualBlocks[0].conv2, self.residualBlocks[1].conv1,
                 self.residualBlocks[1].conv2, self.maskedConv2, self.convs[0], self.convs[1], self.convs[2]]
        key = tuple((c.weight.data_ptr(), c.weight._version, c.bias.data_ptr(), c.bias._version) for c in convs)
        cache = self.__dict__.get("_ar_cache")
        if cache is None or cache[0] != key:
            if self.num_features != 128 or self.num_parameters != 2 or self.maskedConv1.weight.shape[1] != 1:
                raise RuntimeError("the sequential LL kernel is built for pWave++'s configuration (1 -> 128 features -> 2 parameters)")
            lib, dev = nat.lib(), self.maskedConv1.weight.device
            sizes = [4 * 128] + [5 * 128 * 128] * 5 + [128 * 128] * 2
            buf = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            offs, off = [], 0
            for c, sz, (cin, taps) in zip(convs[:8], sizes, [(1, 4)] + [(128, 5)] * 5 + [(128, 0)] * 2):
                ops._launch(dev, "llar_pack", lib.pmctf_llar_pack, c.weight.detach().contiguous().data_ptr(), cin, taps, buf.data_ptr() + 4 * off)
                offs.append(off)
                off += sz
            keep = [c.bias.detach().contiguous() for c in convs] + [convs[8].weight.detach().contiguous()]
            cache = (key, buf, offs, keep)
            self.__dict__["_ar_cache"] = cache
        return cache

    def _ar_desc(self, B, H, W, dev):
        _, buf, offs, keep = self._ar_weights()
        d = nat.LLar()
        base = buf.data_ptr()
        d.w_in, d.b_in = base + 4 * offs[0], keep[0].data_ptr()
        for i in range(5):
            d.w[i], d.b[i] = base + 4 * offs[1 + i], keep[1 + i].data_ptr()
        for i in range(2):
            d.w1[i], d.b1[i] = base + 4 * offs[6 + i], keep[6 + i].data_ptr()
        d.w_out, d.b_out = keep[9].data_ptr(), keep[8].data_ptr()
        Y = torch.zeros((B, H + 2, W + 2), dtype=torch.float32, device=dev)
        hist = torch.zeros((5, B, H + 2, W + 2, 128), dtype=torch.float32, device=dev)
        d.Y = Y.data_ptr()
        for i in range(5):
            d.hist[i] = hist[i].data_ptr()
        d.B, d.H, d.W = B, H, W
        d.log_scale_min = float(np.float32(math.log(0.01)))                       # GaussianEncoder('laplace'), entropy_models.py:204-221
        d.log_scale_step = float(np.float32((math.log(64.0) - math.log(0.01)) / 255))
        d.scale_levels = 256
        return d, (Y, hist)

    def ar_forward(self, x):
        """forward (context_fusion.py:143-158) of a whole band x [B,1,H,W] on the layer-parallel kernels: -> [B,2,H,W] (scales,
        means), each value computed with the arithmetic of the sequential form, i.e. exactly the parameters the bitstream path
        codes with when the band is its own reconstruction."""
        x = ops._chk(x, "ll", 4).contiguous()
        B, _, H, W = x.shape
        d, _keep = self._ar_desc(B, H, W, x.device)
        out = torch.empty((2, B, H, W), dtype=torch.float32, device=x.device)
        ops._launch(x.device, "llar_forward", nat.lib().pmctf_llar_forward, C.byref(d), x.data_ptr(), 0, None, None, out[0].data_ptr(),
                    out[1].data_ptr(), None)
        return out.permute(1, 0, 2, 3).contiguous()

    # PMCTF_LL_SEQUENTIAL=1: always run the coefficient-by-coefficient encoder (A/B runs, tests)
    def ar_encode(self, yq, parallel=None):
        """yq [B,1,H,W] quantised LL band (CUDA) -> (ll_hat [B,1,H,W] as the decoder will reconstruct it, int16 symbols, int16 table
        indexes), the latter two as numpy arrays in the order the entropy coder consumes them: coefficient by coefficient in raster
        order, planes of the batch innermost (one `encoder.encode` per coefficient in pWave.py:548-553)."""
        yq = ops._chk(yq, "ll", 4).contiguous()
        B, Cc, H, W = yq.shape
        if Cc != 1:
            raise RuntimeError("the LL band is single-channel")
        d, (Y, _hist) = self._ar_desc(B, H, W, yq.device)
        sym = torch.empty((B, H * W), dtype=torch.int16, device=yq.device)
        idx = torch.empty_like(sym)
        if parallel is None:
            parallel = os.environ.get("PMCTF_LL_SEQUENTIAL", "0") != "1"
        self.last_encode_path = "sequential"
        if parallel:
            # the encoder knows every coefficient: all of them at once, the history speculated to be round(y); a coefficient whose
            # reconstruction round(symbol + mean) differs from that (round(y) - mean ending in exactly .5) sends the band through the
            # sequential kernel below, which conditions on the reconstruction by construction
            flag = torch.zeros(1, dtype=torch.int32, device=yq.device)
            ops._launch(yq.device, "llar_forward", nat.lib().pmctf_llar_forward, C.byref(d), yq.data_ptr(), 1, sym.data_ptr(), idx.data_ptr(),
                        None, None, flag.data_ptr())
            if int(flag.item()) == 0:
                self.last_encode_path = "parallel"
        if self.last_encode_path == "sequential":
            ops._launch(yq.device, "llar_encode", nat.lib().pmctf_llar_encode, C.byref(d), yq.data_ptr(), sym.data_ptr(), idx.data_ptr())
        ll_hat = Y[:, 1:-1, 1:-1].unsqueeze(1).contiguous()
        return ll_hat, sym.t().contiguous().cpu().numpy().reshape(-1), idx.t().contiguous().cpu().numpy().reshape(-1)

    def ar_decode_band(self, size, rans_decoder, cdf, cdf_sizes, offsets, device):
        """The whole band in ONE launch (pmctf_llar_decode_band): a cluster of eight CTAs per plane with the weights resident in
        shared memory, the band's symbols decoded by the device-side rANS decoder from the sub-stream of `rans_decoder` (a
        models.MLCodec_rans.RansDecoder) that holds them, whose reader is moved past the band afterwards.  -> [B,1,H,W], or None
        when the configuration is outside the kernel's (a coefficient's symbols spread over several sub-streams, more than 16 planes)."""
        B, Cc, H, W = size
        lib, device = nat.lib(), torch.device(device)
        h = getattr(rans_decoder, "_h", None)
        if Cc != 1 or B > 16 or h is None or os.environ.get("PMCTF_LL_SEQUENTIAL", "0") == "1":
            return None
        parts = lib.pmctf_rans_decoder_parts(h)
        if parts != 1 and B >= parts:
            return None                   # a coefficient's B symbols would come from several sub-streams
        part = parts - 1                  # B < parts: every share but the last is empty (py_rans.cpp:45-59)
        x, pos, nwords, words = C.c_ulonglong(), C.c_longlong(), C.c_longlong(), C.c_void_p()
        nat.check(lib.pmctf_rans_decoder_peek(h, part, C.byref(x), C.byref(pos), C.byref(nwords), C.byref(words)), "rans_decoder_peek")
        wnp = np.ctypeslib.as_array(C.cast(words, C.POINTER(C.c_uint32)), shape=(max(int(nwords.value), 1),))[: int(nwords.value)]
        key = (id(cdf), id(cdf_sizes), id(offsets), str(device))
        tabs = self.__dict__.get("_ar_tables")
        if tabs is None or tabs[0] != key:
            from ..models.MLCodec_rans import _tables
            cd, sz, of = _tables(cdf, cdf_sizes, offsets)
            tabs = (key, torch.from_numpy(cd.copy()).to(device), torch.from_numpy(sz.copy()).to(device), torch.from_numpy(of.copy()).to(device),
                    (cdf, cdf_sizes, offsets))
            self.__dict__["_ar_tables"] = tabs
        _, cd, sz, of, _keep = tabs
        wdev = torch.from_numpy(wnp.view(np.int32).copy()).to(device) if wnp.size else torch.zeros(1, dtype=torch.int32, device=device)
        state = torch.tensor([x.value if x.value < (1 << 63) else x.value - (1 << 64), pos.value, 0, 0], dtype=torch.int64, device=device)
        d, _keepbuf = self._ar_desc(B, H, W, device)
        out = torch.empty((B, H * W), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = lib.pmctf_llar_decode_band(C.byref(d), wdev.data_ptr(), int(nwords.value), state.data_ptr(), cd.data_ptr(), cd.shape[0],
                                            cd.shape[1], sz.data_ptr(), of.data_ptr(), out.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        if rc != 0:       # the cluster could not be launched here (e.g. no eight free SMs with 222 KB of shared memory in one GPC):
            return None   # nothing was decoded, the reader has not moved -- the caller takes the per-coefficient path
        st = state.cpu().tolist()
        if st[2] != 0:
            raise RuntimeError("LL band: the entropy-coded stream ended early or is corrupt")
        nat.check(lib.pmctf_rans_decoder_seek(h, part, C.c_ulonglong(st[0] & ((1 << 64) - 1)), st[1]), "rans_decoder_seek")
        return out.view(B, 1, H, W)

    def ar_decode(self, size, decode, device):
        """size = [B,1,H,W]; decode(idx int16 numpy [B]) -> int16 numpy [B] symbols of one coefficient.  One kernel launch + one
        stream synchronisation + one rANS step per coefficient (the fallback of ar_decode_band)."""
        B, Cc, H, W = size
        if Cc != 1:
            raise RuntimeError("the LL band is single-channel")
        device = torch.device(device)
        d, _keep = self._ar_desc(B, H, W, device)
        prev = torch.zeros(B, dtype=torch.float32, pin_memory=True)
        mean = torch.zeros(B, dtype=torch.float32, pin_memory=True)
        idx = torch.zeros(B, dtype=torch.int16, pin_memory=True)
        out = np.zeros((B, H * W), dtype=np.float32)
        lib, st = nat.lib(), torch.cuda.current_stream(device)
        prev_np, mean_np, idx_np = prev.numpy(), mean.numpy(), idx.numpy()
        with torch.cuda.device(device):
            for pos in range(H * W):
                nat.check(lib.pmctf_llar_decode_step(C.byref(d), pos, prev.data_ptr(), mean.data_ptr(), idx.data_ptr(), st.cuda_stream), "llar_decode_step")
                st.synchronize()
                sym = decode(idx_np.copy()).astype(np.float32)
                prev_np[:] = np.rint(sym + mean_np)                             # round(decoded + mean), fp32 like the kernel's encoder side
                out[:, pos] = prev_np
        return torch.from_numpy(out.reshape(B, 1, H, W)).to(device)
