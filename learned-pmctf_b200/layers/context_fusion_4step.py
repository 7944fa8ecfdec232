"""ContextFusionFourStep, the four-step entropy-parameter network of pWave++ (reference: pMCTF/layers/context_fusion_4step.py:9-249),
on the B200 tensor cores.

Same module tree and parameter names as the reference (y_hierarchical_prior_enc.{0,1}.conv{1,2}, conv1_context,
lower_level_subband.1, y_hierarchical_prior_out.block.{0,1}.*, y_spatial_prior_{1,2,3}.{0,1.conv1,1.conv2},
y_spatial_prior_{1,2,3}_out.{0,1}.conv{1,2} / .2), so the `context_fusion.{lvl}.{lh,hl,hh}.*` entries of its state_dicts load
unchanged.  forward / compress / decompress keep the reference's signatures and return values.

The 22 dense 112 -> 112 3x3 convolutions and the 1x1 of the DepthConvBlock head (4.97 MFLOP per coefficient) run as tcgen05
CTA-pair implicit GEMMs (csrc/pmctf_ctx.cu); the 1|2 -> 112 input convolutions, the depthwise head, the 112 -> 2 projections
and the masked quantiser are small CUDA-core kernels.  CUDA tensors only; under autograd the module composes the reference's
formula from torch ops (training of the entropy model is not part of the hot path)."""
from __future__ import annotations

import collections
import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from .. import ops
from .layers import RoundNoGradient

NUM_FEATURES = 112


class ContextResidual(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.conv1 = nn.Conv2d(num_features, num_features, 3, padding=1)
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)
        self.conv2 = nn.Conv2d(num_features, num_features, 3, padding=1)

    def forward(self, x):
        return self.conv2(self.lrelu(self.conv1(x))) + x


class _DepthConv(nn.Module):   # pMCTF/layers/video/layers.py:113-141 (stride 1, in_ch != out_ch)
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(in_ch, in_ch, 1), nn.LeakyReLU(negative_slope=0.01))
        self.depth_conv = nn.Conv2d(in_ch, in_ch, 3, padding=1, groups=in_ch)
        self.conv2 = nn.Conv2d(in_ch, out_ch, 1)
        self.adaptor = nn.Conv2d(in_ch, out_ch, 1)

    def forward(self, x):
        return self.conv2(self.depth_conv(self.conv1(x))) + self.adaptor(x)


class _ConvFFN(nn.Module):     # layers.py:144-157
    def __init__(self, in_ch):
        super().__init__()
        mid = max(min(in_ch * 4, 1024), in_ch * 2)
        self.conv = nn.Sequential(nn.Conv2d(in_ch, mid, 1), nn.LeakyReLU(negative_slope=0.1), nn.Conv2d(mid, in_ch, 1),
                                  nn.LeakyReLU(negative_slope=0.1))

    def forward(self, x):
        return x + self.conv(x)


class DepthConvBlock(nn.Module):   # layers.py:160-172
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.block = nn.Sequential(_DepthConv(in_ch, out_ch), _ConvFFN(out_ch))

    def forward(self, x):
        return self.block(x)


class _Features:
    """A 112-channel feature map in the two chunk-planar layouts the kernels exchange."""

    def __init__(self, N, H, W, device, f32=True, bf16=True):
        self.N, self.H, self.W = N, H, W
        self.f32 = torch.empty((N, 28, H, W, 4), dtype=torch.float32, device=device) if f32 else None
        self.bf16 = torch.empty((N, 14, H, W, 8), dtype=torch.bfloat16, device=device) if bf16 else None

    def nchw(self):   # [N,112,H,W] fp32 view for tests
        return self.f32.permute(0, 1, 4, 2, 3).reshape(self.N, NUM_FEATURES, self.H, self.W)


# The three full feature maps + the bf16-only intermediate of one call, shared by ALL four-step modules of the process (the 24 modules
# of a codec run one after the other on one stream, so stream-ordered reuse is safe) and kept per shape, least recently used shapes
# dropped beyond a byte budget (PMCTF_CTX_WORKSPACE_GB, default 24).
_WS_CACHE = collections.OrderedDict()


def _shared_workspace(N, H, W, dev):
    key = (N, H, W, str(dev), torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is not None:
        _WS_CACHE.move_to_end(key)
        return ws
    need = N * H * W * NUM_FEATURES * (3 * 6 + 2)
    budget = float(os.environ.get("PMCTF_CTX_WORKSPACE_GB", "24")) * 2 ** 30
    while _WS_CACHE and sum(v[4] for v in _WS_CACHE.values()) + need > budget:
        _WS_CACHE.popitem(last=False)
    ws = (_Features(N, H, W, dev), _Features(N, H, W, dev), _Features(N, H, W, dev), _Features(N, H, W, dev, f32=False), need)
    _WS_CACHE[key] = ws
    return ws


def release_workspaces():
    _WS_CACHE.clear()


_GRAPHS_ON = os.environ.get("PMCTF_CTX_GRAPHS", "1") != "0"


class _GraphUnavailable(Exception):
    pass


class ContextFusionFourStep(nn.Module):
    def __init__(self, in_channels=1, ctx_channels=1, num_features=NUM_FEATURES, num_parameters=2, ctx=True, lossy=True,
                 lower_subband=True):
        super().__init__()
        if (in_channels, num_features, num_parameters, ctx) != (1, NUM_FEATURES, 2, True) or ctx_channels not in (1, 2):
            raise NotImplementedError("the B200 four-step kernels are built for the configuration pWave++ uses (in_channels 1, "
                                      "112 features, 2 parameters, context on: pWave.py:70-78)")
        self.num_ch, self.num_parameters, self.ctx_channels, self.ctx, self.lossy = num_features, num_parameters, ctx_channels, ctx, lossy
        self.masks = {}
        self.y_hierarchical_prior_enc = nn.Sequential(ContextResidual(num_features), ContextResidual(num_features))
        self.conv1_context = nn.Conv2d(ctx_channels, num_features, 3, padding=1)
        if ctx_channels > 1 and lower_subband:
            self.lower_level_subband = nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(1, 1, 3, padding=1))
        self.y_hierarchical_prior_out = DepthConvBlock(num_features, num_parameters)
        for k in (1, 2, 3):
            setattr(self, f"y_spatial_prior_{k}", nn.Sequential(nn.Conv2d(1, num_features, 3, padding=1), ContextResidual(num_features)))
            setattr(self, f"y_spatial_prior_{k}_out", nn.Sequential(ContextResidual(num_features), ContextResidual(num_features),
                                                                    nn.Conv2d(num_features, num_parameters, 1)))
        self._key = None
        self._packed = None
        # scale-table constants of GaussianEncoder('laplace') (entropy_models.py:204-221), for the fused symbol staging
        self.log_scale_min = math.log(0.01)
        self.log_scale_step = (math.log(64.0) - math.log(0.01)) / 255
        self.scale_levels = 256

    # ---- packed tensor-core operands (one bf16 image per 112 -> 112 layer), rebuilt when a parameter changes ---------------
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_key"], st["_packed"] = None, None
        st.pop("_pack_hot", None)
        st.pop("_graphs", None)
        st.pop("_ws_hot", None)
        return st

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._key = None

    def invalidate_packed(self):
        self._key = None

    def _tc_layers(self):
        out = []
        for blk in self.y_hierarchical_prior_enc:
            out += [blk.conv1, blk.conv2]
        out.append(self.y_hierarchical_prior_out.block[0].conv1[0])
        for k in (1, 2, 3):
            sp, so = getattr(self, f"y_spatial_prior_{k}"), getattr(self, f"y_spatial_prior_{k}_out")
            out += [sp[1].conv1, sp[1].conv2, so[0].conv1, so[0].conv2, so[1].conv1, so[1].conv2]
        return out

    def _pack(self):
        hot = self.__dict__.get("_pack_hot")
        if hot is not None:                  # inside one forward pass the operands were checked already
            return hot
        layers = self._tc_layers()
        key = tuple((c.weight.data_ptr(), c.weight._version) for c in layers)
        if key != self._key:
            lib = nat.lib()
            dev = layers[0].weight.device
            offs, total = {}, 0
            for c in layers:
                taps = c.weight.shape[-1] ** 2
                offs[id(c)] = (total, taps)
                total += int(lib.pmctf_ctx_packed_bytes(taps))
            buf = torch.empty(total, dtype=torch.uint8, device=dev)
            for c in layers:
                w = ops._chk(c.weight.detach().contiguous(), "conv weight")
                if tuple(w.shape[:2]) != (NUM_FEATURES, NUM_FEATURES):
                    raise RuntimeError(f"tensor-core layer with weight {tuple(w.shape)}")
                off, taps = offs[id(c)]
                ops._launch(dev, "ctx_pack_conv", lib.pmctf_ctx_pack_conv, w.data_ptr(), taps, buf.data_ptr() + off)
            self._key, self._packed = key, (buf, offs)
        return self._packed

    # ---- layer launchers ------------------------------------------------------------------------------------------------
    def _conv112(self, conv, src: _Features, dst: _Features, slope=1.0, res=None, res2=None):
        buf, offs = self._pack()
        off, taps = offs[id(conv)]
        ops._launch(src.bf16.device, "ctx_conv112", nat.lib().pmctf_ctx_conv112, src.bf16.data_ptr(), buf.data_ptr() + off, taps,
                    conv.bias.detach().data_ptr(), res.f32.data_ptr() if res is not None else None,
                    res2.f32.data_ptr() if res2 is not None else None, float(slope),
                    dst.f32.data_ptr() if dst.f32 is not None else None, dst.bf16.data_ptr() if dst.bf16 is not None else None,
                    src.N, src.H, src.W)
        return dst

    def _conv_in(self, conv, x0, x1, dst: _Features):
        ops._launch(x0.device, "ctx_conv_in", nat.lib().pmctf_ctx_conv_in, x0.data_ptr(), x1.data_ptr() if x1 is not None else None,
                    conv.weight.detach().data_ptr(), conv.bias.detach().data_ptr(), dst.f32.data_ptr(), dst.bf16.data_ptr(),
                    dst.N, dst.H, dst.W)
        return dst

    def _resblock(self, blk, src: _Features, tmp: _Features, dst: _Features, res2=None):
        """dst = conv2(lrelu_0.2(conv1(src))) + src (+ res2)"""
        self._conv112(blk.conv1, src, tmp, slope=0.2)
        return self._conv112(blk.conv2, tmp, dst, res=src, res2=res2)

    def _context_features(self, context, prev_subband, N, H, W):
        dev = context.device
        x1 = None
        if prev_subband is not None:
            if not hasattr(self, "lower_level_subband"):
                raise RuntimeError("prev_subband given to a module built with ctx_channels = 1")
            prev = ops._chk(prev_subband, "prev_subband", 4).contiguous()
            if tuple(prev.shape) != (N, 1, H // 2, W // 2) or H % 2 or W % 2:
                raise RuntimeError(f"prev_subband {tuple(prev.shape)} for a subband {(N, 1, H, W)}")
            x1 = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
            c = self.lower_level_subband[1]
            ops._launch(dev, "ctx_lower_subband", nat.lib().pmctf_ctx_lower_subband, prev.data_ptr(), c.weight.detach().data_ptr(),
                        c.bias.detach().data_ptr(), x1.data_ptr(), N, H // 2, W // 2)
        elif self.ctx_channels == 2:
            raise RuntimeError("this module expects prev_subband (ctx_channels = 2)")
        a, b, _, t, _ = self._workspace(N, H, W, dev)
        self._conv_in(self.conv1_context, context, x1, a)
        self._resblock(self.y_hierarchical_prior_enc[0], a, t, b)
        self._resblock(self.y_hierarchical_prior_enc[1], b, t, a)
        return a, b, t   # a = context features; b, t = scratch

    def _hierarchical(self, cf: _Features, scratch: _Features):
        blk = self.y_hierarchical_prior_out.block
        dc, ffn = blk[0], blk[1]
        N, H, W = cf.N, cf.H, cf.W
        t1 = _Features(N, H, W, cf.f32.device, bf16=False) if scratch.f32 is None else scratch
        keep_bf16, t1.bf16 = t1.bf16, None                       # fp32 output only
        self._conv112(dc.conv1[0], cf, t1, slope=0.01)
        t1.bf16 = keep_bf16
        d = nat.CtxDcb()
        g = lambda p: p.detach().data_ptr()  # noqa: E731
        d.dw_w, d.dw_b, d.pw_w, d.pw_b = g(dc.depth_conv.weight), g(dc.depth_conv.bias), g(dc.conv2.weight), g(dc.conv2.bias)
        d.ad_w, d.ad_b = g(dc.adaptor.weight), g(dc.adaptor.bias)
        d.f1_w, d.f1_b, d.f2_w, d.f2_b = g(ffn.conv[0].weight), g(ffn.conv[0].bias), g(ffn.conv[2].weight), g(ffn.conv[2].bias)
        scales = torch.empty((N, 1, H, W), dtype=torch.float32, device=cf.f32.device)
        means = torch.empty_like(scales)
        ops._launch(scales.device, "ctx_dcb_tail", nat.lib().pmctf_ctx_dcb_tail, t1.f32.data_ptr(), cf.f32.data_ptr(), C.byref(d),
                    scales.data_ptr(), means.data_ptr(), N, H, W)
        return scales, means

    def _spatial(self, k, x_hat, cf: _Features, a: _Features, b: _Features, t: _Features):
        sp, so = getattr(self, f"y_spatial_prior_{k}"), getattr(self, f"y_spatial_prior_{k}_out")
        self._conv_in(sp[0], x_hat, None, a)
        self._resblock(sp[1], a, t, b, res2=cf)                   # ... + context (:149)
        self._resblock(so[0], b, t, a)
        N, H, W = cf.N, cf.H, cf.W
        scales = torch.empty((N, 1, H, W), dtype=torch.float32, device=x_hat.device)
        means = torch.empty_like(scales)
        # the last block's second convolution evaluates the 112 -> 2 projection in its epilogue: its feature map is never written
        blk = so[1]
        self._conv112(blk.conv1, a, t, slope=0.2)
        buf, offs = self._pack()
        off, taps = offs[id(blk.conv2)]
        ops._launch(x_hat.device, "ctx_conv112_head", nat.lib().pmctf_ctx_conv112_head, t.bf16.data_ptr(), buf.data_ptr() + off, taps,
                    blk.conv2.bias.detach().data_ptr(), a.f32.data_ptr(), None, 1.0, so[2].weight.detach().data_ptr(),
                    so[2].bias.detach().data_ptr(), scales.data_ptr(), means.data_ptr(), N, H, W)
        return scales, means

    def _mask_step(self, k, x, dec_sym, scales, means, run, sym16=None, idx16=None):
        s = nat.CtxStep()
        s.x = x.data_ptr() if x is not None else None
        s.dec_sym = dec_sym.data_ptr() if dec_sym is not None else None
        s.scales, s.means = scales.data_ptr(), means.data_ptr()
        s.x_hat = run["x_hat"].data_ptr() if run.get("x_hat") is not None else None
        s.s_hat = run["s_hat"].data_ptr() if run.get("s_hat") is not None else None
        s.x_q = run["x_q"].data_ptr() if run.get("x_q") is not None else None
        s.x_res = run["x_res"].data_ptr() if run.get("x_res") is not None else None
        s.sym16 = sym16.data_ptr() if sym16 is not None else None
        s.idx16 = idx16.data_ptr() if idx16 is not None else None
        s.log_scale_min, s.log_scale_step = float(np.float32(self.log_scale_min)), float(np.float32(self.log_scale_step))
        s.scale_levels, s.step, s.lossy = self.scale_levels, k, int(bool(self.lossy))
        s.N, s.H, s.W = scales.shape[0], scales.shape[2], scales.shape[3]
        ops._launch(scales.device, "ctx_mask_step", nat.lib().pmctf_ctx_mask_step, C.byref(s))

    # ---- the reference's entry points ------------------------------------------------------------------------------------
    def get_mask_four_parts(self, height, width, dtype, device):
        key = f"{width}x{height}"
        if key not in self.masks:
            yy, xx = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
            code = 2 * (yy & 1) + (xx & 1)
            self.masks[key] = [(code == k).to(dtype)[None, None] for k in range(4)]
        return self.masks[key]

    def quant(self, x):
        return RoundNoGradient.apply(x) if self.training else torch.round(x)

    def process_with_mask(self, y, scales, means, mask):
        if not self.lossy:
            means = RoundNoGradient.apply(means)
        means_hat = means * mask
        y_res = (y - means_hat) * mask
        y_q = self.quant(y_res)
        return y_res, y_q, y_q + means_hat, scales * mask

    def _forward_torch(self, x, context, prev_subband, write):
        if prev_subband is not None:
            context = torch.cat((context, self.lower_level_subband(prev_subband)), dim=1)
        context = self.y_hierarchical_prior_enc(self.conv1_context(context))
        scales, means = self.y_hierarchical_prior_out(context).chunk(2, dim=1)
        masks = self.get_mask_four_parts(x.size(2), x.size(3), x.dtype, x.device)
        outs, so_far = [], None
        for k in range(4):
            if k > 0:
                t = getattr(self, f"y_spatial_prior_{k}")(so_far) + context
                scales, means = getattr(self, f"y_spatial_prior_{k}_out")(t).chunk(2, dim=1)
            o = self.process_with_mask(x, scales, means, masks[k])
            outs.append(o)
            so_far = o[2] if so_far is None else so_far + o[2]
        if write:
            return (*[o[1] for o in outs], *[o[3] for o in outs], so_far)
        return sum(o[0] for o in outs), sum(o[1] for o in outs), so_far, sum(o[3] for o in outs)

    def _run(self, x, context, prev_subband, stage=None, dec=None):
        """The four steps on the GPU.  Encoder: x given; decoder: dec(k, scales_masked_idx16) -> int16 symbols of step k.
        stage: optional list that receives (sym16, idx16) device tensors per step (compress).
        The encoder form (about 37 launches of our own kernels, none of them data dependent) is captured ONCE per (shape, weights,
        stream) into a CUDA graph and replayed: the module is launch-bound on the host otherwise (the whole pWave.forward spent
        48 ms enqueueing 32 ms of GPU work).  PMCTF_CTX_GRAPHS=0 keeps every launch eager."""
        if dec is None and x is not None and _GRAPHS_ON and x.is_cuda and not torch.cuda.is_current_stream_capturing():
            try:
                return self._run_graphed(x, context, prev_subband, stage)
            except _GraphUnavailable:
                pass
        return self._run_eager(x, context, prev_subband, stage, dec)

    def _run_eager(self, x, context, prev_subband, stage=None, dec=None):
        self.__dict__["_pack_hot"] = None
        self.__dict__["_pack_hot"] = self._pack()
        try:
            return self._run_steps(x, context, prev_subband, stage, dec)
        finally:
            self.__dict__["_pack_hot"] = None

    def _run_graphed(self, x, context, prev_subband, stage):
        x = ops._chk(x, "x", 4).contiguous()
        context = ops._chk(context, "context", 4).contiguous()
        if prev_subband is not None:
            prev_subband = ops._chk(prev_subband, "prev_subband", 4).contiguous()
        dev = x.device
        stream = torch.cuda.current_stream(dev)
        _, offs = self._pack()
        key = (tuple(x.shape), prev_subband is not None, stage is not None, str(dev), stream.cuda_stream, self._key)
        graphs = self.__dict__.setdefault("_graphs", collections.OrderedDict())
        ent = graphs.get(key)
        if ent is None:
            if self.__dict__.get("_graph_failed"):
                raise _GraphUnavailable()
            N, _, H, W = x.shape
            self._run_eager(x, context, prev_subband, [] if stage is not None else None)     # warm: kernel attributes, packed weights
            ws = _shared_workspace(N, H, W, dev)                                              # shared, allocated outside the capture
            sx, sc = x.clone(), context.clone()
            sp = prev_subband.clone() if prev_subband is not None else None
            st = [] if stage is not None else None
            g = torch.cuda.CUDAGraph()
            self.__dict__["_ws_hot"] = ws
            try:
                with torch.cuda.graph(g):
                    run, _ = self._run_eager(sx, sc, sp, st)
            except Exception:
                self.__dict__["_graph_failed"] = True
                raise _GraphUnavailable()
            finally:
                self.__dict__["_ws_hot"] = None
            while len(graphs) >= 12:
                graphs.popitem(last=False)
            ent = graphs[key] = (g, sx, sc, sp, run, st, ws)
        else:
            graphs.move_to_end(key)
        g, sx, sc, sp, run, st, _ws = ent
        sx.copy_(x)
        sc.copy_(context)
        if sp is not None:
            sp.copy_(prev_subband)
        g.replay()
        out = {k: v.clone() for k, v in run.items()}      # the graph owns its outputs: hand copies to the caller
        if stage is not None:
            stage.extend((a.clone(), b.clone()) for a, b in st)
        return out, None

    def _workspace(self, N, H, W, dev):
        hot = self.__dict__.get("_ws_hot")
        if hot is not None:
            return hot
        return _shared_workspace(N, H, W, dev)

    def _run_steps(self, x, context, prev_subband, stage=None, dec=None):
        ref_t = x if x is not None else context
        ctx_t = ops._chk(context, "context", 4).contiguous()
        N, _, H, W = ctx_t.shape
        if x is not None:
            x = ops._chk(x, "x", 4).contiguous()
            if tuple(x.shape) != (N, 1, H, W):
                raise RuntimeError(f"x {tuple(x.shape)} vs context {tuple(ctx_t.shape)}")
        dev = ref_t.device
        cf, b, t = self._context_features(ctx_t, prev_subband, N, H, W)
        a = self._workspace(N, H, W, dev)[2]
        scales, means = self._hierarchical(cf, a)
        run = {k: torch.empty((N, 1, H, W), dtype=torch.float32, device=dev) for k in ("x_hat", "x_q", "s_hat", "x_res")}
        steps = []
        for k in range(4):
            if k > 0:
                scales, means = self._spatial(k, run["x_hat"], cf, a, b, t)
            sym16 = idx16 = None
            if stage is not None or dec is not None:
                idx16 = torch.empty(N * H * W, dtype=torch.int16, device=dev)
                sym16 = torch.empty(N * H * W, dtype=torch.int16, device=dev) if dec is None else None
            if dec is None:
                self._mask_step(k, x, None, scales, means, run, sym16, idx16)
            else:   # decoder: the table indexes of this step's masked scales first, then the symbols the coder returns for them
                self._mask_step(k, None, None, scales, means, {"x_hat": None, "s_hat": None}, None, idx16)
                self._mask_step(k, None, dec(k, idx16), scales, means, run, None, None)
            steps.append((sym16, idx16, scales, means))
            if stage is not None:
                stage.append((sym16, idx16))
        return run, steps

    def forward(self, x, context=None, prev_subband=None, write=False):
        from .. import train
        if train.needs_grad(x, self) or train.needs_grad(context, self):
            return self._forward_torch(x, context, prev_subband, write)
        run, steps = self._run(x, context, prev_subband)
        if write:
            masks = self.get_mask_four_parts(x.size(2), x.size(3), x.dtype, x.device)
            return (*[run["x_q"] * masks[k] for k in range(4)], *[run["s_hat"] * masks[k] for k in range(4)], run["x_hat"])
        return run["x_res"], run["x_q"], run["x_hat"], run["s_hat"]

    def compress(self, x, context=None, prev_subband=None):
        return self.forward(x, context, prev_subband, write=True)

    def compress_staged(self, x, context=None, prev_subband=None):
        """compress() for the native coder: returns (x_hat, [(sym16, idx16)] * 4) -- the int16 symbols and scale-table indexes of
        the four masked planes exactly as GaussianEncoder.encode would derive them (entropy_models.py:37-40,266-275), produced
        inside the quantiser kernel instead of by two conversions and device->host copies per step."""
        stage = []
        run, _ = self._run(x, context, prev_subband, stage=stage)
        return run["x_hat"], stage

    def decompress(self, gaussian_encoder, context=None, prev_subband=None):
        def dec(k, idx16):
            cdf, ln, off = gaussian_encoder.get_cdf_info()
            idx = idx16.cpu().numpy()
            sym = gaussian_encoder.entropy_coder.decoder.decode_stream(idx, cdf, ln, off)
            return torch.from_numpy(sym).to(idx16.device)
        run, _ = self._run(None, context, prev_subband, dec=dec)
        return run["x_hat"]
