"""pWave++ analysis / synthesis transform and quantisation on the B200 kernels
(reference: pMCTF/models/pWave.py:26-349, hot-path part only).

What is here: the 4-level lifting transform (encode / decode), quantise / dequantise, the
q_index -> step interpolation, and the reference's own "transform only" loop
spatial_wavelet_dec.  What is NOT here (out of scope for this tier): the
context / entropy models, the rANS coder and the PostProcess net.  `accelerate()` in
models/video/pMCTF_L.py grafts these methods onto an instance of the reference class, which keeps
all of those in stock torch.

state_dict keys are the reference's: wavelet_transform.{lift_h,lift_v}.*, QP, QP_ll.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import ops
from ..layers import LiftingScheme2D

BANDS = ("lh", "hl", "hh")


class _QCache:
    """q = exp(log q_min + (log q_max - log q_min)/20 * q_index) (pWave.py:209-225), evaluated
    with torch on the CPU copy of the parameter so that the value is the one the reference's CPU
    run produces, then cached as a Python float (kernel scalar argument) per parameter version."""

    def __init__(self):
        self.cache = {}

    def scalar(self, param: torch.Tensor, q_index, qp_num: int) -> float:
        key = (param.data_ptr(), param._version, q_index)
        v = self.cache.get(key)
        if v is None:
            p = param.detach().to("cpu", torch.float32)
            lo, hi = p[0:1], p[1:2]
            step = (torch.log(hi) - torch.log(lo)) / (qp_num - 1)
            v = float(torch.exp(torch.log(lo) + step * q_index).reshape(()))
            if len(self.cache) > 4096:
                self.cache.clear()
            self.cache[key] = v
        return v


class pWaveTransform:
    """Mixin with the hot-path methods; shared by the stand-alone `pWave` below and by instances of
    the reference's own class after `accelerate()`."""

    decomp_levels: int
    lossy: bool
    clip_value: float

    # --- q handling (pWave.py:209-229) ------------------------------------------------------
    @staticmethod
    def get_qp_num():
        return 21

    def get_one_q_scale(self, q_scale, q_index):
        min_q = q_scale[0:1, :, :, :]
        max_q = q_scale[1:2, :, :, :]
        step = (torch.log(max_q) - torch.log(min_q)) / (self.get_qp_num() - 1)
        return torch.exp(torch.log(min_q) + step * q_index)

    def get_curr_q(self, q_scale, q_index):
        if isinstance(q_index, list):
            return torch.cat([self.get_one_q_scale(q_scale, i) for i in q_index], dim=0)
        return self.get_one_q_scale(q_scale, q_index)

    def _q_float(self, q) -> float:
        """Scalar value of a [1,1,1,1] step tensor (or a float).  The reference's forward_one_channel passes the SAME step
        tensor to quantize_subband for every band (pWave.py:255,278) and again to dequantize_subbands (:291): the device ->
        host read is done once per tensor object and version (a small identity cache that keeps the tensor alive, so an id
        can not be recycled), not once per subband."""
        if isinstance(q, torch.Tensor):
            if q.numel() != 1:
                raise NotImplementedError("per-sample q_index lists are not supported on the fused path (batch shares one step)")
            cache = self.__dict__.setdefault("_q_seen", [])
            for t, ver, val in cache:
                if t is q and ver == q._version:
                    return val
            val = float(q.detach().reshape(()).to("cpu"))
            cache.append((q, q._version, val))
            if len(cache) > 32:
                del cache[0]
            return val
        return float(q)

    # --- transform (pWave.py:139-157) ---------------------------------------------------------
    def encode(self, x):
        subbands, ll = {}, x
        for lvl in range(self.decomp_levels):
            d = self.wavelet_transform.forward_lift_2d(ll)
            subbands[lvl] = d
            ll = d["ll"]
        return subbands

    def encode_bands(self, x):
        """encode() without the unused row-pass 'l'/'h' dictionary entries."""
        subbands, ll = {}, x
        for lvl in range(self.decomp_levels):
            d = self.wavelet_transform.forward_lift_2d_bands(ll)
            subbands[lvl] = d
            ll = d["ll"]
        return subbands

    def decode(self, subbands):
        """Like the reference, writes each reconstructed `ll` back into the caller's dict (:153-157)."""
        y = None
        for lvl in range(self.decomp_levels - 1, -1, -1):
            y = self.wavelet_transform.backward_lift_2d(subbands[lvl])
            if lvl > 0:
                subbands[lvl - 1]["ll"] = y
        return y

    def decode_dequant(self, subbands_hat, q_scale, q_scale_ll):
        """dequantize_subbands (pWave.py:191-202) + decode (:150-157) with the divisions fused into
        the first loads of each inverse level."""
        if self._train(subbands_hat[0]["lh"], q_scale, q_scale_ll):
            return self.decode(self.dequantize_subbands(subbands_hat, q_scale, q_scale_ll))
        q = self._q_float(q_scale) if self.lossy else 1.0
        qll = self._q_float(q_scale_ll) if self.lossy else 1.0
        top = self.decomp_levels - 1
        ll = subbands_hat[top]["ll"]
        for lvl in range(top, -1, -1):
            sb = dict(subbands_hat[lvl])
            sb["ll"] = ll
            ll = self.wavelet_transform.backward_lift_2d(sb, ll_div=qll if lvl == top else 1.0, q=q)
        return ll

    # --- quantisation (pWave.py:168-202) -------------------------------------------------------
    def _train(self, *ts):
        from .. import train
        return train.needs_grad(*ts, self)

    def post_process(self, x_hat):
        """dequantModule(x_hat / 256) * 256 (pWave.py:299-300): the scaling is fused into our PostProcess kernels; a stock torch
        module (the reference's own, or a test double) is called the reference's way."""
        from ..layers.postprocessing import PostProcess
        if isinstance(self.dequantModule, PostProcess):
            return self.dequantModule(x_hat, 1.0 / self.dynamic_range, self.dynamic_range)
        return self.dequantModule(x_hat / self.dynamic_range) * self.dynamic_range

    def _round(self, s):
        """RoundNoGradient (layers.py:71-80) on a clamped band: the fused kernel with q = 1 on the evaluation path."""
        if self._train(s):
            from ..layers import RoundNoGradient
            return RoundNoGradient.apply(s) if self.lossy else s
        return ops.quantize(s, 1.0, self.clip_value, self.lossy, do_round=self.lossy)

    def q_pair(self, q_index=None, qp_scale=None):
        """(q, q_ll) as the reference derives them in forward/compress (pWave.py:231-238,383-392).  Outside autograd the pair is
        cached per (q_index, qp_scale tensor, parameter versions) and the SAME tensor objects are handed out again, so that the
        one device -> host read of their values (_q_float) happens once, not once per coded plane."""
        if q_index is None:
            return self.QP[-1], self.QP_ll[-1]
        cacheable = not torch.is_grad_enabled() and not isinstance(q_index, list)
        if cacheable:
            cache = self.__dict__.setdefault("_qpair_cache", {})
            key = (q_index, id(qp_scale) if isinstance(qp_scale, torch.Tensor) else qp_scale,
                   qp_scale._version if isinstance(qp_scale, torch.Tensor) else None, self.QP.data_ptr(), self.QP._version,
                   self.QP_ll.data_ptr(), self.QP_ll._version)
            hit = cache.get(key)
            if hit is not None:
                return hit[0], hit[1]
        q, qll = self.get_curr_q(self.QP, q_index), self.get_curr_q(self.QP_ll, q_index)
        if qp_scale is not None:
            q, qll = q * qp_scale, qll * qp_scale
        if cacheable:
            if len(cache) > 64:
                cache.clear()
            cache[key] = (q, qll, qp_scale)      # qp_scale kept alive: its id is part of the key
        return q, qll

    def quantize_subband(self, subband, q_scale):
        """clamp(s * q, +-clip), not rounded (pWave.py:184-189)."""
        if self._train(subband, q_scale):
            from .. import train
            return train.quantize(subband, q_scale, self.clip_value, self.lossy, do_round=False)
        return ops.quantize(subband, self._q_float(q_scale), self.clip_value, self.lossy, do_round=False)

    def quantize_subbands(self, subbands, q_scale, q_scale_ll):
        """round(clamp(s * q)) for every coded band (pWave.py:168-182)."""
        train_mode = self._train(subbands[0]["lh"], q_scale, q_scale_ll)
        if train_mode:
            from .. import train
            q, qll = q_scale, q_scale_ll     # tensors stay in the graph: QP / QP_ll / hp_q_scale receive gradients
        else:
            q, qll = self._q_float(q_scale), self._q_float(q_scale_ll)
        out = {}
        for lvl in range(self.decomp_levels - 1, -1, -1):
            out[lvl] = {}
            for b in (("ll",) + BANDS if lvl == self.decomp_levels - 1 else BANDS):
                if train_mode:
                    out[lvl][b] = train.quantize(subbands[lvl][b], qll if b == "ll" else q, self.clip_value, self.lossy, self.lossy)
                else:
                    out[lvl][b] = ops.quantize(subbands[lvl][b], qll if b == "ll" else q, self.clip_value, self.lossy,
                                               do_round=self.lossy)
        return out

    def dequantize_subbands(self, subbands_hat, q_scale, q_scale_ll):
        if self._train(next(iter(subbands_hat[0].values())), q_scale, q_scale_ll):
            if not self.lossy:
                return {lvl: dict(v) for lvl, v in subbands_hat.items()}
            return {lvl: {b: v / (q_scale_ll if b == "ll" else q_scale) for b, v in subbands_hat[lvl].items()} for lvl in subbands_hat}
        q, qll = self._q_float(q_scale), self._q_float(q_scale_ll)
        out = {}
        for lvl in range(self.decomp_levels - 1, -1, -1):
            out[lvl] = {b: ops.dequantize(v, qll if b == "ll" else q, self.lossy) for b, v in subbands_hat[lvl].items()}
        return out

    def dequantize_subband(self, subband, q_scale):
        if self._train(subband, q_scale):
            return subband / q_scale if self.lossy else subband
        return ops.dequantize(subband, self._q_float(q_scale), self.lossy)

    def _step_table(self, q, planes: int, device):
        """per-plane step table (CUDA float [P]) from a float (cached) or a tensor of P steps"""
        if isinstance(q, torch.Tensor):
            if q.numel() != planes:
                raise RuntimeError(f"expected one step per plane ({planes}), got {q.numel()}")
            return q.detach().to(device=device, dtype=torch.float32).reshape(planes).contiguous()
        cache = self.__dict__.setdefault("_qtabs", {})
        key = (float(q), planes, str(device))
        t = cache.get(key)
        if t is None:
            if len(cache) > 64:
                cache.clear()
            t = cache[key] = torch.full((planes,), float(q), dtype=torch.float32, device=device)
        return t

    def band_layout(self, H: int, W: int):
        """[(level, band, offset, elements)] of one plane's coefficients in coding order (ll of the coarsest level, then lh / hl /
        hh from the coarsest level down: pWave.py:254-290) -- the layout of the int16 symbol rows code_planes() emits."""
        out, off = [], 0
        top = self.decomp_levels - 1
        for lvl in range(top, -1, -1):
            n = (H >> (lvl + 1)) * (W >> (lvl + 1))
            for b in (("ll",) + BANDS if lvl == top else BANDS):
                out.append((lvl, b, off, n))
                off += n
        return out

    def code_planes(self, x, q, qll, sym16=None):
        """spatial_wavelet_dec (pWave.py:314-349, without PostProcess) on a batch of planes [P,1,H,W]: analysis, then ONE launch per
        band that quantises (per-plane steps: q / qll are floats or CUDA tensors with P entries, so frames of different
        temporal levels can share the batch), gathers the exact symbol statistics, optionally emits the symbols as int16
        (sym16: CUDA int16 [P, H*W], rows in band_layout() order -- what the reference hands to its entropy coder,
        entropy_models.py:37-40) and dequantises; then synthesis.  -> (x_hat, int64 [P,2] = sum|sym|, #nonzero)."""
        P, _, H, W = x.shape
        y = self.encode_bands(x)
        qt, qllt = self._step_table(q, P, x.device), self._step_table(qll, P, x.device)
        layout = self.band_layout(H, W)
        stats = torch.zeros((len(layout), P, 2), dtype=torch.int64, device=x.device)
        hat = {lvl: {} for lvl in range(self.decomp_levels)}
        for i, (lvl, b, off, _) in enumerate(layout):
            hat[lvl][b] = ops.quantize_code(y[lvl][b], qllt if b == "ll" else qt, stats[i], self.clip_value, self.lossy, dequant=True,
                                            sym16=sym16, sym16_offset=off)
        x_hat = self.decode(hat)          # the bands are dequantised already
        return x_hat, stats.sum(0)

    # --- the coder around the transform (pWave.py:244-312, 381-529): entropy-parameter networks + rANS --------------------------
    def has_entropy_model(self) -> bool:
        return hasattr(self, "context_fusion") and hasattr(self, "context_prediction") and hasattr(self, "em")

    def _coded_forward(self, x, q_scale, q_scale_ll):
        """forward_one_channel (pWave.py:244-312): the LL band through its autoregressive model, then lh / hl / hh from the coarsest
        level down, each through its four-step model with the long-term context of everything coded before it; the rate is the
        Laplace estimate of the entropy model (gaussian_model.py:50-55)."""
        top = self.decomp_levels - 1
        y = self.encode(x)
        hat = {lvl: {} for lvl in range(self.decomp_levels)}
        bits = {lvl: {} for lvl in range(top, -1, -1)}
        ll_hat = self._round(self.quantize_subband(y[top]["ll"], q_scale_ll))
        scales, means = self.context_fusion[str(top)]["ll"](ll_hat).chunk(2, dim=1)
        bits[top]["ll"] = self.em.get_y_laplace_bits(ll_hat - means, scales)
        hat[top]["ll"] = ll_hat
        bits["bits_total"] = torch.sum(bits[top]["ll"], dim=(1, 2, 3))
        self.context_prediction.init_sequential(list(ll_hat.size()), ll_hat.device)
        context = self.context_prediction.forward_one_subband(ll_hat, "ll", top)["context"]
        for lvl in range(top, -1, -1):
            for i, b in enumerate(BANDS):
                ctx_b = context.chunk(3, dim=1)[i]
                prev = hat[lvl + 1][b] if lvl < top else None
                s = self.quantize_subband(y[lvl][b], q_scale)
                _, s_q, s_hat, sc = self.context_fusion[str(lvl)][b](s, context=ctx_b, prev_subband=prev)
                hat[lvl][b] = s_hat
                bits[lvl][b] = self.em.get_y_laplace_bits(s_q, sc)
                bits["bits_total"] = bits["bits_total"] + torch.sum(bits[lvl][b], dim=(1, 2, 3))
                context = self.context_prediction.forward_one_subband(s_hat, b, lvl)["context"]
        x_hat = self.decode(self.dequantize_subbands(hat, q_scale, q_scale_ll))
        if self.lossy and hasattr(self, "dequantModule"):
            x_hat = self.post_process(x_hat)
        n = x_hat.size(0)
        return {"x_hat": x_hat, "bits": bits, "likelihoods": bits, "subbands": hat,
                "bpp_total": bits["bits_total"].sum() / (x_hat.size(2) * x_hat.size(3) * n), "bits_total": bits["bits_total"].sum() / n,
                "mse": torch.mean((x - x_hat) ** 2)}

    def update(self, force=False):
        self.em.update(force)

    def _ll_sequential(self, size, device, dtype, symbols=None):
        """The LL band, coefficient by coefficient in raster order (pWave.py:531-584): parameters from the causal neighbourhood
        of what is known so far, then the residual to the mean is encoded (symbols given) or decoded.  Returns the band."""
        B, Cc, H, W = size
        net = self.context_fusion[str(self.decomp_levels - 1)]["ll"]
        enc = self.em.gaussian_encoder
        if hasattr(net, "ar_encode") and torch.device(device).type == "cuda":   # one kernel per band / per coefficient (csrc/pmctf_llar.cu)
            cdf, ln, off = enc.get_cdf_info()
            if symbols is not None:
                ll_hat, sym16, idx16 = net.ar_encode(symbols)
                # one coefficient (B symbols) per share-out over the sub-streams, as the reference's per-coefficient encode calls
                enc.entropy_coder.encoder.encode_with_indexes(sym16, idx16, cdf, ln, off, chunk=B)
                return ll_hat
            dec = enc.entropy_coder.decoder
            band = net.ar_decode_band(size, dec, cdf, ln, off, device) if hasattr(net, "ar_decode_band") else None
            if band is None:          # several sub-streams / large batches: one launch + one host rANS step per coefficient
                band = net.ar_decode(size, lambda idx: dec.decode_stream(idx, cdf, ln, off), device)
            return band.to(dtype)
        pad = 1
        if symbols is not None:
            plane = torch.nn.functional.pad(symbols, (pad, pad, pad, pad))
            out = torch.zeros_like(symbols)
        else:
            plane = torch.zeros((B, Cc, H + 2 * pad, W + 2 * pad), dtype=dtype, device=device)
        for h in range(H):
            for w in range(W):
                scale, mean = net.forward_sequential(plane, h, w).chunk(2, dim=1)
                if symbols is not None:
                    cur = plane[:, 0:1, h + pad:h + pad + 1, w + pad:w + pad + 1]
                    res = torch.round(torch.round(cur) - mean)
                    out[:, :, h, w] = torch.round(res + mean)[:, :, 0, 0]
                    enc.encode(res, scale)
                else:
                    rec = enc.decode_stream(scale, dtype, device) + mean
                    plane[:, :, h + pad, w + pad] = torch.round(rec)[:, :, 0, 0]
        net.sequential_init = False
        return out if symbols is not None else plane[:, :, pad:-pad, pad:-pad].contiguous()

    @torch.no_grad()
    def compress(self, x, sideinfo=None, file_name=None, q_index=None, skip_decoding=False, qp_scale=None):
        """pWave.compress (pWave.py:381-463): same arguments; writes the container of stream_helper.encode_image (:201-207) when
        file_name is given and returns x_hat.  The four masked symbol planes of every band go to the native rANS coder as int16
        symbols + table indexes produced by the quantiser kernel (ContextFusionFourStep.compress_staged)."""
        _, num_channels, height, width = sideinfo
        q, qll = self.q_pair(q_index, qp_scale)
        x_in = torch.cat([x[:, i:i + 1] for i in range(3)], dim=0) if num_channels == 3 else x
        top = self.decomp_levels - 1
        y = self.encode(x_in)
        hat = {lvl: {} for lvl in range(self.decomp_levels)}
        ll = torch.round(self.quantize_subband(y[top]["ll"], qll))
        coder, enc = self.em.entropy_coder, self.em.gaussian_encoder
        coder.reset()
        if skip_decoding:       # encoder-only shortcut of the reference: parameters of the whole band in one pass
            scales, means = self.context_fusion[str(top)]["ll"](ll).chunk(2, dim=1)
            res = torch.round(ll - means)
            ll_hat = torch.round(res + means)
            enc.encode(res, scales)
        else:
            ll_hat = self._ll_sequential(list(ll.size()), ll.device, ll.dtype, symbols=ll)
        hat[top]["ll"] = ll_hat
        self.context_prediction.init_sequential(list(ll.size()), ll.device)
        context = self.context_prediction.forward_one_subband(ll_hat, "ll", top)["context"]
        cdf, ln, off = enc.get_cdf_info()
        for lvl in range(top, -1, -1):
            for i, b in enumerate(BANDS):
                ctx_b = context.chunk(3, dim=1)[i].contiguous()
                prev = hat[lvl + 1][b] if lvl < top else None
                s = self.quantize_subband(y[lvl][b], q)
                net = self.context_fusion[str(lvl)][b]
                if hasattr(net, "compress_staged") and s.is_cuda:
                    s_hat, staged = net.compress_staged(s, context=ctx_b, prev_subband=prev)
                    host = [(a.cpu(), c.cpu()) for a, c in staged]           # 4 + 4 small copies per band, one synchronisation
                    for sym16, idx16 in host:
                        coder.encoder.encode_with_indexes(sym16.numpy(), idx16.numpy(), cdf, ln, off)
                else:
                    o = net.compress(s, context=ctx_b, prev_subband=prev)
                    s_hat = o[8]
                    for k in range(4):
                        enc.encode(o[k], o[4 + k])
                hat[lvl][b] = s_hat
                context = self.context_prediction.forward_one_subband(s_hat, b, lvl)["context"]
        x_hat = self.decode(self.dequantize_subbands(hat, q, qll))
        if self.lossy and hasattr(self, "dequantModule"):
            x_hat = self.post_process(x_hat)
        coder.flush()
        stream = coder.get_encoded_stream()
        self.last_stream = stream
        if file_name is not None:
            import struct
            with open(file_name, "wb") as f:                                 # stream_helper.py:201-207
                f.write(struct.pack(">3I", height, width, num_channels))
                f.write(struct.pack(">I", len(stream)))
                f.write(stream)
        if num_channels == 3:
            x_hat = torch.cat([x_hat[i:i + 1] for i in range(3)], dim=1)
        return x_hat

    @torch.no_grad()
    def decompress(self, file_name, padding=64, q_index=None, qp_scale=None):
        """pWave.decompress (pWave.py:467-529); file_name may also be the bytes object of such a file."""
        import struct
        q, qll = self.q_pair(q_index, qp_scale)
        blob = file_name if isinstance(file_name, (bytes, bytearray)) else open(file_name, "rb").read()
        height, width, num_channel = struct.unpack(">3I", blob[:12])
        (n,) = struct.unpack(">I", blob[12:16])
        self.em.entropy_coder.set_stream(bytes(blob[16:16 + n]))
        p0 = next(self.parameters())
        dtype, device = p0.dtype, p0.device
        new_h, new_w = (height + padding - 1) // padding * padding, (width + padding - 1) // padding * padding
        top = self.decomp_levels - 1
        size = [num_channel, 1, new_h >> (top + 1), new_w >> (top + 1)]
        hat = {lvl: {} for lvl in range(top, -1, -1)}
        hat[top]["ll"] = ll = self._ll_sequential(size, device, dtype)
        self.context_prediction.init_sequential(list(ll.size()), device)
        context = self.context_prediction.forward_one_subband(ll, "ll", top)["context"]
        for lvl in range(top, -1, -1):
            for i, b in enumerate(BANDS):
                ctx_b = context.chunk(3, dim=1)[i].contiguous()
                prev = hat[lvl + 1][b] if lvl < top else None
                hat[lvl][b] = s_hat = self.context_fusion[str(lvl)][b].decompress(self.em.gaussian_encoder, context=ctx_b, prev_subband=prev)
                context = self.context_prediction.forward_one_subband(s_hat, b, lvl)["context"]
        x_hat = self.decode(self.dequantize_subbands(hat, q, qll))
        if self.lossy and hasattr(self, "dequantModule"):
            x_hat = self.post_process(x_hat)
        if num_channel == 3:
            x_hat = torch.cat([x_hat[i:i + 1] for i in range(3)], dim=1)
        return {"x_hat": x_hat}

    # --- the reference's transform-only loop (pWave.py:314-349) --------------------------------
    def spatial_wavelet_dec(self, x, q_scale=None, q_scale_ll=None, post_process=True, return_symbols=False):
        """encode -> round(clamp(s*q)) on every band -> dequantise -> decode [-> PostProcess].
        `post_process` applies self.dequantModule when the instance has one (reference class)."""
        if q_scale is None:
            q_scale, q_scale_ll = self.QP[-1], self.QP_ll[-1]
        y = self.encode_bands(x)
        hat = self.quantize_subbands(y, q_scale, q_scale_ll)
        x_hat = self.decode_dequant(hat, q_scale, q_scale_ll)
        if post_process and self.lossy and hasattr(self, "dequantModule"):
            x_hat = self.post_process(x_hat)
        return (x_hat, hat) if return_symbols else x_hat


class pWave(pWaveTransform, nn.Module):
    """Stand-alone hot-path subset of the reference's pWave (pWave.py:26-98): same constructor,
    same parameter names for the transform and the quantiser."""

    def __init__(self, bitdepth=8, decomp_levels=4, lossy=True, postprocess=False, entropy_model=False):
        super().__init__()
        self.bitdepth = 8
        self.dynamic_range = float(2 ** bitdepth)
        self.lossy = lossy
        self.in_channels = 1
        self.decomp_levels = decomp_levels
        self.wavelet_transform = LiftingScheme2D(bitdepth=bitdepth, lossy=lossy, in_channels=1)
        self.clip_value = 8192.0 if lossy else float(torch.iinfo(torch.int16).max)  # pWave.py:55-58
        self._qc = _QCache()
        if entropy_model:           # pWave.py:60-82: the whole coder, module tree and names as in the reference
            from ..entropy_models.gaussian_model import CompressionModel
            from ..layers.context_fusion import ContextFusionSubband
            from ..layers.context_fusion_4step import ContextFusionFourStep
            from ..layers.long_context import SubbandContext
            self.context_prediction = SubbandContext(in_channels=1, decomp_levels=decomp_levels)
        if (postprocess or entropy_model) and lossy:   # pWave.py:61-62; optional so that hot-path-only checkpoints keep loading
            from ..layers.postprocessing import PostProcess
            self.dequantModule = PostProcess(in_channels=1, out_channels=1)
        if entropy_model:
            self.num_params = 2
            # the reference's coder configuration (pWave.py:66: no worker thread, one sub-stream).  PMCTF_EC_THREAD=1 queues and codes
            # on a worker thread under the GPU work of the next band (same bytes); PMCTF_STREAM_PART=n writes n sub-streams coded
            # concurrently (the reference's container for stream_part = n, py_rans.cpp:67-113)
            self.em = CompressionModel(y_distribution="laplace", ec_thread=os.environ.get("PMCTF_EC_THREAD", "0") == "1",
                                       stream_part=int(os.environ.get("PMCTF_STREAM_PART", "1")))
            self.context_fusion = nn.ModuleDict({
                str(lvl): nn.ModuleDict({b: ContextFusionFourStep(in_channels=1, num_features=112, num_parameters=2, lossy=lossy,
                                                                  ctx_channels=2 if lvl < decomp_levels - 1 else 1) for b in BANDS})
                for lvl in range(decomp_levels)})
            self.context_fusion[str(decomp_levels - 1)]["ll"] = ContextFusionSubband(num_features=128, num_parameters=2, context=False,
                                                                                     in_channels=1)
        self.QP = nn.Parameter(torch.ones((2, 1, 1, 1), dtype=torch.float) * 1 / 16)      # pWave.py:84-85
        self.QP_ll = nn.Parameter(torch.ones((2, 1, 1, 1), dtype=torch.float) * 1 / 16)

    def forward(self, x, q_index=None, qp_scale=None):
        """pWave.forward (pWave.py:231-242): same signature, same step derivation, then forward_one_channel."""
        if q_index is not None:
            q, qll = self.q_pair(q_index, qp_scale)
            return self.forward_one_channel(x, q, qll)
        return self.forward_one_channel(x)

    def forward_one_channel(self, x, q_scale=None, q_scale_ll=None):
        """pWave.forward_one_channel (pWave.py:244-312) with the reference's own call sequence on the hot path -- encode ->
        quantize_subband(ll, q_ll) -> round -> per band (coarsest level first, lh/hl/hh) quantize_subband(s, q) -> symbols ->
        dequantize_subbands -> decode -- and the reference's return keys.  The entropy-parameter networks and the
        PostProcess net are SURVEY.md section 8f rows (not built yet): the symbols are the zero-mean ones
        (round(clamp(s*q)), which is what spatial_wavelet_dec, pWave.py:314-349, defines), `x_hat` is the synthesis output
        before PostProcess, and the keys that only the entropy model can fill (`bits`, `likelihoods`, `bits_total`,
        `bpp_total`) are NaN / None instead of invented numbers."""
        if q_scale is None:
            q_scale, q_scale_ll = self.QP[-1], self.QP_ll[-1]
        if self.has_entropy_model():
            return self._coded_forward(x, q_scale, q_scale_ll)
        top = self.decomp_levels - 1
        y = self.encode(x)
        subbands_hat = {lvl: {} for lvl in range(self.decomp_levels)}
        ll = self.quantize_subband(y[top]["ll"], q_scale_ll)
        subbands_hat[top]["ll"] = self._round(ll)
        for lvl in range(top, -1, -1):
            for b in BANDS:
                subbands_hat[lvl][b] = self._round(self.quantize_subband(y[lvl][b], q_scale))
        x_hat = self.decode(self.dequantize_subbands(subbands_hat, q_scale, q_scale_ll))
        if self.lossy and hasattr(self, "dequantModule"):          # pWave.py:298-300
            x_hat = self.post_process(x_hat)
        nan = torch.full((), float("nan"), device=x.device)
        return {"x_hat": x_hat, "bits": None, "likelihoods": None, "subbands": subbands_hat, "bpp_total": nan, "bits_total": nan,
                "mse": torch.mean((x - x_hat) ** 2)}
