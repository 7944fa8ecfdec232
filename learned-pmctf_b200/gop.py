"""Dyadic GOP schedule of the MCTF hot path (reference: test_pMCTF_flex.py:131-291, training twin
train_pMCTF_L.py:161-208) as a batched on-device pipeline.

The reference walks (stage, group) pairs one at a time in Python.  Pairs inside a stage are
independent for the lifting work (SURVEY.md section 8e "Inside a GOP"), so here every stage is ONE
batched call per plane type: the even/odd frames of the previous level are passed to the kernels as
strided views (no copies), the H frames of a stage are coded by `hp_coder` as one batch, and
the inverse writes straight into the interleaved positions of the next finer level.

Hot path only (SURVEY.md section 8d "(1) hot-path-only"): motion vectors are an INPUT here (the reference
gets them from SpyNet + the MV codec, out of scope), and the entropy model is replaced by exact
integer symbol statistics.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import ops

# per-frame statistics gathered across ranks (float64)
STAT_FIELDS = ("frame_type", "sym_abs_sum", "sym_nonzero", "sse_y", "sse_cb", "sse_cr", "psnr_yuv", "pixels")
N_STATS = len(STAT_FIELDS)


def num_stages(gop_size: int) -> int:
    """log2(gop_size), asserting a power of two (test_pMCTF_flex.py:91-94)."""
    s = 1
    while 2 ** s < gop_size:
        s += 1
    if 2 ** s != gop_size:
        raise ValueError(f"gop_size must be a power of two >= 2, got {gop_size}")
    return s


def dyadic_schedule(gop_size: int):
    """[[(ref_idx, cur_idx), ...] per stage]: stage s pairs frame g*2^(s+1) (-> L) with +2^s (-> H)
    (test_pMCTF_flex.py:138-146)."""
    out = []
    n = gop_size
    for s in range(num_stages(gop_size)):
        n //= 2
        step = 2 ** s
        out.append([(g * 2 * step, g * 2 * step + step) for g in range(n)])
    return out


def get_padding_size(height: int, width: int, p: int = 128):
    """(left, right, top, bottom) zero padding to a multiple of p (pMCTF/utils/stream_helper.py:23-32)."""
    new_h = (height + p - 1) // p * p
    new_w = (width + p - 1) // p * p
    return 0, new_w - width, 0, new_h - height


def psnr_from_sse(sse: float, pixels: int) -> float:
    """PSNR for 8-bit samples (pMCTF/utils/util.py PSNR: 10 log10(255^2 / mse), 100 dB at mse == 0 is NOT
    applied by the reference; a zero mse gives inf there, here as well)."""
    mse = sse / pixels
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 * 255.0 / mse)


class GopCodec:
    """MCTF analysis -> pWave++ analysis -> quantise -> dequantise -> pWave++ synthesis -> MCTF synthesis
    for whole GOPs.  `model` is a learned_pmctf_b200.pMCTF (or an accelerate()d reference model)."""

    def __init__(self, model, gop_size: int = 16, q_index: Optional[int] = None, concurrent_chroma: bool = True):
        self.m = model
        self.gop_size = gop_size
        self.stages = num_stages(gop_size)
        self.q_index = q_index
        self._qcache = {}
        self._qtab = {}
        self._hostsym = {}
        # luma and chroma never meet on the path (the motion field is read-only), so code_gop() runs the chroma chain on a
        # side stream: the tail of every persistent launch (CTAs that run out of tiles) and the small launches of the deep
        # wavelet levels are filled by the other chain's kernels
        self.concurrent_chroma = concurrent_chroma
        self._side = {}

    # step sizes exactly as forward_one_stage derives them (pMCTF_L.py:343-349, pWave.py:231-238)
    def _q_params(self, coder: str, stage: int):
        m = self.m
        if coder == "hp":
            ps = [m.hp_coder.QP, m.hp_coder.QP_ll]
            if self.q_index is not None and m.quant_stage:
                ps.append(m.hp_q_scale[min(m.num_me_stages - 1, stage)])
            return ps
        return [m.lp_coder.QP, m.lp_coder.QP_ll]

    def q_pair(self, coder: str, stage: int):
        """Host scalars of the quantisation steps, cached per (coder, stage, q_index) AND per version of the parameters they
        derive from, so that load_state_dict / an optimiser step is seen."""
        key = (coder, stage, self.q_index) + tuple((p.data_ptr(), p._version) for p in self._q_params(coder, stage))
        v = self._qcache.get(key)
        if v is None:
            m = self.m
            if coder == "hp":
                me = min(m.num_me_stages - 1, stage)
                scale = m.hp_qp_scale(me, self.q_index) if (self.q_index is not None and m.quant_stage) else None
                q, qll = m.hp_coder.q_pair(self.q_index, scale)
            else:
                q, qll = m.lp_coder.q_pair(self.q_index)
            v = (float(q.detach().reshape(-1)[0].cpu()), float(qll.detach().reshape(-1)[0].cpu()))
            if len(self._qcache) > 256:
                self._qcache.clear()
            self._qcache[key] = v
        return v

    # ---------------------------------------------------------------------------------------
    def analysis(self, Y: torch.Tensor, C: torch.Tensor, mvs: Sequence[torch.Tensor]):
        """Temporal decomposition.  Y [G,1,H,W], C [G,2,1,H/2,W/2] (Cb, Cr), mvs[s] [G>>(s+1),2,H,W] (luma
        fields; chroma uses the fused 2x2-mean/2).  -> (L_y, L_c, [(H_y, H_c) per stage])."""
        m = self.m
        Ly, Lc, Hs = Y, C, []
        for s in range(self.stages):
            me = min(m.num_me_stages - 1, s)
            n = Ly.size(0) // 2
            if mvs[s].size(0) != n:
                raise RuntimeError(f"stage {s}: expected {n} motion fields, got {mvs[s].size(0)}")
            L2y, Hy, _, _ = m.forward_MCTF(Ly[0::2], Ly[1::2], mvs[s], stage_idx=me, want_pred=False)
            L2c, Hc, _, _ = m.forward_MCTF(Lc[0::2], Lc[1::2], mvs[s], stage_idx=me, mv_down=True, want_pred=False)
            Hs.append((Hy, Hc))
            Ly, Lc = L2y, L2c
        return Ly, Lc, Hs

    def _code(self, coder, x, q, qll, stats, sym16=None):
        planes = x.reshape(-1, 1, x.size(-2), x.size(-1))
        x_hat, st = coder.code_planes(planes, q, qll, sym16=sym16)
        if stats is not None:
            stats.append(st)
        return x_hat.view(x.shape)

    def _hp_step_tables(self, planes_per_frame: int, device):
        """Per-plane (q, q_ll) tables of the H frames of ALL temporal levels in coding order (level 0 frames first): the step of
        hp_coder is scaled per temporal level (pMCTF_L.py:343-347), so the frames of a GOP can share one coder batch only with
        per-plane steps.  Cached per parameter version."""
        pairs = [self.q_pair("hp", s) for s in range(self.stages)]
        key = (tuple(pairs), planes_per_frame, str(device))
        t = self._qtab.get(key)
        if t is None:
            q, qll = [], []
            for s, (a, b) in enumerate(pairs):
                n = (self.gop_size >> (s + 1)) * planes_per_frame
                q += [a] * n
                qll += [b] * n
            if len(self._qtab) > 32:
                self._qtab.clear()
            t = self._qtab[key] = (torch.tensor(q, dtype=torch.float32, device=device), torch.tensor(qll, dtype=torch.float32, device=device))
        return t

    def code(self, Ly, Lc, Hs, want_stats: bool = True):
        """Spatial coding of every temporal subband frame: hp_coder on the H frames of each stage (step scaled
        per temporal level), lp_coder on the final L (code_lt, test_pMCTF_flex.py:205).  Returns the decoded
        frames and per-plane symbol statistics."""
        m = self.m
        sy, sc = ([] if want_stats else None), ([] if want_stats else None)
        Hhat = []
        for s, (Hy, Hc) in enumerate(Hs):
            q, qll = self.q_pair("hp", s)
            Hhat.append((self._code(m.hp_coder, Hy, q, qll, sy), self._code(m.hp_coder, Hc, q, qll, sc)))
        q, qll = self.q_pair("lp", 0)
        Ly_hat = self._code(m.lp_coder, Ly, q, qll, sy)
        Lc_hat = self._code(m.lp_coder, Lc, q, qll, sc)
        return Ly_hat, Lc_hat, Hhat, sy, sc

    def synthesis(self, Ly, Lc, Hhat, mvs):
        """Temporal reconstruction, coarsest stage first (test_pMCTF_flex.py:268-291)."""
        m = self.m
        for s in range(self.stages - 1, -1, -1):
            me = min(m.num_me_stages - 1, s)
            Hy, Hc = Hhat[s]
            n = Ly.size(0)
            by = torch.empty((2 * n,) + tuple(Ly.shape[1:]), dtype=torch.float32, device=Ly.device)
            bc = torch.empty((2 * n,) + tuple(Lc.shape[1:]), dtype=torch.float32, device=Lc.device)
            m.inverse_MCTF(Ly, Hy, mvs[s], stage_idx=me, out_ref=by[0::2], out_cur=by[1::2])
            m.inverse_MCTF(Lc, Hc, mvs[s], downscale=True, stage_idx=me, out_ref=bc[0::2], out_cur=bc[1::2])
            Ly, Lc = by, bc
        return Ly, Lc

    def _stream_for(self, key, device):
        st = self._side.get(key)
        if st is None:
            st = self._side[key] = torch.cuda.Stream(device)
        return st

    def _chain(self, X, mvs, chroma: bool, want_stats: bool, sym16=None):
        """analysis -> code -> synthesis of ONE plane type (luma [G,1,H,W] or chroma [G,2,1,h,w]), everything on the
        current stream.  The H frames of every temporal level are written into ONE buffer (level 0 first: the coding order of
        test_pMCTF_flex.py:138-146) and coded by hp_coder as ONE batch with per-plane steps, so the deep wavelet levels launch
        G-1 planes at a time instead of G/2, G/4, ... 1.  sym16: optional int16 [G * planes_per_frame, coefficients] buffer that
        receives the symbols (rows in coding order: H frames, then the final L)."""
        m = self.m
        G = X.size(0)
        ppf = X.numel() // (G * X.size(-2) * X.size(-1))          # planes per frame: 1 (luma) or 2 (Cb, Cr)
        Hbuf = torch.empty((G - 1,) + tuple(X.shape[1:]), dtype=torch.float32, device=X.device)
        L, off = X, 0
        for s in range(self.stages):
            me = min(m.num_me_stages - 1, s)
            n = L.size(0) // 2
            if mvs[s].size(0) != n:
                raise RuntimeError(f"stage {s}: expected {n} motion fields, got {mvs[s].size(0)}")
            L, _, _, _ = m.forward_MCTF(L[0::2], L[1::2], mvs[s], stage_idx=me, mv_down=chroma, want_pred=False, out_H=Hbuf[off:off + n])
            off += n
        st = [] if want_stats else None
        qt, qllt = self._hp_step_tables(ppf, X.device)
        nh = (G - 1) * ppf
        Hhat = self._code(m.hp_coder, Hbuf, qt, qllt, st, None if sym16 is None else sym16[:nh])
        q, qll = self.q_pair("lp", 0)
        L = self._code(m.lp_coder, L, q, qll, st, None if sym16 is None else sym16[nh:])
        end = G - 1
        for s in range(self.stages - 1, -1, -1):
            me = min(m.num_me_stages - 1, s)
            n = L.size(0)
            b = torch.empty((2 * n,) + tuple(L.shape[1:]), dtype=torch.float32, device=L.device)
            m.inverse_MCTF(L, Hhat[end - n:end], mvs[s], downscale=chroma, stage_idx=me, out_ref=b[0::2], out_cur=b[1::2])
            end -= n
            L = b
        return L, st

    # ---------------------------------------------------------------------------------------
    def coding_order(self):
        """GOP frame index of row i of the symbol buffers / statistics: the H frames level by level, then the low-pass frame."""
        return self._frame_of_plane()

    def _frame_of_plane(self):
        """GOP frame index of every coded plane in the order `code` emits statistics."""
        order = []
        for s, pairs in enumerate(dyadic_schedule(self.gop_size)):
            order += [cur for _, cur in pairs]
        order.append(0)
        return order

    @torch.no_grad()
    def code_gop(self, Y, C, mvs, orig_y_u8=None, orig_c_u8=None, want_stats=True, symbols=None):
        """One GOP through the whole hot path.  Returns (rec_Y, rec_C, stats) with stats an fp64 device
        tensor [G, N_STATS] (None if want_stats is False).  orig_*_u8: the un-padded 8-bit originals
        ([G,h0,w0], [G,2,h0/2,w0/2]) for the distortion columns; without them SSE is taken against Y / C.
        symbols: optional dict that receives the quantised symbols as int16 device tensors, "y" [G, H*W] and "c" [G*2, H*W/4]
        (rows in CODING order -- row i belongs to frame coding_order()[i], chroma rows Cb, Cr per frame -- each row in
        pWave.band_layout() order): what the reference copies to the host for its entropy coder (entropy_models.py:37-40)."""
        G = self.gop_size
        if Y.size(0) != G or C.size(0) != G:
            raise RuntimeError(f"expected {G} frames, got {Y.size(0)} / {C.size(0)}")
        for s in range(self.stages):   # resolve the quantisation steps (host scalars) before any stream forks
            self.q_pair("hp", s)
        self.q_pair("lp", 0)
        sym_y = sym_c = None
        if symbols is not None:
            sym_y = torch.empty((G, Y.size(-2) * Y.size(-1)), dtype=torch.int16, device=Y.device)
            sym_c = torch.empty((2 * G, C.size(-2) * C.size(-1)), dtype=torch.int16, device=Y.device)
            symbols["y"], symbols["c"] = sym_y, sym_c
        if self.concurrent_chroma and Y.is_cuda:
            cur = torch.cuda.current_stream(Y.device)
            side = self._stream_for((Y.device, "chroma", cur.cuda_stream), Y.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                rec_c, sc = self._chain(C, mvs, True, want_stats, sym_c)
            rec_y, sy = self._chain(Y, mvs, False, want_stats, sym_y)
            cur.wait_stream(side)
            for t in [rec_c] + (sc or []):
                t.record_stream(cur)
        else:
            rec_y, sy = self._chain(Y, mvs, False, want_stats, sym_y)
            rec_c, sc = self._chain(C, mvs, True, want_stats, sym_c)
        if not want_stats:
            return rec_y, rec_c, None
        dev = Y.device
        order = torch.tensor(self._frame_of_plane(), device=dev)
        sym = torch.zeros((G, 2), dtype=torch.int64, device=dev)
        sym.index_add_(0, order, torch.cat(sy))                       # luma planes: one per frame
        sym.index_add_(0, order, torch.cat(sc).view(-1, 2, 2).sum(1))  # chroma: Cb + Cr per frame
        if orig_y_u8 is not None:
            sse_y = ops.frame_sse(rec_y, orig_y_u8)
            sse_c = ops.frame_sse(rec_c, orig_c_u8.reshape(-1, orig_c_u8.size(-2), orig_c_u8.size(-1))).view(G, 2)
            px = orig_y_u8.size(-2) * orig_y_u8.size(-1)
        else:
            sse_y = ((rec_y.clamp(0, 255).round() - Y) ** 2).sum((1, 2, 3)).to(torch.int64)
            sse_c = ((rec_c.clamp(0, 255).round() - C) ** 2).sum((2, 3, 4)).to(torch.int64)
            px = Y.size(-2) * Y.size(-1)
        st = torch.zeros((G, N_STATS), dtype=torch.float64, device=dev)
        st[1:, 0] = 1.0  # frame 0 is the coded low-pass ("I"-type 0), the rest are H frames (type 1), :236,254
        st[:, 1:3] = sym.to(torch.float64)
        st[:, 3] = sse_y.to(torch.float64)
        st[:, 4:6] = sse_c.to(torch.float64)
        mse = torch.stack([st[:, 3] / px, st[:, 4] / (px // 4), st[:, 5] / (px // 4)], 1)
        psnr = 10.0 * torch.log10(255.0 * 255.0 / mse)
        st[:, 6] = (6.0 * psnr[:, 0] + psnr[:, 1] + psnr[:, 2]) / 8.0  # test_pMCTF_flex.py:322
        st[:, 7] = px
        return rec_y, rec_c, st

    # ---------------------------------------------------------------------------------------
    # GOPs in flight (GOPs are independent: test_pMCTF_flex.py:131-134); each lane = a luma + a chroma stream.  Measured on B200:
    # the luma | chroma overlap of one GOP already fills the launch tails (145 -> 162 frames/s); a second lane adds nothing (160).
    GOP_LANES = 1

    def _lane(self, device, g: int):
        cur = torch.cuda.current_stream(device)
        if not self.concurrent_chroma or self.GOP_LANES <= 1:
            return cur
        return self._stream_for((device, "lane", cur.cuda_stream, g % self.GOP_LANES), device)

    @torch.no_grad()
    def code_sequence(self, Y, C, mvs_per_gop, y_u8=None, c_u8=None):
        """Whole GOPs of a device-resident sequence: Y [F,1,H,W], C [F,2,1,H/2,W/2], mvs_per_gop[g] as code_gop takes them.
        Consecutive GOPs alternate between GOP_LANES stream pairs, so that the serial tail of one GOP's launch chain is
        covered by another GOP's kernels.  -> fp64 statistics [F/G, G, N_STATS] on the device, ordered on the current stream."""
        G = self.gop_size
        n_gops = Y.size(0) // G
        if Y.size(0) % G or len(mvs_per_gop) != n_gops:
            raise RuntimeError("frame count must be a multiple of the GOP size, one motion-field list per GOP")
        dev = Y.device
        cur = torch.cuda.current_stream(dev)
        out, lanes = [], []
        for g in range(n_gops):
            lane = self._lane(dev, g)
            sl = slice(g * G, (g + 1) * G)
            if lane is not cur:
                if lane not in lanes:
                    lanes.append(lane)
                    lane.wait_stream(cur)
                with torch.cuda.stream(lane):
                    _, _, st = self.code_gop(Y[sl], C[sl], mvs_per_gop[g], None if y_u8 is None else y_u8[sl],
                                             None if c_u8 is None else c_u8[sl])
            else:
                _, _, st = self.code_gop(Y[sl], C[sl], mvs_per_gop[g], None if y_u8 is None else y_u8[sl],
                                         None if c_u8 is None else c_u8[sl])
            out.append(st)
        for lane in lanes:
            cur.wait_stream(lane)
        for st in out:
            st.record_stream(cur)
        return torch.stack(out)

    @torch.no_grad()
    def _host_symbols(self, frames: int, ny: int, nc: int):
        """Pinned host buffers that receive the int16 symbols of a sequence (allocated once per shape, reused by later calls)."""
        key = (frames, ny, nc)
        b = self._hostsym.get(key)
        if b is None:
            self._hostsym.clear()
            b = self._hostsym[key] = (torch.empty((frames, ny), dtype=torch.int16).pin_memory(),
                                      torch.empty((frames, 2, nc), dtype=torch.int16).pin_memory())
        return b

    def code_sequence_host(self, y_u8: torch.Tensor, c_u8: torch.Tensor, mvs_host: List[Sequence[torch.Tensor]],
                           psize: int = 128, return_symbols: bool = False):
        """End-to-end entry point on HOST buffers: y_u8 [F,h0,w0], c_u8 [F,2,h0/2,w0/2] uint8 (pinned for
        speed) and per-GOP host motion fields; frames are uploaded GOP by GOP on a copy stream (overlapping the
        kernels of earlier GOPs), unpacked + zero padded on the device, coded (GOPs alternating between GOP_LANES stream
        pairs), and the [F, N_STATS] statistics are returned on the host.
        return_symbols=True additionally brings the path's PRODUCT to the host: the quantised symbols of every coded plane as
        int16, copied GOP by GOP on a second copy stream into pinned buffers while later GOPs compute -> (stats,
        {"y": int16 [F, Hp*Wp], "c": int16 [F, 2, Hp*Wp/4], "coding_order": frame index of each row within its GOP}); rows
        g*G .. g*G+G-1 hold GOP g in coding order, every row in pWave.band_layout() order.  The buffers are reused by the
        next call on the same codec."""
        G = self.gop_size
        F_, h0, w0 = y_u8.shape
        if F_ % G:
            raise RuntimeError("frame count must be a multiple of the GOP size (the reference pads the sequence, :96-99)")
        _, pr, _, pb = get_padding_size(h0, w0, psize)
        hp, wp = h0 + pb, w0 + pr
        dev = next(self.m.parameters()).device
        cur = torch.cuda.current_stream(dev)
        copy = self._stream_for((dev, "copy"), dev)
        copy.wait_stream(cur)
        d2h = self._stream_for((dev, "d2h"), dev) if return_symbols else None
        hy, hc = self._host_symbols(F_, hp * wp, (hp // 2) * (wp // 2)) if return_symbols else (None, None)

        def upload(g):  # H2D of GOP g on the copy stream, overlapping the kernels of the GOPs before it
            with torch.cuda.stream(copy):
                t = (y_u8[g * G:(g + 1) * G].to(dev, non_blocking=True), c_u8[g * G:(g + 1) * G].to(dev, non_blocking=True),
                     [m.to(dev, non_blocking=True) for m in mvs_host[g]])
                ev = torch.cuda.Event()
                ev.record(copy)
            return t, ev

        out, lanes = [], []
        nxt = upload(0)
        for g in range(F_ // G):
            (yd, cd, mvd), ev = nxt
            lane = self._lane(dev, g)
            if lane is not cur and lane not in lanes:
                lanes.append(lane)
                lane.wait_stream(cur)
            lane.wait_event(ev)
            for t in [yd, cd] + mvd:
                t.record_stream(lane)
            if g + 1 < F_ // G:
                nxt = upload(g + 1)
            sym = {} if return_symbols else None
            with torch.cuda.stream(lane):
                Y = ops.unpack_u8(yd, hp, wp)
                C = ops.unpack_u8(cd.view(-1, h0 // 2, w0 // 2), hp // 2, wp // 2).view(G, 2, 1, hp // 2, wp // 2)
                _, _, st = self.code_gop(Y, C, mvd, yd, cd, symbols=sym)
            if return_symbols:   # device -> pinned host on its own stream, overlapping the next GOP's kernels
                d2h.wait_stream(lane)
                with torch.cuda.stream(d2h):
                    hy[g * G:(g + 1) * G].copy_(sym["y"], non_blocking=True)
                    hc[g * G:(g + 1) * G].copy_(sym["c"].view(G, 2, -1), non_blocking=True)
                sym["y"].record_stream(d2h), sym["c"].record_stream(d2h)
            out.append(st)
        for lane in lanes:
            cur.wait_stream(lane)
        for st in out:
            st.record_stream(cur)
        res = torch.cat(out).cpu()              # synchronises: every kernel of the sequence has finished
        if return_symbols:
            d2h.synchronize()
        ops.check_tc_error(dev, "GopCodec.code_sequence_host")
        if return_symbols:
            return res, {"y": hy, "c": hc, "coding_order": self.coding_order()}
        return res


# -------------------------------------------------------------------------------------------------
def training_loss_hot_path(model, clips: torch.Tensor, mvs: Optional[Sequence[torch.Tensor]] = None, q_index: int = 8,
                           lmbda: float = 0.05):
    """The hot-path part of one training iteration (train_pMCTF_L.py:161-226, luma only as :137): dyadic MCTF analysis of a
    batch of clips [B,G,1,H,W] (pairs of a stage batched along N), hp/lp pWave++ transform with the straight-through
    quantiser, synthesis, inverse MCTF, and a rate-proxy + distortion loss.  Differentiable: under model.train() every op
    runs on the training kernels (train.py).  mvs[s]: [B * pairs_s, 2, H, W] motion fields (zeros if None; the reference
    gets them from SpyNet + MV codec, out of scope)."""
    B, G, _, H, W = clips.shape
    S = num_stages(G)
    level = [clips[:, f] for f in range(G)]                     # each [B,1,H,W]
    coded, rate = [], 0.0
    for s in range(S):
        n = len(level) // 2
        ref = torch.cat(level[0::2], 0)                          # [n*B,1,H,W], pair-major
        cur = torch.cat(level[1::2], 0)
        mv = mvs[s] if mvs is not None else torch.zeros((n * B, 2, H, W), dtype=torch.float32, device=clips.device)
        me = min(model.num_me_stages - 1, s)
        L, Hh, _, _ = model.forward_MCTF(ref, cur, mv, stage_idx=me)
        q, qll = model.hp_coder.q_pair(q_index, model.hp_qp_scale(me, q_index))
        x_hat, hat = model.hp_coder.spatial_wavelet_dec(Hh, q, qll, post_process=False, return_symbols=True)
        rate = rate + sum(v.abs().mean() for lvl in hat for v in hat[lvl].values())
        coded.append((x_hat, mv, me))
        level = list(L.split(B, 0))
    q, qll = model.lp_coder.q_pair(q_index)
    L_hat, hat = model.lp_coder.spatial_wavelet_dec(level[0], q, qll, post_process=False, return_symbols=True)
    rate = rate + sum(v.abs().mean() for lvl in hat for v in hat[lvl].values())
    rec = [L_hat]
    for s in range(S - 1, -1, -1):
        x_hat, mv, me = coded[s]
        r, c = model.inverse_MCTF(torch.cat(rec, 0), x_hat, mv, stage_idx=me)
        rec = [t for pair in zip(r.split(B, 0), c.split(B, 0)) for t in pair]
    dist = sum(((a - clips[:, f]) ** 2).mean() for f, a in enumerate(rec)) / G
    return dist + lmbda * rate, dist


def synthetic_sequence(seq_id: int, n_frames: int, h0: int = 1080, w0: int = 1920, device="cuda"):
    """SURVEY.md section 8d synthetic input: band-limited noise (3x 9x9 box blur, rescaled to 16..235) translated by a
    per-sequence constant velocity, plus N(0, 2^2) per-frame noise, rounded to 8 bits.
    -> (y_u8 [F,h0,w0], c_u8 [F,2,h0/2,w0/2]) on `device`."""
    g = torch.Generator(device=device)
    g.manual_seed(1000 + seq_id)
    vx, vy = 2 + seq_id % 3, -(1 + seq_id % 2)  # px / frame
    mx, my = abs(vx) * n_frames + 8, abs(vy) * n_frames + 8

    def texture(h, w):
        t = torch.rand((1, 1, h, w), device=device, generator=g)
        for _ in range(3):
            t = torch.nn.functional.avg_pool2d(t, 9, 1, 4)
        t = (t - t.amin()) / (t.amax() - t.amin())
        return 16.0 + t * (235.0 - 16.0)

    def frames(base, h, w, sx, sy, pad_x, pad_y):
        out = torch.empty((n_frames, h, w), dtype=torch.uint8, device=device)
        for f in range(n_frames):
            ox = pad_x + sx * f if sx >= 0 else pad_x * 2 - 8 + sx * f
            oy = pad_y + sy * f if sy >= 0 else pad_y * 2 - 8 + sy * f
            ox, oy = max(0, min(ox, base.size(-1) - w)), max(0, min(oy, base.size(-2) - h))
            fr = base[0, 0, oy:oy + h, ox:ox + w] + 2.0 * torch.randn((h, w), device=device, generator=g)
            out[f] = fr.round().clamp(0, 255).to(torch.uint8)
        return out

    y = frames(texture(h0 + 2 * my, w0 + 2 * mx), h0, w0, vx, vy, mx, my)
    cb = frames(texture(h0 // 2 + my, w0 // 2 + mx), h0 // 2, w0 // 2, vx // 2, vy // 2, mx // 2, my // 2)
    cr = frames(texture(h0 // 2 + my, w0 // 2 + mx), h0 // 2, w0 // 2, vx // 2, vy // 2, mx // 2, my // 2)
    return y, torch.stack([cb, cr], 1)


def synthetic_motion(seq_id: int, gop_idx: int, gop_size: int, hp: int, wp: int, device="cuda"):
    """Injected motion fields: N(0, 4^2) px smoothed 5x5, clipped to +-32, one per pair.
    -> list over stages of [pairs, 2, hp, wp]."""
    g = torch.Generator(device=device)
    g.manual_seed(2000 + 97 * seq_id + gop_idx)
    out = []
    n = gop_size
    for _ in range(num_stages(gop_size)):
        n //= 2
        f = 4.0 * torch.randn((n, 2, hp, wp), device=device, generator=g)
        out.append(torch.nn.functional.avg_pool2d(f, 5, 1, 2).mul_(5.0).clamp_(-32.0, 32.0).contiguous())
    return out
