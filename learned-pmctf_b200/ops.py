"""Tensor-level operators of the hot path.  Every function validates its arguments like the
C ABI does, allocates outputs/workspaces with torch on the input's device, and launches on
torch's current CUDA stream.  CUDA tensors only: there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _native as nat

SCALE_L = float(np.float32(1.149604398860241))  # lifting_1d.py:57-58,98-101 (bior4.4)
SCALE_H = float(np.float32(0.869864451624781))
SCALE_P = float(np.float32(1 / math.sqrt(2)))   # wavelet_transform_temporal_mctf.py:24-25
SCALE_U = 0.5


_MODES = {"default": nat.CONV_DEFAULT, "ffma": nat.CONV_FFMA, "tensor": nat.CONV_TENSOR}


def set_conv_mode(mode: str) -> None:
    """'tensor' (default): conv2/conv3 of PredictUpdate as exact fixed-point implicit GEMMs on the tcgen05 tensor cores;
    'ffma': sequential fp32 FMA chains on the CUDA cores.  This is the process-wide DEFAULT (include/pmctf_b200.h
    PMCTF_CONV_*); a module can pin its own arithmetic with `module.conv_mode = "ffma" | "tensor"`, which travels in the
    conv_mode field of its descriptor and overrides the default for that module's calls only."""
    nat.check(nat.lib().pmctf_set_conv_mode({"ffma": nat.CONV_FFMA, "tensor": nat.CONV_TENSOR}[mode]), "set_conv_mode")


def get_conv_mode() -> str:
    return "tensor" if nat.lib().pmctf_get_conv_mode() == nat.CONV_TENSOR else "ffma"


def conv_mode_code(mode) -> int:
    """'default' | 'ffma' | 'tensor' | None -> PMCTF_CONV_* value of a descriptor's conv_mode field."""
    return _MODES[mode or "default"]


def tc_error_flag() -> int:
    """Watchdog word of the current device: non-zero once a tensor-core step kernel gave up waiting for its MMAs.  Read it
    after synchronising; every later launch on that device fails with PMCTF_ETIMEOUT until clear_tc_error()."""
    return int(nat.lib().pmctf_tc_error_flag())


def check_tc_error(device=None, what: str = "tensor-core lifting step") -> None:
    """Raise if a tensor-core kernel on `device` timed out (its outputs are incomplete).  Called by GopCodec at its
    synchronisation points; cheap (reads one word of pinned host memory)."""
    with _guard(device):
        if tc_error_flag() != 0:
            raise RuntimeError(f"{what}: a tensor-core kernel gave up waiting for its MMAs on {device or 'the current device'}; "
                               "results since the last check are incomplete (ops.clear_tc_error() re-arms the device)")


def clear_tc_error(device=None) -> None:
    with _guard(device):
        nat.check(nat.lib().pmctf_tc_clear_error(), "tc_clear_error")


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOGUARD = _NoGuard()


def _guard(device):
    """Device guard like the one ATen ops carry: the C ABI keys its per-device state on the CURRENT device and launches
    on the stream it is given, so a call on tensors of another device must switch to it first.  No-op (no allocation) when
    the tensors already live on the current device."""
    if device is None:
        return _NOGUARD
    device = torch.device(device)
    if device.type != "cuda":
        return _NOGUARD
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx == torch.cuda.current_device():
        return _NOGUARD
    return torch.cuda.device(idx)


def _same_device(*ts):
    """All CUDA operands of one call must share a device (the kernels dereference raw pointers)."""
    dev = None
    for t in ts:
        if isinstance(t, torch.Tensor):
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"operands live on different devices: {dev} and {t.device}")
    return dev


def _launch(device, what: str, fn, *args) -> None:
    """One C-ABI call on `device`: device guard, the current stream of THAT device as the trailing stream argument,
    non-zero return -> RuntimeError."""
    with _guard(device):
        nat.check(fn(*args, _stream(device)), what)


def _stream(device=None) -> int:
    """Handle of torch's current stream ON THE OPERANDS' DEVICE (not of the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def _chk(t: torch.Tensor, name: str, ndim: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: the pMCTF hot path runs on CUDA tensors only (got {getattr(t, 'device', type(t))}); "
                           "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    return t


def _no_grad_only(*ts):
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts):
        raise NotImplementedError("learned_pmctf_b200: backward of the fused CUDA ops is not implemented yet; "
                                  "call under torch.no_grad() (as test_pMCTF_flex.py does)")


# ---------------------------------------------------------------------------------------------
# Optional live timing of the fused lifting-step launches (bench.py's roofline): while a KernelTimer is
# active every call that launches lift_step_kernel is bracketed by CUDA events on the launching stream.
PU_FLOPS_PER_PX = 9792  # 2*MAC of conv1..conv4


class KernelTimer:
    def __init__(self):
        self.records = []  # (start_event, end_event, launches, pixels_through_PU)

    def __enter__(self):
        global _TIMER
        _TIMER = self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = None

    def summary(self):
        """-> dict(launches, pixels, ms) after the stream has been synchronised."""
        ms = sum(a.elapsed_time(b) for a, b, _, _ in self.records)
        return {"launches": sum(r[2] for r in self.records), "pixels": sum(r[3] for r in self.records), "ms": ms}


_TIMER: Optional[KernelTimer] = None


class _timed:
    __slots__ = ("launches", "pixels", "ev", "dev")

    def __init__(self, launches: int, pixels: int, device=None):
        self.launches, self.pixels, self.dev = launches, pixels, device

    def __enter__(self):
        if _TIMER is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *exc):
        if _TIMER is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.dev))
            _TIMER.records.append((self.ev, e, self.launches, self.pixels))


_LIN_CACHE: dict = {}


def linspace_table(n: int, device) -> torch.Tensor:
    """torch.linspace(-1, 1, n) evaluated with the scalar formula of ATen's CUDA kernel
    (video_net.py:36-39 runs it on the frame's device).  Cached per (n, device), like the
    reference's backward_grid cache (video_net.py:10,33-40)."""
    key = (n, str(device))
    t = _LIN_CACHE.get(key)
    if t is None:
        step = np.float32(2.0) / np.float32(n - 1)
        i = np.arange(n)
        lo = np.float32(-1.0) + step * i.astype(np.float32)
        hi = np.float32(1.0) - step * (n - i - 1).astype(np.float32)
        t = torch.from_numpy(np.where(i < n // 2, lo, hi).astype(np.float32)).to(device)
        _LIN_CACHE[key] = t
    return t


_WS_CACHE: dict = {}


def workspace(floats: int, device, tag: str = "") -> torch.Tensor:
    """Grow-only per-(device, stream, tag) scratch buffer; stream-ordered reuse is safe because
    every consumer is launched on the same stream.  release_workspaces() drops them."""
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), _stream(device), tag)
    t = _WS_CACHE.get(key)
    if t is None or t.numel() < floats:
        t = torch.empty(max(floats, 1), dtype=torch.float32, device=device)
        _WS_CACHE[key] = t
    return t


def release_workspaces(device=None) -> None:
    """Drop the cached scratch buffers (of one device, or all): they are otherwise kept for the life of the process, keyed on
    (device, stream), so a program that cycles through many streams should call this when a stream is retired."""
    if device is None:
        _WS_CACHE.clear()
        return
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    for k in [k for k in _WS_CACHE if k[1] == idx]:
        del _WS_CACHE[k]


def plane_of(t: torch.Tensor) -> nat.Plane:
    """[N,1,H,W] or [G,K,1,H,W] tensor (any strides, no copy) -> strided plane descriptor."""
    if t.dim() == 4 and t.size(1) == 1:
        return nat.Plane(t.data_ptr(), t.stride(0), t.stride(2), t.stride(3), 0, 0)
    if t.dim() == 5 and t.size(2) == 1:
        return nat.Plane(t.data_ptr(), t.stride(1), t.stride(3), t.stride(4), t.stride(0), t.size(1))
    raise RuntimeError(f"expected single-channel planes [N,1,H,W] or [G,K,1,H,W], got {tuple(t.shape)}")


def _planes_shape(t: torch.Tensor):
    """-> (number of planes, H, W)"""
    return (t.size(0) if t.dim() == 4 else t.size(0) * t.size(1)), t.size(-2), t.size(-1)


# ---------------------------------------------------------------------------------------------
def pack_pu(params, out: torch.Tensor):
    """params: (w1,b1,...,w4,b4) CUDA tensors in the state_dict's OIHW layout."""
    ps = [_chk(p.detach().contiguous(), "PredictUpdate parameter") for p in params]
    shapes = [(16, 1, 3, 3), (16,), (16, 16, 3, 3), (16,), (16, 16, 3, 3), (16,), (1, 16, 3, 3), (1,)]
    for p, s in zip(ps, shapes):
        if tuple(p.shape) != s:
            raise RuntimeError(f"PredictUpdate parameter has shape {tuple(p.shape)}, expected {s} (lifting_1d.py:28-34)")
    assert out.numel() == nat.PU_PACKED_FLOATS and out.is_contiguous()
    _dev = _same_device(out, *ps)
    _launch(_dev, "pack_pu_weights", nat.lib().pmctf_pack_pu_weights, *[p.data_ptr() for p in ps], out.data_ptr())
    return ps  # keep alive until the stream has consumed them


def release_pu(packed: torch.Tensor, blocks: int = 1) -> None:
    """Forget the parameters registered for `blocks` consecutive packed blocks (the owner is about to free or rewrite them)."""
    for i in range(blocks):
        nat.lib().pmctf_release_pu_weights(packed.data_ptr() + 4 * i * nat.PU_PACKED_FLOATS)


def flow_warp(im: torch.Tensor, flow: torch.Tensor, sign: float = 1.0, lin_x=None, lin_y=None, round_out=False):
    """flow_warp / torch_warp: video_net.py:32-55."""
    _no_grad_only(im, flow)
    im, flow = _chk(im, "im", 4).contiguous(), _chk(flow, "flow", 4).contiguous()
    N, Cc, H, W = im.shape
    if flow.shape[1:] != (2, H, W) or N % flow.shape[0]:
        raise RuntimeError(f"flow shape {tuple(flow.shape)} does not match image {tuple(im.shape)}")
    lx = linspace_table(W, im.device) if lin_x is None else _chk(lin_x, "lin_x")
    ly = linspace_table(H, im.device) if lin_y is None else _chk(lin_y, "lin_y")
    out = torch.empty_like(im)
    _dev = _same_device(im, flow, lx, ly)
    _launch(_dev, "flow_warp", nat.lib().pmctf_flow_warp, im.data_ptr(), flow.data_ptr(), lx.data_ptr(), ly.data_ptr(), out.data_ptr(),
                                        N, Cc, H, W, flow.shape[0], sign, int(round_out))
    return out


def chroma_mv_down(mv: torch.Tensor):
    """bilineardownsacling(mv) / 2: video_net.py:66-71, pMCTF_L.py:317,336,401."""
    _no_grad_only(mv)
    mv = _chk(mv, "mv", 4).contiguous()
    N, two, H, W = mv.shape
    if two != 2:
        raise RuntimeError("mv must be [N,2,H,W]")
    out = torch.empty((N, 2, H // 2, W // 2), dtype=torch.float32, device=mv.device)
    _dev = mv.device
    _launch(_dev, "chroma_mv_down", nat.lib().pmctf_chroma_mv_down, mv.data_ptr(), out.data_ptr(), N, H, W)
    return out


def predict_update(x: torch.Tensor, packed: torch.Tensor, in_mul: float = 1.0, conv_mode=None):
    """PredictUpdate.forward: lifting_1d.py:36-49.  conv_mode: None/'default' | 'ffma' | 'tensor' for this call."""
    _no_grad_only(x)
    x = _chk(x, "x", 4).contiguous()
    N, Cc, H, W = x.shape
    if Cc != 1:
        raise RuntimeError("PredictUpdate on the hot path is single-channel (in_ch=1)")
    out = torch.empty_like(x)
    _dev = _same_device(x, packed)
    _launch(_dev, "predict_update", nat.lib().pmctf_predict_update, x.data_ptr(), packed.data_ptr(), in_mul, out.data_ptr(), N, H, W,
            conv_mode_code(conv_mode))
    return out


def temporal_filter(x: torch.Tensor, t: nat.Temporal, which: int):
    _no_grad_only(x)
    x = _chk(x, "x", 4).contiguous()
    N, Cc, H, W = x.shape
    if Cc != 1:
        raise RuntimeError("TemporalLifting is single-channel")
    out = torch.empty_like(x)
    _dev = x.device
    _launch(_dev, "temporal_filter", nat.lib().pmctf_temporal_filter, x.data_ptr(), C.byref(t), which, out.data_ptr(), N, H, W)
    return out


def _mv_args(frame: torch.Tensor, mv: torch.Tensor, mv_down: bool):
    if frame.dim() not in (4, 5) or frame.size(-3) != 1:
        raise RuntimeError("MCTF planes are single-channel ([N,1,H,W] or [G,K,1,H,W]; chroma is batched over N, "
                           "test_pMCTF_flex.py:153-164)")
    N, H, W = _planes_shape(frame)
    exp = (2, 2 * H, 2 * W) if mv_down else (2, H, W)
    if mv.dim() != 4 or tuple(mv.shape[1:]) != exp or N % mv.shape[0]:
        raise RuntimeError(f"mv shape {tuple(mv.shape)} does not match frame {tuple(frame.shape)} (mv_down={mv_down})")
    return N, H, W


def _out_like(ref: torch.Tensor, out: Optional[torch.Tensor], name: str) -> torch.Tensor:
    if out is None:
        return torch.empty(ref.shape, dtype=torch.float32, device=ref.device)
    _chk(out, name)
    if out.shape != ref.shape:
        raise RuntimeError(f"{name}: shape {tuple(out.shape)} != {tuple(ref.shape)}")
    return out


def forward_mctf(ref, cur, mv, t: nat.Temporal, mv_down=False, want_pred=True, lin_x=None, lin_y=None,
                 out_L=None, out_H=None):
    """pMCTF.forward_MCTF: pMCTF_L.py:297-312 -> (L_t, H_t, pred, inv).  ref / cur (and the optional
    preallocated out_L / out_H) may be arbitrary strided views: no copies are made."""
    _no_grad_only(ref, cur, mv)
    ref, cur, mv = _chk(ref, "ref"), _chk(cur, "cur"), _chk(mv, "mv", 4).contiguous()
    if ref.shape != cur.shape:
        raise RuntimeError("ref and cur must have the same shape")
    N, H, W = _mv_args(ref, mv, mv_down)
    lx = linspace_table(W, ref.device) if lin_x is None else lin_x
    ly = linspace_table(H, ref.device) if lin_y is None else lin_y
    L, Hh = _out_like(ref, out_L, "out_L"), _out_like(ref, out_H, "out_H")
    pred = torch.empty(ref.shape, dtype=torch.float32, device=ref.device) if want_pred else None
    inv = torch.empty(ref.shape, dtype=torch.float32, device=ref.device) if want_pred else None
    pr, pc, pL, pH = plane_of(ref), plane_of(cur), plane_of(L), plane_of(Hh)
    pp, pi = (plane_of(pred), plane_of(inv)) if want_pred else (None, None)
    _dev = _same_device(ref, cur, mv, L, Hh, lx, ly)
    with _timed(2, 2 * N * H * W, _dev):
        _launch(_dev, "forward_mctf", nat.lib().pmctf_forward_mctf, C.byref(pr), C.byref(pc), mv.data_ptr(), mv.shape[0], int(mv_down),
                                               lx.data_ptr(), ly.data_ptr(), C.byref(t), C.byref(pL), C.byref(pH),
                                               C.byref(pp) if want_pred else None, C.byref(pi) if want_pred else None,
                                               N, H, W)
    return L, Hh, pred, inv


def inverse_mctf(L, Hh, mv, t: nat.Temporal, mv_down=False, lin_x=None, lin_y=None, out_ref=None, out_cur=None):
    """pMCTF.inverse_MCTF: pMCTF_L.py:314-330 -> (ref, cur); strided views accepted as in forward_mctf."""
    _no_grad_only(L, Hh, mv)
    L, Hh, mv = _chk(L, "L_t"), _chk(Hh, "H_t"), _chk(mv, "mv", 4).contiguous()
    if L.shape != Hh.shape:
        raise RuntimeError("L_t and H_t must have the same shape")
    N, H, W = _mv_args(L, mv, mv_down)
    lx = linspace_table(W, L.device) if lin_x is None else lin_x
    ly = linspace_table(H, L.device) if lin_y is None else lin_y
    ref, cur = _out_like(L, out_ref, "out_ref"), _out_like(L, out_cur, "out_cur")
    pL, pH, pr, pc = plane_of(L), plane_of(Hh), plane_of(ref), plane_of(cur)
    _dev = _same_device(L, Hh, mv, ref, cur, lx, ly)
    with _timed(2, 2 * N * H * W, _dev):
        _launch(_dev, "inverse_mctf", nat.lib().pmctf_inverse_mctf, C.byref(pL), C.byref(pH), mv.data_ptr(), mv.shape[0], int(mv_down),
                                               lx.data_ptr(), ly.data_ptr(), C.byref(t), C.byref(pr), C.byref(pc),
                                               N, H, W)
    return ref, cur


def iwave1d_forward(x: torch.Tensor, p: nat.IWave):
    """iWave1D.forward_lift on a (possibly permuted) [N,1,H,W] view: lifting_1d.py:103-145.
    Outputs are allocated with the same dimension order as the input's memory, so a transposed
    input gives transposed-view outputs exactly like the reference's permute chain."""
    _no_grad_only(x)
    _chk(x, "x", 4)
    N, Cc, H, W = x.shape
    if Cc != 1 or H % 2 or H < 4:
        raise RuntimeError(f"forward_lift needs [N,1,H,W] with even H >= 4, got {tuple(x.shape)}")
    h2 = H // 2
    if x.stride(3) <= x.stride(2):
        l = torch.empty((N, 1, h2, W), dtype=torch.float32, device=x.device)
        h = torch.empty_like(l)
    else:  # transposed view: keep W as the slow axis in memory
        l = torch.empty((N, 1, W, h2), dtype=torch.float32, device=x.device).permute(0, 1, 3, 2)
        h = torch.empty((N, 1, W, h2), dtype=torch.float32, device=x.device).permute(0, 1, 3, 2)
    ws = workspace(N * h2 * W, x.device, "iw1d")
    px, pl, ph = plane_of(x), plane_of(l), plane_of(h)
    _dev = x.device
    _launch(_dev, "iwave1d_forward", nat.lib().pmctf_iwave1d_forward, C.byref(px), C.byref(p), C.byref(pl), C.byref(ph), N, h2, W,
                                              ws.data_ptr(), ws.numel())
    return l, h


def iwave1d_backward(l: torch.Tensor, h: torch.Tensor, p: nat.IWave):
    """iWave1D.backward_lift: lifting_1d.py:147-189."""
    _no_grad_only(l, h)
    _chk(l, "l", 4), _chk(h, "h", 4)
    if l.shape != h.shape or l.size(1) != 1 or l.size(2) < 2:
        raise RuntimeError(f"backward_lift needs two [N,1,h,W] bands with h >= 2, got {tuple(l.shape)} / {tuple(h.shape)}")
    N, _, h2, W = l.shape
    if l.stride(3) <= l.stride(2):
        x = torch.empty((N, 1, 2 * h2, W), dtype=torch.float32, device=l.device)
    else:
        x = torch.empty((N, 1, W, 2 * h2), dtype=torch.float32, device=l.device).permute(0, 1, 3, 2)
    ws = workspace(2 * N * h2 * W, l.device, "iw1d")
    pl, ph, px = plane_of(l), plane_of(h), plane_of(x)
    _dev = _same_device(l, h)
    _launch(_dev, "iwave1d_backward", nat.lib().pmctf_iwave1d_backward, C.byref(pl), C.byref(ph), C.byref(p), C.byref(px), N, h2, W,
                                               ws.data_ptr(), ws.numel())
    return x


def lift2d_forward(x: torch.Tensor, p: nat.IWave, want_lh_rows: bool = False):
    """LiftingScheme2D.forward_lift_2d: wavelet_transform.py:25-43 -> dict (ll, lh, hl, hh [, l, h])."""
    _no_grad_only(x)
    x = _chk(x, "x", 4).contiguous()
    N, Cc, H, W = x.shape
    if Cc != 1 or H % 2 or W % 2 or H < 4 or W < 4:
        raise RuntimeError(f"forward_lift_2d needs [N,1,H,W] with even H, W >= 4, got {tuple(x.shape)}")
    bands = torch.empty((4, N, 1, H // 2, W // 2), dtype=torch.float32, device=x.device)
    rows = torch.empty((2, N, 1, H // 2, W), dtype=torch.float32, device=x.device) if want_lh_rows else None
    ws = workspace(2 * N * H * W, x.device, "l2d")
    _dev = x.device
    with _timed(8, 4 * N * H * W, _dev):  # 4 row steps on N*H/2*W px + 4 column steps on 2N*W/2*H/2 px
        _launch(_dev, "lift2d_forward", nat.lib().pmctf_lift2d_forward, x.data_ptr(), C.byref(p), bands[0].data_ptr(), bands[1].data_ptr(),
                                                 bands[2].data_ptr(), bands[3].data_ptr(),
                                                 rows[0].data_ptr() if want_lh_rows else None,
                                                 rows[1].data_ptr() if want_lh_rows else None,
                                                 N, H, W, ws.data_ptr(), ws.numel())
    d = {"ll": bands[0], "lh": bands[1], "hl": bands[2], "hh": bands[3]}
    if want_lh_rows:  # the reference returns the transposed views (wavelet_transform.py:32,37,42)
        d["l"], d["h"] = rows[0].permute(0, 1, 3, 2), rows[1].permute(0, 1, 3, 2)
    return d


def lift2d_backward(ll, lh, hl, hh, p: nat.IWave, ll_div: float = 1.0, q: float = 1.0):
    """LiftingScheme2D.backward_lift_2d: wavelet_transform.py:45-57, optionally with the
    dequantise of pWave.py:191-202 fused (ll / ll_div, details / q)."""
    _no_grad_only(ll, lh, hl, hh)
    ts = [_chk(t, n, 4).contiguous() for t, n in ((ll, "ll"), (lh, "lh"), (hl, "hl"), (hh, "hh"))]
    if any(t.shape != ts[0].shape for t in ts) or ts[0].size(1) != 1:
        raise RuntimeError("backward_lift_2d needs four [N,1,h,w] subbands of equal shape")
    N, _, h2, w2 = ts[0].shape
    if h2 < 2 or w2 < 2:
        raise RuntimeError("subbands must be at least 2x2 (reflection padding)")
    H, W = 2 * h2, 2 * w2
    x = torch.empty((N, 1, H, W), dtype=torch.float32, device=ts[0].device)
    ws = workspace(2 * N * H * W, x.device, "l2d")
    _dev = _same_device(*ts)
    with _timed(8, 4 * N * H * W, _dev):
        _launch(_dev, "lift2d_backward", nat.lib().pmctf_lift2d_backward_q, ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr(),
                                                    ll_div, q, C.byref(p), x.data_ptr(), N, H, W, ws.data_ptr(), ws.numel())
    return x


def quantize(s: torch.Tensor, q: float, clip: float = 8192.0, lossy: bool = True, do_round: bool = True):
    """[round](clamp(s*q, +-clip)): pWave.py:184-189,256-257,337; layers.py:71-92."""
    _no_grad_only(s)
    s = _chk(s, "subband").contiguous()
    out = torch.empty_like(s)
    if s.numel() == 0:
        return out
    _dev = s.device
    _launch(_dev, "quantize", nat.lib().pmctf_quantize, s.data_ptr(), q, clip, int(lossy), int(do_round), out.data_ptr(), s.numel())
    return out


def dequantize(s_hat: torch.Tensor, q: float, lossy: bool = True):
    """s_hat / q: pWave.py:191-202."""
    _no_grad_only(s_hat)
    s_hat = _chk(s_hat, "subband").contiguous()
    out = torch.empty_like(s_hat)
    if s_hat.numel() == 0:
        return out
    _dev = s_hat.device
    _launch(_dev, "dequantize", nat.lib().pmctf_dequantize, s_hat.data_ptr(), q, int(lossy), out.data_ptr(), s_hat.numel())
    return out


def quantize_stats(s: torch.Tensor, q: float, stats: torch.Tensor, clip: float = 8192.0, lossy: bool = True):
    """round(clamp(s*q)) on [P, ...] planes; stats (int64 [P,2], caller-zeroed) accumulates sum|sym| and #nonzero."""
    _no_grad_only(s)
    s = _chk(s, "subband").contiguous()
    out = torch.empty_like(s)
    planes = s.size(0) if s.dim() > 0 else 0
    if s.numel() == 0:
        return out
    if stats.dtype != torch.int64 or not stats.is_cuda or not stats.is_contiguous() or stats.numel() < 2 * planes:
        raise RuntimeError("stats must be a contiguous CUDA int64 tensor with 2 entries per plane")
    _dev = _same_device(s, stats)
    _launch(_dev, "quantize_stats", nat.lib().pmctf_quantize_stats, s.data_ptr(), q, clip, int(lossy), out.data_ptr(), planes, s.numel() // planes,
                                             stats.data_ptr())
    return out


def quantize_code(s: torch.Tensor, q_tab: torch.Tensor, stats: Optional[torch.Tensor] = None, clip: float = 8192.0, lossy: bool = True,
                  dequant: bool = True, sym16: Optional[torch.Tensor] = None, sym16_offset: int = 0):
    """One band of a plane batch [P, ...] with per-plane steps q_tab (CUDA float [P]): -> fp32 tensor of the dequantised symbols
    (or the symbols, dequant=False).  sym16: int16 [P, plane_total] buffer that receives this band's symbols at column
    `sym16_offset` of every plane's row; stats: int64 [P,2] accumulators (caller-zeroed)."""
    _no_grad_only(s)
    s = _chk(s, "subband").contiguous()
    out = torch.empty_like(s)
    planes = s.size(0) if s.dim() > 0 else 0
    if s.numel() == 0:
        return out
    elems = s.numel() // planes
    if q_tab.dtype != torch.float32 or not q_tab.is_cuda or q_tab.numel() != planes or not q_tab.is_contiguous():
        raise RuntimeError("q_tab must be a contiguous CUDA float32 tensor with one step per plane")
    if stats is not None and (stats.dtype != torch.int64 or not stats.is_cuda or not stats.is_contiguous() or stats.numel() < 2 * planes):
        raise RuntimeError("stats must be a contiguous CUDA int64 tensor with 2 entries per plane")
    sp, stride = None, 0
    if sym16 is not None:
        if sym16.dtype != torch.int16 or not sym16.is_cuda or sym16.dim() != 2 or sym16.size(0) != planes or sym16.stride(1) != 1 \
                or sym16_offset + elems > sym16.size(1):
            raise RuntimeError("sym16 must be a CUDA int16 tensor [planes, coefficients per plane] with room for this band")
        sp, stride = sym16.data_ptr() + 2 * sym16_offset, sym16.stride(0)
    _dev = _same_device(s, q_tab, stats, sym16)
    _launch(_dev, "quantize_code", nat.lib().pmctf_quantize_code, s.data_ptr(), q_tab.data_ptr(), clip, int(lossy), int(dequant),
            out.data_ptr(), sp, stride, planes, elems, stats.data_ptr() if stats is not None else None)
    return out


_WSB_CACHE: dict = {}


def byte_workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    """Grow-only per-(device, stream, tag) byte buffer (256-byte aligned by the caching allocator)."""
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream(device), tag)
    t = _WSB_CACHE.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _WSB_CACHE[key] = t
    return t


def postprocess(x: torch.Tensor, desc: "nat.PostProcessD", in_mul: float = 1.0, out_mul: float = 1.0):
    """PostProcess.forward (postprocessing.py:35-44) on [N,1,H,W]: y = PostProcess(x * in_mul) * out_mul.  The 64 -> 64 layers run
    on the tensor cores (bf16 operands, fp32 accumulation); one plane's feature maps live in a cached workspace (1 KB / pixel)."""
    _no_grad_only(x)
    x = _chk(x, "x", 4).contiguous()
    N, Cc, H, W = x.shape
    if Cc != 1:
        raise RuntimeError("PostProcess is single-channel (pWave.py:62)")
    y = torch.empty_like(x)
    need = int(nat.lib().pmctf_postprocess_workspace(H, W))
    ws = byte_workspace(need, x.device, "pp")
    _launch(x.device, "postprocess", nat.lib().pmctf_postprocess, x.data_ptr(), C.byref(desc), in_mul, out_mul, y.data_ptr(), N, H, W,
            ws.data_ptr(), ws.numel())
    return y


def pp_layout_bf16(x: torch.Tensor) -> torch.Tensor:
    """[N,64,H,W] float -> the operand layout of the PostProcess kernels: bf16 [N,8,H,W,8] (a pixel's 8 channels = one 16-byte record)."""
    N, Cc, H, W = x.shape
    return x.reshape(N, Cc // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def pp_layout_f32(x: torch.Tensor) -> torch.Tensor:
    """[N,64,H,W] float -> fp32 [N,16,H,W,4] (residual stream layout)."""
    N, Cc, H, W = x.shape
    return x.reshape(N, Cc // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous().float()


def pp_unlayout(t: torch.Tensor) -> torch.Tensor:
    """[N,G,H,W,g] (either layout) -> [N,G*g,H,W] fp32."""
    N, G, H, W, g = t.shape
    return t.permute(0, 1, 4, 2, 3).reshape(N, G * g, H, W).float()


def pp_conv64(x_bf16: torch.Tensor, packed_w: torch.Tensor, bias: torch.Tensor, co: int = 64, residual=None, slope: float = 1.0,
              want_f32: bool = True, want_bf16: bool = False):
    """One tensor-core 3x3 convolution 64 -> 64 on a bf16 feature map in operand layout [N,8,H,W,8] (building block of
    PostProcess; tests).  residual / fp32 output: [N,16,H,W,4]."""
    if x_bf16.dtype != torch.bfloat16 or not x_bf16.is_cuda or x_bf16.dim() != 5 or x_bf16.size(1) != 8 or x_bf16.size(4) != 8 \
            or not x_bf16.is_contiguous():
        raise RuntimeError("pp_conv64 expects a contiguous CUDA bfloat16 tensor [N,8,H,W,8] (ops.pp_layout_bf16)")
    N, _, H, W, _ = x_bf16.shape
    of = torch.empty((N, 16, H, W, 4), dtype=torch.float32, device=x_bf16.device) if want_f32 else None
    ob = torch.empty((N, 8, H, W, 8), dtype=torch.bfloat16, device=x_bf16.device) if want_bf16 else None
    _launch(x_bf16.device, "pp_conv64", nat.lib().pmctf_pp_conv64, x_bf16.data_ptr(), packed_w.data_ptr(), bias.data_ptr(), co,
            residual.data_ptr() if residual is not None else None, slope, of.data_ptr() if of is not None else None,
            ob.data_ptr() if ob is not None else None, None, 1.0, 1.0, None, N, H, W)
    return of, ob


def unpack_u8(src: torch.Tensor, hp: int, wp: int, out: Optional[torch.Tensor] = None):
    """uint8 planes [n,h0,w0] -> fp32 [n,1,hp,wp], zero padded bottom/right (test_pMCTF_flex.py:151-192)."""
    if not src.is_cuda or src.dtype != torch.uint8 or src.dim() != 3:
        raise RuntimeError("unpack_u8 expects a CUDA uint8 tensor [n,h0,w0]")
    src = src.contiguous()
    n, h0, w0 = src.shape
    if out is None:
        out = torch.empty((n, 1, hp, wp), dtype=torch.float32, device=src.device)
    elif out.numel() != n * hp * wp or not out.is_contiguous() or out.dtype != torch.float32:
        raise RuntimeError("unpack_u8: bad output buffer")
    _dev = _same_device(src, out)
    _launch(_dev, "unpack_u8", nat.lib().pmctf_unpack_u8, src.data_ptr(), out.data_ptr(), n, h0, w0, hp, wp)
    return out


def frame_sse(rec: torch.Tensor, orig_u8: torch.Tensor, sse: Optional[torch.Tensor] = None):
    """Exact per-plane sum of (round(clamp(rec,0,255)) - orig)^2 over the un-padded area (test_pMCTF_flex.py:300-310)."""
    rec = _chk(rec, "rec").contiguous()
    if not orig_u8.is_cuda or orig_u8.dtype != torch.uint8 or orig_u8.dim() != 3:
        raise RuntimeError("frame_sse expects the originals as a CUDA uint8 tensor [n,h0,w0]")
    orig_u8 = orig_u8.contiguous()
    n, h0, w0 = orig_u8.shape
    hp, wp = rec.size(-2), rec.size(-1)
    if rec.numel() != n * hp * wp:
        raise RuntimeError(f"frame_sse: {tuple(rec.shape)} does not hold {n} planes")
    if sse is None:
        sse = torch.zeros(n, dtype=torch.int64, device=rec.device)
    _dev = _same_device(rec, orig_u8, sse)
    _launch(_dev, "frame_sse", nat.lib().pmctf_frame_sse, rec.data_ptr(), orig_u8.data_ptr(), n, h0, w0, hp, wp, sse.data_ptr())
    return sse
