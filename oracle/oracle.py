"""ctypes front-end of the CPU oracle (oracle/pmctf_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under learned-pmctf_b200/ imports this module.

The functions mirror the reference's call surface for the hot path on numpy
float32 arrays; `sd` arguments are state-dict style mappings name -> ndarray with the reference's
own key names (e.g. "conv1.weight", "P_1.conv2.bias", "conv_P1.weight").
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpmctf_oracle.so")

_f32p = C.POINTER(C.c_float)


class _PU(C.Structure):
    _fields_ = [(n, _f32p) for n in ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")]


class _IWave(C.Structure):
    _fields_ = [("tap", (C.c_float * 3) * 4), ("bias", C.c_float * 4), ("pu", _PU * 4),
                ("scale_l", C.c_float), ("scale_h", C.c_float), ("dynamic_range", C.c_float), ("lossy", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, portable x86-64 flags)."""
    src = os.path.join(_HERE, "pmctf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def set_threads(n: int) -> None:
    """Number of OpenMP threads the oracle uses (torchrun exports OMP_NUM_THREADS=1, which would silently serialise it)."""
    lib()
    C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))


def set_conv_mode(mode: str) -> None:
    """'ffma': conv2/conv3 as sequential fp32 FMA chains; 'tensor': exact fixed-point (see pmctf_oracle.c)."""
    lib().orc_set_conv_mode({"ffma": 0, "tensor": 1}[mode])


def get_conv_mode() -> str:
    return "tensor" if lib().orc_get_conv_mode() == 1 else "ffma"


def _a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(x: np.ndarray):
    return x.ctypes.data_as(_f32p)


# lifting_1d.py:57-58 (bior4.4); scale_l/scale_h are fp32 0-dim tensors outside the state_dict (:98-101)
SCALE_L = float(np.float32(1.149604398860241))
SCALE_H = float(np.float32(0.869864451624781))
# wavelet_transform_temporal_mctf.py:24-25
SCALE_P = float(np.float32(1 / math.sqrt(2)))
SCALE_U = 0.5


class PU:
    """Holds one PredictUpdate's weights alive for the C side (lifting_1d.py:25-49)."""

    def __init__(self, sd, prefix=""):
        self.arrs = [_a(sd[f"{prefix}conv{i}.{k}"]) for i in (1, 2, 3, 4) for k in ("weight", "bias")]
        assert self.arrs[0].shape == (16, 1, 3, 3) and self.arrs[2].shape == (16, 16, 3, 3) and self.arrs[6].shape == (1, 16, 3, 3)
        self.c = _PU(*[_p(a) for a in self.arrs])


class IWave:
    """iWave1D parameters (lifting_1d.py:52-101)."""

    def __init__(self, sd, prefix="", lossy=True, dynamic_range=256.0):
        self.pus = [PU(sd, f"{prefix}{n}.") for n in ("P_1", "U_1", "P_2", "U_2")]
        c = _IWave()
        for i, n in enumerate(("conv_P1", "conv_U1", "conv_P2", "conv_U2")):
            w = _a(sd[f"{prefix}{n}.weight"]).reshape(3)
            for j in range(3):
                c.tap[i][j] = float(w[j])
            c.bias[i] = float(_a(sd[f"{prefix}{n}.bias"]).reshape(1)[0])
            c.pu[i] = self.pus[i].c
        c.scale_l, c.scale_h, c.dynamic_range, c.lossy = SCALE_L, SCALE_H, float(dynamic_range), int(lossy)  # 2 ** bitdepth, lifting_1d.py:62
        self.c = c
        self.lossy = lossy


def linspace(n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().orc_linspace(_p(out), C.c_int(n))
    return out


def tanh(x):
    x = _a(x)
    y = np.empty_like(x)
    lib().orc_tanh(_p(x), _p(y), C.c_long(x.size))
    return y


def flow_warp(im, flow, sign=1.0, lin_x=None, lin_y=None, round_out=False):
    """video_net.py:32-55.  im [N,C,H,W], flow [1 or N,2,H,W]."""
    im, flow = _a(im), _a(flow)
    N, Cc, H, W = im.shape
    lx = _a(lin_x) if lin_x is not None else linspace(W)
    ly = _a(lin_y) if lin_y is not None else linspace(H)
    out = np.empty_like(im)
    lib().orc_flow_warp(_p(im), _p(flow), _p(lx), _p(ly), _p(out), N, Cc, H, W, flow.shape[0],
                        C.c_float(sign), int(round_out))
    return out


def chroma_mv_down(mv):
    """bilineardownsacling(mv) / 2: video_net.py:66-71, pMCTF_L.py:317."""
    mv = _a(mv)
    N, two, H, W = mv.shape
    out = np.empty((N, 2, H // 2, W // 2), np.float32)
    lib().orc_chroma_mv_down(_p(mv), _p(out), N, H, W)
    return out


def predict_update(x, pu: PU, in_mul=1.0):
    """PredictUpdate.forward on [N,1,H,W] (lifting_1d.py:36-49)."""
    x = _a(x)
    N, _, H, W = x.shape
    out = np.empty_like(x)
    lib().orc_predict_update(_p(x), C.byref(pu.c), _p(out), N, H, W, C.c_float(in_mul))
    return out


def temporal_filter(x, pu: PU, scale, lossy=True):
    """TemporalLifting.predict_filter / update_filter (wavelet_transform_temporal_mctf.py:27-45)."""
    x = _a(x)
    N, _, H, W = x.shape
    out = np.empty_like(x)
    lib().orc_temporal_filter(_p(x), C.byref(pu.c), C.c_float(scale), int(lossy), _p(out), N, H, W)
    return out


def forward_mctf(ref, cur, mv, P_t: PU, U_t: PU, lossy=True, lin_x=None, lin_y=None):
    """pMCTF.forward_MCTF (pMCTF_L.py:297-312) -> (L_t, H_t, pred, inv)."""
    ref, cur, mv = _a(ref), _a(cur), _a(mv)
    N, _, H, W = ref.shape
    lx = _a(lin_x) if lin_x is not None else linspace(W)
    ly = _a(lin_y) if lin_y is not None else linspace(H)
    L, Hh, pred, inv = (np.empty_like(ref) for _ in range(4))
    lib().orc_forward_mctf(_p(ref), _p(cur), _p(mv), mv.shape[0], _p(lx), _p(ly), C.byref(P_t.c), C.byref(U_t.c),
                           C.c_float(SCALE_P), C.c_float(SCALE_U), int(lossy), _p(L), _p(Hh), _p(pred), _p(inv), N, H, W)
    return L, Hh, pred, inv


def inverse_mctf(L, Hh, mv, P_t: PU, U_t: PU, lossy=True, downscale=False, lin_x=None, lin_y=None):
    """pMCTF.inverse_MCTF (pMCTF_L.py:314-330) -> (ref, cur)."""
    L, Hh, mv = _a(L), _a(Hh), _a(mv)
    if downscale:
        mv = chroma_mv_down(mv)
    N, _, H, W = L.shape
    lx = _a(lin_x) if lin_x is not None else linspace(W)
    ly = _a(lin_y) if lin_y is not None else linspace(H)
    ref, cur = np.empty_like(L), np.empty_like(L)
    lib().orc_inverse_mctf(_p(L), _p(Hh), _p(mv), mv.shape[0], _p(lx), _p(ly), C.byref(P_t.c), C.byref(U_t.c),
                           C.c_float(SCALE_P), C.c_float(SCALE_U), int(lossy), _p(ref), _p(cur), N, H, W)
    return ref, cur


def iwave1d_forward(x, w: IWave):
    """iWave1D.forward_lift (lifting_1d.py:103-145): [N,1,H,W] -> l,h [N,1,H/2,W]."""
    x = _a(x)
    N, _, H, W = x.shape
    l, h = np.empty((N, 1, H // 2, W), np.float32), np.empty((N, 1, H // 2, W), np.float32)
    lib().orc_iwave1d_forward(_p(x), C.byref(w.c), _p(l), _p(h), N, H, W)
    return l, h


def iwave1d_backward(l, h, w: IWave):
    """iWave1D.backward_lift (lifting_1d.py:147-189)."""
    l, h = _a(l), _a(h)
    N, _, h2, W = l.shape
    x = np.empty((N, 1, 2 * h2, W), np.float32)
    lib().orc_iwave1d_backward(_p(l), _p(h), C.byref(w.c), _p(x), N, 2 * h2, W)
    return x


def lift2d_forward(x, w: IWave):
    """LiftingScheme2D.forward_lift_2d (wavelet_transform.py:25-43) -> dict ll, lh, hl, hh (+ l, h row-pass)."""
    x = _a(x)
    N, _, H, W = x.shape
    q = lambda: np.empty((N, 1, H // 2, W // 2), np.float32)  # noqa: E731
    ll, lh, hl, hh = q(), q(), q(), q()
    l, h = np.empty((N, 1, H // 2, W), np.float32), np.empty((N, 1, H // 2, W), np.float32)
    lib().orc_lift2d_forward(_p(x), C.byref(w.c), _p(ll), _p(lh), _p(hl), _p(hh), _p(l), _p(h), N, H, W)
    return {"ll": ll, "lh": lh, "hl": hl, "hh": hh, "l": l, "h": h}


def lift2d_backward(sb, w: IWave):
    """LiftingScheme2D.backward_lift_2d (wavelet_transform.py:45-57)."""
    ll, lh, hl, hh = (_a(sb[k]) for k in ("ll", "lh", "hl", "hh"))
    N, _, h2, w2 = ll.shape
    x = np.empty((N, 1, 2 * h2, 2 * w2), np.float32)
    lib().orc_lift2d_backward(_p(ll), _p(lh), _p(hl), _p(hh), C.byref(w.c), _p(x), N, 2 * h2, 2 * w2)
    return x


def pwave_encode(x, w: IWave, levels=4):
    """pWave.encode (pWave.py:139-148)."""
    sub, ll = {}, _a(x)
    for lvl in range(levels):
        sub[lvl] = lift2d_forward(ll, w)
        ll = sub[lvl]["ll"]
    return sub


def pwave_decode(sub, w: IWave, levels=4):
    """pWave.decode (pWave.py:150-157); like the reference it writes `ll` back into `sub`."""
    y = None
    for lvl in range(levels - 1, -1, -1):
        y = lift2d_backward(sub[lvl], w)
        if lvl > 0:
            sub[lvl - 1]["ll"] = y
    return y


def quantize(s, q, clip=8192.0, lossy=True, do_round=True):
    """round(clamp(s*q, +-clip)) (pWave.py:184-189,256-257,337; layers.py:71-92)."""
    s = _a(s)
    out = np.empty_like(s)
    lib().orc_quantize(_p(s), C.c_float(q), C.c_float(clip), int(lossy), int(do_round), _p(out), C.c_long(s.size))
    return out


def dequantize(s_hat, q, lossy=True):
    """s_hat / q (pWave.py:191-202)."""
    s_hat = _a(s_hat)
    out = np.empty_like(s_hat)
    lib().orc_dequantize(_p(s_hat), C.c_float(q), int(lossy), _p(out), C.c_long(s_hat.size))
    return out


def spatial_wavelet_dec(x, w: IWave, q, q_ll, levels=4, lossy=True):
    """pWave.spatial_wavelet_dec WITHOUT the PostProcess net (pWave.py:314-349, stop before :347):
    encode -> round(clamp(s*q)) on every band -> dequantise -> decode.  Returns (x_hat, symbols)."""
    y = pwave_encode(x, w, levels)
    hat = {lvl: {} for lvl in range(levels)}
    hat[levels - 1]["ll"] = quantize(y[levels - 1]["ll"], q_ll, lossy=lossy)
    for lvl in range(levels - 1, -1, -1):
        for b in ("lh", "hl", "hh"):
            hat[lvl][b] = quantize(y[lvl][b], q, lossy=lossy)
    rec = {lvl: {b: dequantize(v, q_ll if b == "ll" else q, lossy) for b, v in hat[lvl].items()} for lvl in hat}
    return pwave_decode(rec, w, levels), hat


def one_q_scale(q_scale, q_index, qp_num=21):
    """get_one_q_scale (pWave.py:209-215 / pMCTF_L.py:195-200) in float32 numpy (log/exp may differ
    from torch's Sleef by 1 ulp; parity tests take q from the golden file instead)."""
    q_scale = _a(q_scale).reshape(2)
    lo, hi = np.log(q_scale[0]), np.log(q_scale[1])
    step = np.float32((hi - lo) / np.float32(qp_num - 1))
    return float(np.exp(np.float32(lo + step * np.float32(q_index))))


# ---------------------------------------------------------------------------------------------------
# Host-side pieces of the GOP loop (numpy: byte / integer arithmetic)
def unpack_u8(src, hp, wp):
    """np_image_to_tensor + zero pad bottom/right (test_pMCTF_flex.py:151-192)."""
    src = np.asarray(src, np.uint8)
    out = np.zeros(src.shape[:-2] + (hp, wp), np.float32)
    out[..., : src.shape[-2], : src.shape[-1]] = src
    return out


def frame_sse(rec, orig_u8):
    """Per-plane sum of (round(clamp(rec,0,255)) - orig)^2 over the un-padded area (test_pMCTF_flex.py:300-310)."""
    orig = np.asarray(orig_u8)
    h0, w0 = orig.shape[-2:]
    r = np.rint(np.clip(_a(rec), 0, 255))[..., :h0, :w0].astype(np.int64)
    return ((r.reshape(orig.shape) - orig.astype(np.int64)) ** 2).sum((-2, -1))


def symbol_stats(hat):
    """(sum |sym|, #nonzero) per plane over every coded band of one spatial_wavelet_dec call."""
    n = next(iter(hat[0].values())).shape[0]
    st = np.zeros((n, 2), np.int64)
    for lvl in hat:
        for v in hat[lvl].values():
            a = np.abs(v.reshape(n, -1)).astype(np.int64)
            st[:, 0] += a.sum(1)
            st[:, 1] += (a != 0).sum(1)
    return st


def code_gop(Y, C, mvs, temporal, hp_w: IWave, lp_w: IWave, q_hp, q_lp, num_me_stages=4, trace=None):
    """The reference's GOP loop restricted to the hot path, one pair at a time in its own order
    (test_pMCTF_flex.py:131-291): stage s pairs frame g*2^(s+1) with +2^s; H frames are coded with the hp
    transform (step pair q_hp[s]), the final L with the lp transform; temporal decoding in reverse.
    Y [G,1,H,W]; C [G,2,1,h,w]; mvs[s] [pairs,2,H,W]; temporal[i] = (P_t, U_t).
    -> (rec_Y, rec_C, sym_stats int64 [G,2]).  `trace` (a dict) receives the temporal subbands before coding:
    trace["H"][frame] = (H_y, H_c) of every high-pass frame, trace["L"] = (L_y, L_c) of the final low-pass frame,
    trace["sym"][frame] = (luma symbols, chroma symbols) as spatial_wavelet_dec returns them."""
    Y, C = _a(Y).copy(), _a(C).copy()
    G = Y.shape[0]
    S = int(round(math.log2(G)))
    coded = {}
    sym = np.zeros((G, 2), np.int64)
    fy = {i: Y[i:i + 1] for i in range(G)}
    fc = {i: C[i] for i in range(G)}
    for s in range(S):
        step = 2 ** s
        P, U = temporal[min(num_me_stages - 1, s)]
        for g in range(G // (2 * step)):
            r, c = g * 2 * step, g * 2 * step + step
            mv = _a(mvs[s][g:g + 1])
            L, Hh, _, _ = forward_mctf(fy[r], fy[c], mv, P, U)
            Lc, Hc, _, _ = forward_mctf(fc[r], fc[c], chroma_mv_down(mv), P, U)
            q, qll = q_hp[s]
            Hh_hat, hat = spatial_wavelet_dec(Hh, hp_w, q, qll)
            Hc_hat, hatc = spatial_wavelet_dec(Hc, hp_w, q, qll)
            sym[c] += symbol_stats(hat)[0] + symbol_stats(hatc).sum(0)
            fy[r], fc[r] = L, Lc
            coded[c] = (Hh_hat, Hc_hat, mv)
            if trace is not None:
                trace.setdefault("H", {})[c] = (Hh, Hc)
                trace.setdefault("sym", {})[c] = (hat, hatc)
    q, qll = q_lp
    if trace is not None:
        trace["L"] = (fy[0], fc[0])
    L_hat, hat = spatial_wavelet_dec(fy[0], lp_w, q, qll)
    Lc_hat, hatc = spatial_wavelet_dec(fc[0], lp_w, q, qll)
    sym[0] += symbol_stats(hat)[0] + symbol_stats(hatc).sum(0)
    if trace is not None:
        trace.setdefault("sym", {})[0] = (hat, hatc)
    ry, rc = {0: L_hat}, {0: Lc_hat}
    for s in range(S - 1, -1, -1):
        step = 2 ** s
        P, U = temporal[min(num_me_stages - 1, s)]
        for g in reversed(range(G // (2 * step))):
            r, c = g * 2 * step, g * 2 * step + step
            Hh_hat, Hc_hat, mv = coded[c]
            ry[r], ry[c] = inverse_mctf(ry[r], Hh_hat, mv, P, U)
            rc[r], rc[c] = inverse_mctf(rc[r], Hc_hat, mv, P, U, downscale=True)
    return np.concatenate([ry[i] for i in range(G)]), np.stack([rc[i] for i in range(G)]), sym


# ---------------------------------------------------------------------------------------------------
# PostProcess (postprocessing.py:20-44) -- fp32 tolerance oracle of the section 8f row
class _PostProcessC(C.Structure):
    _fields_ = [("conv1_w", C.c_void_p), ("conv1_b", C.c_void_p), ("res_w", C.c_void_p * 12), ("res_b", C.c_void_p * 12),
                ("conv2_w", C.c_void_p), ("conv2_b", C.c_void_p), ("conv3_w", C.c_void_p), ("conv3_b", C.c_void_p)]


class PostProcess:
    """Weights of a PostProcess module from a state_dict-like mapping (keys conv1.weight, resBlocks.0.conv1.weight, ...)."""

    def __init__(self, sd, prefix=""):
        g = lambda k: _a(sd[prefix + k])  # noqa: E731
        self.keep = {k: g(k) for k in ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"] +
                     [f"resBlocks.{i}.conv{j}.{t}" for i in range(6) for j in (1, 2) for t in ("weight", "bias")]}
        c = _PostProcessC()
        c.conv1_w, c.conv1_b = self.keep["conv1.weight"].ctypes.data, self.keep["conv1.bias"].ctypes.data
        c.conv2_w, c.conv2_b = self.keep["conv2.weight"].ctypes.data, self.keep["conv2.bias"].ctypes.data
        c.conv3_w, c.conv3_b = self.keep["conv3.weight"].ctypes.data, self.keep["conv3.bias"].ctypes.data
        for i in range(6):
            for j in (1, 2):
                c.res_w[2 * i + j - 1] = self.keep[f"resBlocks.{i}.conv{j}.weight"].ctypes.data
                c.res_b[2 * i + j - 1] = self.keep[f"resBlocks.{i}.conv{j}.bias"].ctypes.data
        self.c = c


def postprocess(x, pp: PostProcess, in_mul=1.0 / 256.0, out_mul=256.0):
    """dequantModule(x / 256) * 256 (pWave.py:300) on [N,1,H,W]."""
    x = _a(x)
    N, _, H, W = x.shape
    y = np.empty_like(x)
    lib().orc_postprocess(_p(x), C.byref(pp.c), C.c_float(in_mul), C.c_float(out_mul), _p(y), N, H, W)
    return y


def conv3x3(x, w, b, slope=1.0, res=None):
    """nn.Conv2d(3x3, padding=1) (+ residual, + LeakyReLU) on ONE planar map [ci,H,W] -> [co,H,W]."""
    x, w, b = _a(x), _a(w), _a(b)
    ci, H, W = x.shape
    co = w.shape[0]
    out = np.empty((co, H, W), np.float32)
    r = _a(res) if res is not None else None
    lib().orc_conv3x3(_p(x), ci, _p(w), _p(b), co, _p(out), H, W, C.c_float(slope), _p(r) if r is not None else None)
    return out


def llar_encode(sd, yq):
    """The LL band's autoregressive model, coefficient by coefficient, under the fp32 contract stated at orc_llar_encode
    (pmctf_oracle.c): reference pMCTF/layers/context_fusion.py:56-204 as driven by pMCTF/models/pWave.py:531-553.
    sd: state_dict of a ContextFusionSubband (numpy arrays; keys maskedConv1.*, residualBlocks.{0,1}.conv{1,2}.*, maskedConv2.*,
    convs.{0,1,2}.*); yq [H,W] fp32 quantised band.  -> (scales, means, symbols, reconstruction), each [H,W] fp32."""
    f = lambda k: np.ascontiguousarray(sd[k], dtype=np.float32)   # noqa: E731
    names = ["residualBlocks.0.conv1", "residualBlocks.0.conv2", "residualBlocks.1.conv1", "residualBlocks.1.conv2", "maskedConv2"]
    w128 = np.ascontiguousarray(np.stack([f(n + ".weight") for n in names]))
    b128 = np.ascontiguousarray(np.stack([f(n + ".bias") for n in names]))
    w1 = np.ascontiguousarray(np.stack([f("convs.0.weight").reshape(128, 128), f("convs.1.weight").reshape(128, 128)]))
    b1 = np.ascontiguousarray(np.stack([f("convs.0.bias"), f("convs.1.bias")]))
    wout, bout = f("convs.2.weight").reshape(2, 128).copy(), f("convs.2.bias")
    w_in, b_in = f("maskedConv1.weight").reshape(128, 9).copy(), f("maskedConv1.bias")
    assert w128.shape == (5, 128, 128, 3, 3) and w_in.shape == (128, 9) and wout.shape == (2, 128)
    yq = np.ascontiguousarray(yq, dtype=np.float32)
    H, W = yq.shape
    outs = [np.empty((H, W), dtype=np.float32) for _ in range(4)]
    P = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    fn = lib().orc_llar_encode
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_int] + [C.c_void_p] * 4
    fn(P(w_in), P(b_in), P(w128), P(b128), P(w1), P(b1), P(wout), P(bout), P(yq), H, W, *[P(o) for o in outs])
    return tuple(outs)
