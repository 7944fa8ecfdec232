"""Training path (BASELINE.json configs[4], reference: train_pMCTF_L.py:161-251): differentiable primitives on the
B200 training kernels (csrc/pmctf_train.cu) and the reference's formulas composed from them.

Evaluation runs the fused tensor-core lifting step; as soon as autograd needs a gradient the modules switch to the
functions below, which mirror the reference line by line but call our own conv / warp kernels (forward AND backward) where
the reference calls cuDNN / grid_sample.  Element-wise glue (tanh, adds, scalings, straight-through round/clamp) is plain
torch.  CUDA tensors only, like the rest of the package."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _native as nat
from . import ops
from .layers.layers import ClampNoGradient, RoundNoGradient


def needs_grad(*ts) -> bool:
    """True when autograd is recording and a gradient can be asked for: an input tensor requires grad, or a module in
    training mode (model.train(), as train_pMCTF_L.py:118 sets it) has trainable parameters.  Evaluation (model.eval(),
    normally under torch.no_grad() like test_pMCTF_flex.py:130) always takes the fused kernels."""
    if not torch.is_grad_enabled():
        return False
    for t in ts:
        if isinstance(t, torch.Tensor):
            if t.requires_grad:
                return True
        elif isinstance(t, torch.nn.Module):
            if t.training and any(p.requires_grad for p in t.parameters()):
                return True
    return False


def _conv_raw(x, w, b):
    x = ops._chk(x, "x", 4).contiguous()
    w = ops._chk(w, "weight", 4).contiguous()
    N, cin, H, W = x.shape
    cout = w.size(0)
    if tuple(w.shape) != (cout, cin, 3, 3):
        raise RuntimeError(f"conv3x3: weight {tuple(w.shape)} does not match input channels {cin}")
    y = torch.empty((N, cout, H, W), dtype=torch.float32, device=x.device)
    ops._launch(ops._same_device(x, w, b), "conv3x3", nat.lib().pmctf_conv3x3, x.data_ptr(), w.data_ptr(),
                b.contiguous().data_ptr() if b is not None else None, y.data_ptr(), N, cin, cout, H, W)
    return y


class _Conv3x3(torch.autograd.Function):
    """nn.Conv2d(cin, cout, 3, padding=1) (layers.py:54-56) on pmctf_conv3x3 / pmctf_conv3x3_wgrad."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return _conv_raw(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _conv_raw(g, w.transpose(0, 1).flip(2, 3).contiguous(), None)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            N, cin, H, W = x.shape
            cout = w.size(0)
            gw = torch.zeros_like(w)
            gb = torch.zeros(cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            ops._launch(ops._same_device(x, g), "conv3x3_wgrad", nat.lib().pmctf_conv3x3_wgrad, x.contiguous().data_ptr(), g.data_ptr(),
                        gw.data_ptr(), gb.data_ptr() if gb is not None else None, N, cin, cout, H, W)
        return gx, gw, gb


def conv3x3(x, conv: torch.nn.Conv2d):
    return _Conv3x3.apply(x, conv.weight, conv.bias)


class _FlowWarp(torch.autograd.Function):
    """flow_warp (video_net.py:32-55) with its adjoint (pmctf_flow_warp_bwd)."""

    @staticmethod
    def forward(ctx, im, flow):
        im, flow = im.contiguous(), flow.contiguous()
        ctx.save_for_backward(im, flow)
        with torch.no_grad():
            return ops.flow_warp(im, flow)

    @staticmethod
    def backward(ctx, g):
        im, flow = ctx.saved_tensors
        g = g.contiguous()
        N, Cc, H, W = im.shape
        gim = torch.zeros_like(im) if ctx.needs_input_grad[0] else None
        gfl = torch.zeros_like(flow) if ctx.needs_input_grad[1] else None
        lx, ly = ops.linspace_table(W, im.device), ops.linspace_table(H, im.device)
        ops._launch(ops._same_device(g, im, flow), "flow_warp_bwd", nat.lib().pmctf_flow_warp_bwd, g.data_ptr(), im.data_ptr(),
                    flow.data_ptr(), lx.data_ptr(), ly.data_ptr(), gim.data_ptr() if gim is not None else None,
                    gfl.data_ptr() if gfl is not None else None, N, Cc, H, W, flow.size(0), 1.0)
        return gim, gfl


def flow_warp(im, flow):
    return _FlowWarp.apply(im, flow)


def _conv_fused(x, w, b, mode, aux=None, aux2=None, want_y2=False):
    """pmctf_conv3x3_fused: conv3x3 + the element-wise step behind it (modes: include/pmctf_b200.h)"""
    x, w = x.contiguous(), w.contiguous()
    N, cin, H, W = x.shape
    cout = w.size(0)
    y = torch.empty((N, cout, H, W), dtype=torch.float32, device=x.device)
    y2 = torch.empty_like(y) if want_y2 else None
    ops._launch(ops._same_device(x, w), "conv3x3_fused", nat.lib().pmctf_conv3x3_fused, x.data_ptr(), w.data_ptr(),
                b.contiguous().data_ptr() if b is not None else None, y.data_ptr(), y2.data_ptr() if y2 is not None else None,
                aux.contiguous().data_ptr() if aux is not None else None, aux2.contiguous().data_ptr() if aux2 is not None else None, mode,
                N, cin, cout, H, W)
    return (y, y2) if want_y2 else y


def _wgrad(x, g, w, has_bias):
    N, cin, H, W = x.shape
    cout = w.size(0)
    gw = torch.zeros_like(w)
    gb = torch.zeros(cout, dtype=torch.float32, device=x.device) if has_bias else None
    ops._launch(ops._same_device(x, g), "conv3x3_wgrad", nat.lib().pmctf_conv3x3_wgrad, x.contiguous().data_ptr(), g.contiguous().data_ptr(),
                gw.data_ptr(), gb.data_ptr() if gb is not None else None, N, cin, cout, H, W)
    return gw, gb


def _flip(w):
    return w.transpose(0, 1).flip(2, 3).contiguous()


class _PredictUpdate(torch.autograd.Function):
    """PredictUpdate.forward (lifting_1d.py:36-49) and its backward as ONE autograd node on fused kernels: every element-wise step
    over the 16-channel maps (two tanh, the residual add, the two tanh derivatives, the gradient sum of the residual branch) lives
    in the epilogue of the convolution before it.  4 launches forward, 8 backward (4 data gradients, 4 weight gradients); saved for
    backward: x, a1 = tanh(c1), a2 = tanh(c2), r = c1 + c3."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3, w4, b4):
        x = ops._chk(x, "x", 4).contiguous()
        c1, a1 = _conv_fused(x, w1, b1, 2, want_y2=True)
        a2 = _conv_fused(a1, w2, b2, 1)
        r = _conv_fused(a2, w3, b3, 3, aux=c1)
        y = _conv_fused(r, w4, b4, 0)
        ctx.save_for_backward(x, a1, a2, r, w1, w2, w3, w4)
        return y

    @staticmethod
    def backward(ctx, g):
        x, a1, a2, r, w1, w2, w3, w4 = ctx.saved_tensors
        g = g.contiguous()
        g_r = _conv_fused(g, _flip(w4), None, 0)
        gw4, gb4 = _wgrad(r, g, w4, True)
        g_c2 = _conv_fused(g_r, _flip(w3), None, 4, aux=a2)                    # through conv3, then tanh'(c2) = 1 - a2^2
        gw3, gb3 = _wgrad(a2, g_r, w3, True)
        g_c1 = _conv_fused(g_c2, _flip(w2), None, 5, aux=a1, aux2=g_r)          # through conv2 and tanh'(c1), plus the residual branch
        gw2, gb2 = _wgrad(a1, g_c2, w2, True)
        gx = _conv_fused(g_c1, _flip(w1), None, 0) if ctx.needs_input_grad[0] else None
        gw1, gb1 = _wgrad(x, g_c1, w1, True)
        return gx, gw1, gb1, gw2, gb2, gw3, gb3, gw4, gb4


FUSED_PU = True     # False: the un-fused composition below (one autograd node per convolution, element-wise steps in torch)


# ---- the reference's formulas on these primitives ---------------------------------------------------------------------
def predict_update(pu, x):
    """PredictUpdate.forward (lifting_1d.py:36-49)."""
    if FUSED_PU and all(c.bias is not None for c in (pu.conv1, pu.conv2, pu.conv3, pu.conv4)):
        return _PredictUpdate.apply(x, pu.conv1.weight, pu.conv1.bias, pu.conv2.weight, pu.conv2.bias, pu.conv3.weight, pu.conv3.bias,
                                    pu.conv4.weight, pu.conv4.bias)
    c1 = conv3x3(x, pu.conv1)
    a = torch.tanh(c1)
    a = torch.tanh(conv3x3(a, pu.conv2))
    return conv3x3(c1 + conv3x3(a, pu.conv3), pu.conv4)


def temporal_filter(tl, x, which: int):
    """TemporalLifting.predict_filter / update_filter (wavelet_transform_temporal_mctf.py:27-45)."""
    net, scale = (tl.P_t, tl.scale_p) if which == 0 else (tl.U_t, tl.scale_u)
    t = predict_update(net, x) * tl.scale
    if not tl.lossy:
        return x + RoundNoGradient.apply(t)
    return (x + t) * float(scale.detach())


def chroma_mv(mv):
    """bilineardownsacling(mv) / 2 (video_net.py:66-71): bilinear x0.5 with align_corners=False is the 2x2 box mean."""
    return F.avg_pool2d(mv, 2) / 2


def forward_mctf(m, ref, cur, mv, stage_idx=0, mv_down=False):
    """pMCTF.forward_MCTF (pMCTF_L.py:297-312)."""
    tl = m.temporal_filtering[min(m.num_me_stages - 1, stage_idx)]
    if mv_down:
        mv = chroma_mv(mv)
    shp = ref.shape
    ref4, cur4 = ref.reshape(-1, 1, shp[-2], shp[-1]), cur.reshape(-1, 1, shp[-2], shp[-1])
    pred = flow_warp(ref4, mv)
    if not m.lossy:
        pred = RoundNoGradient.apply(pred)
    pred = temporal_filter(tl, pred, 0)
    H_t = cur4 - pred
    inv = flow_warp(H_t, -mv)
    if not m.lossy:
        inv = RoundNoGradient.apply(inv)
    inv = temporal_filter(tl, inv, 1)
    L_t = ref4 + inv
    return L_t.reshape(shp), H_t.reshape(shp), pred.reshape(shp), inv.reshape(shp)


def inverse_mctf(m, L_t, H_t, mv, downscale=False, stage_idx=0):
    """pMCTF.inverse_MCTF (pMCTF_L.py:314-330)."""
    tl = m.temporal_filtering[min(m.num_me_stages - 1, stage_idx)]
    if downscale:
        mv = chroma_mv(mv)
    shp = L_t.shape
    L4, H4 = L_t.reshape(-1, 1, shp[-2], shp[-1]), H_t.reshape(-1, 1, shp[-2], shp[-1])
    inv = flow_warp(H4, -mv)
    if not m.lossy:
        inv = RoundNoGradient.apply(inv)
    ref = L4 - temporal_filter(tl, inv, 1)
    pred = flow_warp(ref, mv)
    if not m.lossy:
        pred = RoundNoGradient.apply(pred)
    cur = H4 + temporal_filter(tl, pred, 0)
    return ref.reshape(shp), cur.reshape(shp)


def _skip(conv, x):
    """(3,1) conv on the row-reflect-padded input (lifting_1d.py:91,105-106) as three shifted views."""
    xp = torch.cat([x[:, :, 1:2], x, x[:, :, -2:-1]], dim=2)
    w = conv.weight.reshape(3)
    return w[0] * xp[:, :, :-2] + w[1] * xp[:, :, 1:-1] + w[2] * xp[:, :, 2:] + conv.bias.reshape(1, 1, 1, 1)


def _lift_term(lift, conv, pu, x):
    """skip + 0.1 * 256 * PU(skip / 256) (lifting_1d.py:105-111)."""
    skip = _skip(conv, x)
    t = predict_update(pu, skip / lift.dynamic_range) * lift.dynamic_range * 0.1
    if not lift.lossy:
        return RoundNoGradient.apply(skip + t)
    return skip + t


def iwave1d_forward(lift, x):
    """iWave1D.forward_lift (lifting_1d.py:103-145)."""
    x_e, x_o = x[:, :, ::2, :], x[:, :, 1::2, :]
    x_o = x_o + _lift_term(lift, lift.conv_P1, lift.P_1, x_e)
    x_e = x_e + _lift_term(lift, lift.conv_U1, lift.U_1, x_o)
    x_o = x_o + _lift_term(lift, lift.conv_P2, lift.P_2, x_e)
    x_e = x_e + _lift_term(lift, lift.conv_U2, lift.U_2, x_o)
    if lift.lossy:
        x_e, x_o = x_e * float(lift.scale_l.detach()), x_o * float(lift.scale_h.detach())
    return x_e, x_o


def iwave1d_backward(lift, l, h):
    """iWave1D.backward_lift (lifting_1d.py:147-189)."""
    if lift.lossy:
        l, h = l / float(lift.scale_l.detach()), h / float(lift.scale_h.detach())
    l = l - _lift_term(lift, lift.conv_U2, lift.U_2, h)
    h = h - _lift_term(lift, lift.conv_P2, lift.P_2, l)
    l = l - _lift_term(lift, lift.conv_U1, lift.U_1, h)
    h = h - _lift_term(lift, lift.conv_P1, lift.P_1, l)
    n, c, hh, w = l.shape
    x = torch.stack([l, h], dim=3).reshape(n, c, 2 * hh, w)  # merge (lifting_1d.py:16-22)
    return x


def lift2d_forward(wt, x):
    """LiftingScheme2D.forward_lift_2d (wavelet_transform.py:25-43)."""
    l, h = iwave1d_forward(wt.lift_h, x)
    ll, lh = iwave1d_forward(wt.lift_v, l.permute(0, 1, 3, 2).contiguous())
    hl, hh = iwave1d_forward(wt.lift_v, h.permute(0, 1, 3, 2).contiguous())
    p = lambda t: t.permute(0, 1, 3, 2)  # noqa: E731
    return {"ll": p(ll), "lh": p(lh), "hl": p(hl), "hh": p(hh), "l": l.permute(0, 1, 3, 2), "h": h.permute(0, 1, 3, 2)}


def lift2d_backward(wt, sb):
    """LiftingScheme2D.backward_lift_2d (wavelet_transform.py:45-57)."""
    p = lambda t: t.permute(0, 1, 3, 2).contiguous()  # noqa: E731
    l = iwave1d_backward(wt.lift_v, p(sb["ll"]), p(sb["lh"]))
    h = iwave1d_backward(wt.lift_v, p(sb["hl"]), p(sb["hh"]))
    return iwave1d_backward(wt.lift_h, l.permute(0, 1, 3, 2).contiguous(), h.permute(0, 1, 3, 2).contiguous())


def quantize(s, q, clip, lossy=True, do_round=True):
    """quantize_subband (+ RoundNoGradient): straight-through clamp / round (pWave.py:184-189, layers.py:71-92)."""
    v = s * q if lossy else s
    v = ClampNoGradient.apply(v, -clip, clip)
    return RoundNoGradient.apply(v) if do_round else v
