"""BASELINE ONLY -- the hot path written with stock PyTorch ops (F.conv2d -> cuDNN, F.grid_sample, ATen element-wise), i.e.
what the reference itself dispatches to on a GPU ("the bar is the stock ATen/cuDNN library path").  It follows
the reference's formulas (pMCTF_L.py:297-330, lifting_1d.py:36-189, wavelet_transform.py:25-57, pWave.py:314-349) and takes
its weights from a learned_pmctf_b200.pMCTF instance.  Used only by `bench.py --torch-baseline` to put a number beside ours;
nothing in the product imports it."""
import torch
import torch.nn.functional as F


def warp(im, flow):
    N, C, H, W = im.shape
    gx = torch.linspace(-1, 1, W, device=im.device).view(1, 1, 1, W).expand(N, 1, H, W)
    gy = torch.linspace(-1, 1, H, device=im.device).view(1, 1, H, 1).expand(N, 1, H, W)
    grid = torch.cat([gx, gy], 1) + torch.cat([flow[:, 0:1] / ((W - 1) / 2), flow[:, 1:2] / ((H - 1) / 2)], 1)
    return F.grid_sample(im, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


def pu(p, x):
    c1 = F.conv2d(x, p.conv1.weight, p.conv1.bias, padding=1)
    a = torch.tanh(c1)
    a = torch.tanh(F.conv2d(a, p.conv2.weight, p.conv2.bias, padding=1))
    return F.conv2d(c1 + F.conv2d(a, p.conv3.weight, p.conv3.bias, padding=1), p.conv4.weight, p.conv4.bias, padding=1)


def tfilter(net, x, scale):
    return (x + pu(net, x) * 0.1) * scale


def forward_mctf(tl, ref, cur, mv):
    mvn = mv.repeat_interleave(ref.size(0) // mv.size(0), 0) if mv.size(0) != ref.size(0) else mv
    Hh = cur - tfilter(tl.P_t, warp(ref, mvn), float(tl.scale_p))
    return ref + tfilter(tl.U_t, warp(Hh, -mvn), float(tl.scale_u)), Hh


def inverse_mctf(tl, L, Hh, mv):
    mvn = mv.repeat_interleave(L.size(0) // mv.size(0), 0) if mv.size(0) != L.size(0) else mv
    ref = L - tfilter(tl.U_t, warp(Hh, -mvn), float(tl.scale_u))
    return ref, Hh + tfilter(tl.P_t, warp(ref, mvn), float(tl.scale_p))


def _term(conv, p, x):
    skip = F.conv2d(torch.cat([x[:, :, 1:2], x, x[:, :, -2:-1]], 2), conv.weight, conv.bias)
    return skip + pu(p, skip / 256.0) * 256.0 * 0.1


def fwd1d(w, x):
    e, o = x[:, :, ::2], x[:, :, 1::2]
    o = o + _term(w.conv_P1, w.P_1, e)
    e = e + _term(w.conv_U1, w.U_1, o)
    o = o + _term(w.conv_P2, w.P_2, e)
    e = e + _term(w.conv_U2, w.U_2, o)
    return e * float(w.scale_l), o * float(w.scale_h)


def bwd1d(w, l, h):
    l, h = l / float(w.scale_l), h / float(w.scale_h)
    l = l - _term(w.conv_U2, w.U_2, h)
    h = h - _term(w.conv_P2, w.P_2, l)
    l = l - _term(w.conv_U1, w.U_1, h)
    h = h - _term(w.conv_P1, w.P_1, l)
    x = torch.zeros((l.size(0), 1, 2 * l.size(2), l.size(3)), device=l.device)   # merge (lifting_1d.py:16-22)
    x[:, :, ::2], x[:, :, 1::2] = l, h
    return x


def spatial_wavelet_dec(coder, x, q, qll, levels=4):
    w = coder.wavelet_transform.lift_h
    T = lambda t: t.permute(0, 1, 3, 2)  # noqa: E731
    ll, bands = x, []
    for _ in range(levels):
        l, h = fwd1d(w, ll)
        a, b = fwd1d(w, T(l))
        c, d = fwd1d(w, T(h))
        bands.append((T(b), T(c), T(d)))
        ll = T(a)
    rq = lambda v, s: torch.round((v * s).clamp(-8192, 8192))  # noqa: E731
    y = rq(ll, qll) / qll
    for lvl in range(levels - 1, -1, -1):
        lh, hl, hh = (rq(v, q) / q for v in bands[lvl])
        y = bwd1d(w, T(bwd1d(w, T(y), T(lh))), T(bwd1d(w, T(hl), T(hh))))
    return y


@torch.no_grad()
def code_gop(model, codec, Y, C, mvs):
    """Same schedule as learned_pmctf_b200.gop.GopCodec.code_gop (without the statistics)."""
    S = codec.stages

    def chroma_mv(mv):
        return F.interpolate(mv, scale_factor=0.5, mode="bilinear", align_corners=False) / 2

    Ly, Lc, Hs = Y, C.reshape(C.size(0), 2, 1, C.size(-2), C.size(-1)), []
    for s in range(S):
        tl = model.temporal_filtering[min(model.num_me_stages - 1, s)]
        L2y, Hy = forward_mctf(tl, Ly[0::2], Ly[1::2], mvs[s])
        cr, cc = Lc[0::2].reshape(-1, 1, *Lc.shape[-2:]), Lc[1::2].reshape(-1, 1, *Lc.shape[-2:])
        L2c, Hc = forward_mctf(tl, cr, cc, chroma_mv(mvs[s]))
        Hs.append((Hy, Hc))
        Ly, Lc = L2y, L2c.reshape(-1, 2, 1, *Lc.shape[-2:])
    Hhat = []
    for s, (Hy, Hc) in enumerate(Hs):
        q, qll = codec.q_pair("hp", s)
        Hhat.append((spatial_wavelet_dec(model.hp_coder, Hy, q, qll), spatial_wavelet_dec(model.hp_coder, Hc, q, qll)))
    q, qll = codec.q_pair("lp", 0)
    Ly = spatial_wavelet_dec(model.lp_coder, Ly, q, qll)
    Lc = spatial_wavelet_dec(model.lp_coder, Lc.reshape(-1, 1, *Lc.shape[-2:]), q, qll)
    for s in range(S - 1, -1, -1):
        tl = model.temporal_filtering[min(model.num_me_stages - 1, s)]
        Hy, Hc = Hhat[s]
        r, c = inverse_mctf(tl, Ly, Hy, mvs[s])
        Ly = torch.stack([r, c], 1).reshape(-1, 1, *r.shape[-2:])
        r, c = inverse_mctf(tl, Lc, Hc, chroma_mv(mvs[s]))
        n = r.size(0) // 2
        Lc = torch.stack([r.reshape(n, 2, *r.shape[-2:]), c.reshape(n, 2, *r.shape[-2:])], 1).reshape(-1, 1, *r.shape[-2:])
    return Ly, Lc.reshape(-1, 2, 1, *Lc.shape[-2:])


class _StockCoder:
    def __init__(self, coder):
        self.c = coder

    def q_pair(self, *a, **k):
        return self.c.q_pair(*a, **k)

    def spatial_wavelet_dec(self, x, q, qll, post_process=False, return_symbols=False):
        """Differentiable version (straight-through round / clamp) of spatial_wavelet_dec above."""
        w = self.c.wavelet_transform.lift_h
        T = lambda t: t.permute(0, 1, 3, 2)  # noqa: E731
        ll, bands = x, []
        for _ in range(4):
            l, h = fwd1d(w, ll)
            a, b = fwd1d(w, T(l))
            c, d = fwd1d(w, T(h))
            bands.append((T(b), T(c), T(d)))
            ll = T(a)
        ste = lambda v: v + (torch.round(v.clamp(-8192, 8192)) - v).detach()  # noqa: E731
        hat = {3: {"ll": ste(ll * qll)}}
        for lvl in range(4):
            hat.setdefault(lvl, {}).update({k: ste(v * q) for k, v in zip(("lh", "hl", "hh"), bands[lvl])})
        y = hat[3]["ll"] / qll
        for lvl in range(3, -1, -1):
            y = bwd1d(w, T(bwd1d(w, T(y), T(hat[lvl]["lh"] / q))), T(bwd1d(w, T(hat[lvl]["hl"] / q), T(hat[lvl]["hh"] / q))))
        return (y, hat) if return_symbols else y


class StockModel:
    """Adapter with the method surface gop.training_loss_hot_path uses, on stock torch ops (baseline for the training step)."""

    def __init__(self, model):
        self.m = model
        self.num_me_stages = model.num_me_stages
        self.hp_coder, self.lp_coder = _StockCoder(model.hp_coder), _StockCoder(model.lp_coder)

    def hp_qp_scale(self, *a):
        return self.m.hp_qp_scale(*a)

    def forward_MCTF(self, ref, cur, mv, stage_idx=0):
        L, Hh = forward_mctf(self.m.temporal_filtering[stage_idx], ref, cur, mv)
        return L, Hh, None, None

    def inverse_MCTF(self, L, Hh, mv, stage_idx=0):
        return inverse_mctf(self.m.temporal_filtering[stage_idx], L, Hh, mv)
