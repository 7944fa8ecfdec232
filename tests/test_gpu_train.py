"""Training path (BASELINE.json configs[4]): the differentiable conv / warp kernels and the modules composed from them
against plain torch fp32 autograd (F.conv2d with TF32 off, F.grid_sample) -- floating-point kernels, so the reference is
torch and the bar is a tolerance: relative error <= 2e-4 on outputs and gradients (fp32 accumulation order differs)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import learned_pmctf_b200 as pkg
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return pkg


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def ref_warp(im, flow):
    """flow_warp of the reference (video_net.py:32-55) in plain torch."""
    N, C, H, W = im.shape
    gx = torch.linspace(-1, 1, W, device=im.device).view(1, 1, 1, W).expand(N, 1, H, W)
    gy = torch.linspace(-1, 1, H, device=im.device).view(1, 1, H, 1).expand(N, 1, H, W)
    grid = torch.cat([gx, gy], 1) + torch.cat([flow[:, 0:1] / ((W - 1) / 2), flow[:, 1:2] / ((H - 1) / 2)], 1)
    return F.grid_sample(im, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


@pytest.mark.parametrize("cin,cout", [(1, 16), (16, 16), (16, 1)])
@pytest.mark.parametrize("shape", [(2, 19, 45), (1, 8, 32), (3, 40, 33)])
def test_conv3x3_forward_backward(P, cin, cout, shape):
    from learned_pmctf_b200 import train
    torch.manual_seed(cin * 100 + cout + shape[1])
    N, H, W = shape
    x = torch.randn(N, cin, H, W, device="cuda", requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.3).requires_grad_()
    b = torch.randn(cout, device="cuda", requires_grad=True)
    g = torch.randn(N, cout, H, W, device="cuda")
    y = train._Conv3x3.apply(x, w, b)
    y.backward(g)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    yr = F.conv2d(x, w, b, padding=1)
    yr.backward(g)
    for a, r_, name in zip(got, (yr.detach(), x.grad, w.grad, b.grad), ("y", "dx", "dw", "db")):
        assert rel(a, r_) <= 2e-5, (name, rel(a, r_))


@pytest.mark.parametrize("n,fn", [(1, 1), (2, 2), (4, 2)])
def test_flow_warp_backward(P, n, fn):
    torch.manual_seed(n)
    H, W = 37, 52
    im = torch.rand(n, 1, H, W, device="cuda", requires_grad=True)
    flow = (torch.randn(fn, 2, H, W, device="cuda") * 4).requires_grad_()
    with torch.no_grad():
        flow[:, :, :2] -= 30      # out-of-frame band: clipped coordinates must get zero flow gradient
    g = torch.randn(n, 1, H, W, device="cuda")
    out = P.flow_warp(im, flow)
    out.backward(g)
    got = (out.detach(), im.grad.clone(), flow.grad.clone())
    im.grad = flow.grad = None
    fl = flow.repeat_interleave(n // fn, 0) if fn != n else flow
    ref = ref_warp(im, fl)
    ref.backward(g)
    assert float((got[0] - ref.detach()).abs().max()) <= 2e-4
    assert rel(got[1], im.grad) <= 1e-5
    # bilinear kinks make d/dflow discontinuous at integer coordinates: compare away from them
    assert float(((got[2] - flow.grad).abs() > 1e-3 * flow.grad.abs().max()).float().mean()) <= 2e-3


def torch_pu(pu, x):
    c1 = F.conv2d(x, pu.conv1.weight, pu.conv1.bias, padding=1)
    a = torch.tanh(c1)
    a = torch.tanh(F.conv2d(a, pu.conv2.weight, pu.conv2.bias, padding=1))
    return F.conv2d(c1 + F.conv2d(a, pu.conv3.weight, pu.conv3.bias, padding=1), pu.conv4.weight, pu.conv4.bias, padding=1)


def test_training_step_matches_torch_autograd(P):
    """A miniature of train_pMCTF_L.py's step on the hot path: forward MCTF -> hp transform + STE quantiser -> synthesis ->
    inverse MCTF -> rate proxy + distortion loss -> backward.  Every parameter gradient is compared with the same graph
    built from F.conv2d / grid_sample."""
    torch.manual_seed(0)
    m = P.pMCTF(num_me_stages=2).cuda().train()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 4 and p.shape[-1] == 3:
                p.normal_(0, 0.08)
            elif p.dim() == 1 and p.numel() > 1:
                p.normal_(0, 0.05)
    B, H, W = 2, 64, 64
    ref = torch.rand(B, 1, H, W, device="cuda") * 255
    cur = (ref.roll((1, -2), (2, 3)) + torch.randn_like(ref) * 3).clamp(0, 255)
    mv = (torch.randn(B, 2, H, W, device="cuda") * 2).requires_grad_()

    def run(use_ours):
        for p in m.parameters():
            p.grad = None
        mv.grad = None
        if use_ours:
            L, Hh, pred, _ = m.forward_MCTF(ref, cur, mv, stage_idx=0)
            x_hat, hat = m.hp_coder.spatial_wavelet_dec(Hh, m.hp_coder.QP[-1:], m.hp_coder.QP_ll[-1:], post_process=False, return_symbols=True)
            r, c = m.inverse_MCTF(L, x_hat, mv, stage_idx=0)
        else:
            tl, wt = m.temporal_filtering[0], m.hp_coder.wavelet_transform.lift_h
            flt = lambda net, x, sc: (x + torch_pu(net, x) * 0.1) * sc  # noqa: E731
            pred = flt(tl.P_t, ref_warp(ref, mv), float(tl.scale_p))
            Hh = cur - pred
            L = ref + flt(tl.U_t, ref_warp(Hh, -mv), float(tl.scale_u))

            def term(conv, pu, x):
                xp = torch.cat([x[:, :, 1:2], x, x[:, :, -2:-1]], 2)
                skip = F.conv2d(xp, conv.weight, conv.bias)
                return skip + torch_pu(pu, skip / 256.0) * 256.0 * 0.1

            def fwd1d(x):
                e, o = x[:, :, ::2], x[:, :, 1::2]
                o = o + term(wt.conv_P1, wt.P_1, e); e = e + term(wt.conv_U1, wt.U_1, o)
                o = o + term(wt.conv_P2, wt.P_2, e); e = e + term(wt.conv_U2, wt.U_2, o)
                return e * float(wt.scale_l), o * float(wt.scale_h)

            def bwd1d(l, h):
                l, h = l / float(wt.scale_l), h / float(wt.scale_h)
                l = l - term(wt.conv_U2, wt.U_2, h); h = h - term(wt.conv_P2, wt.P_2, l)
                l = l - term(wt.conv_U1, wt.U_1, h); h = h - term(wt.conv_P1, wt.P_1, l)
                return torch.stack([l, h], 3).reshape(l.size(0), 1, 2 * l.size(2), l.size(3))
            T = lambda t: t.permute(0, 1, 3, 2)  # noqa: E731
            ll, bands = Hh, []
            for _ in range(4):
                l, h = fwd1d(ll)
                a, b_ = fwd1d(T(l)); c_, d = fwd1d(T(h))
                bands.append((T(b_), T(c_), T(d)))
                ll = T(a)
            q, qll = m.hp_coder.QP[-1:], m.hp_coder.QP_ll[-1:]
            ste = lambda v: v + (torch.round(v.clamp(-8192, 8192)) - v).detach()  # noqa: E731
            hat = {3: {"ll": ste(ll * qll)}}
            for lvl in range(4):
                hat.setdefault(lvl, {}).update({k: ste(v * q) for k, v in zip(("lh", "hl", "hh"), bands[lvl])})
            y = hat[3]["ll"] / qll
            for lvl in range(3, -1, -1):
                l = bwd1d(T(y), T(hat[lvl]["lh"] / q)); h = bwd1d(T(hat[lvl]["hl"] / q), T(hat[lvl]["hh"] / q))
                y = bwd1d(T(l), T(h))
            x_hat = y
            inv = flt(tl.U_t, ref_warp(x_hat, -mv), float(tl.scale_u))
            r = L - inv
            c = x_hat + flt(tl.P_t, ref_warp(r, mv), float(tl.scale_p))
        rate = sum(v.abs().mean() for lvl in hat for v in hat[lvl].values())
        loss = ((r - ref) ** 2).mean() + ((c - cur) ** 2).mean() + 0.05 * rate + 0.01 * ((pred - cur) ** 2).mean()
        loss.backward()
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        return loss.detach(), grads, mv.grad.clone()

    l1, g1, gm1 = run(True)
    l2, g2, gm2 = run(False)
    assert abs(float(l1 - l2)) <= 1e-4 * abs(float(l2))
    assert set(g1) == set(g2) and len(g1) >= 16 + 40 + 2
    # forward and inverse lifting use the same nets with opposite signs, so some gradients are small differences of large
    # terms: compare with an absolute floor tied to the largest gradient of the model
    gscale = max(float(v.abs().max()) for v in g2.values())
    report = sorted(((float((g1[k] - g2[k]).abs().max()), float(g2[k].abs().max()), k) for k in g1), reverse=True)[:5]
    for k in g1:
        err, mag = float((g1[k] - g2[k]).abs().max()), float(g2[k].abs().max())
        assert err <= 5e-3 * mag + 2e-5 * gscale, (k, err, mag, gscale, report)
    assert rel(gm1, gm2) <= 5e-2
    # eval mode under no_grad still takes the fused tensor-core kernels and agrees with the training forward
    m.eval()
    with torch.no_grad():
        Lf, Hf, _, _ = m.forward_MCTF(ref, cur, mv.detach(), stage_idx=0)
    m.train()
    Lt, Ht, _, _ = m.forward_MCTF(ref, cur, mv, stage_idx=0)
    assert float((Lf - Lt).abs().max()) <= 1e-3 and float((Hf - Ht).abs().max()) <= 1e-3


def test_config4_shape_training_step_runs(P):
    """configs[4] shape: batch 8 of 256x256 luma clips, GOP-8 (3 levels): one optimiser step through the batched GOP
    analysis/coding/synthesis decreases nothing in particular but must produce finite gradients for every hot-path parameter."""
    from learned_pmctf_b200 import gop as Gm
    torch.manual_seed(1)
    m = P.pMCTF(num_me_stages=3).cuda().train()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 4 and p.shape[-1] == 3:
                p.normal_(0, 0.05)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    B, G, H, W = 8, 8, 256, 256
    clips = torch.rand(B, G, 1, H, W, device="cuda") * 255
    loss_total = 0.0
    for b in range(2):  # two clips are enough for the smoke; the bench line runs all eight
        frames = [clips[b, f:f + 1] for f in range(G)]
        level, coded = frames, {}
        for s in range(Gm.num_stages(G)):
            nxt = []
            for g in range(len(level) // 2):
                mv = torch.zeros(1, 2, H, W, device="cuda")
                L, Hh, _, _ = m.forward_MCTF(level[2 * g], level[2 * g + 1], mv, stage_idx=s)
                q, qll = m.hp_coder.q_pair(8, m.hp_qp_scale(s, 8))
                coded[(s, g)] = m.hp_coder.spatial_wavelet_dec(Hh, q, qll, post_process=False)
                nxt.append(L)
            level = nxt
        q, qll = m.lp_coder.q_pair(8)
        rec = [m.lp_coder.spatial_wavelet_dec(level[0], q, qll, post_process=False)]
        for s in range(Gm.num_stages(G) - 1, -1, -1):
            out = []
            for g, Lr in enumerate(rec):
                r, c = m.inverse_MCTF(Lr, coded[(s, g)], torch.zeros(1, 2, H, W, device="cuda"), stage_idx=s)
                out += [r, c]
            rec = out
        loss = sum(((a - o) ** 2).mean() for a, o in zip(rec, frames)) / G
        loss.backward()
        loss_total += float(loss)
    bad = [k for k, p in m.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not bad, bad[:5]
    assert np.isfinite(loss_total)
    opt.step()


def test_batched_training_loss_backward(P):
    """gop.training_loss_hot_path (pairs batched along N) gives finite gradients for every parameter it uses and the same loss as
    evaluating the clips one by one."""
    from learned_pmctf_b200 import gop as Gm
    torch.manual_seed(3)
    m = P.pMCTF(num_me_stages=2).cuda().train()
    clips = torch.rand(2, 4, 1, 64, 64, device="cuda") * 255
    mvs = [torch.randn(2 * 2, 2, 64, 64, device="cuda"), torch.randn(2 * 1, 2, 64, 64, device="cuda")]
    loss, dist = Gm.training_loss_hot_path(m, clips, mvs, q_index=8)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    one = [Gm.training_loss_hot_path(m, clips[b:b + 1], [mvs[0][b::2], mvs[1][b::2]], q_index=8)[1] for b in range(2)]
    assert abs(float(dist) - float(sum(one) / 2)) <= 1e-4 * float(dist)


@pytest.mark.parametrize("shape", [(2, 19, 45), (1, 64, 64), (3, 8, 130)])
def test_fused_predict_update_node(P, shape):
    """The single autograd node of PredictUpdate (fused tanh / residual / tanh' epilogues) against the same formula on F.conv2d:
    output, input gradient and all eight parameter gradients."""
    from learned_pmctf_b200 import train
    torch.manual_seed(shape[1] + shape[2])
    N, H, W = shape
    pu = P.PredictUpdate(1).cuda()
    with torch.no_grad():
        for p in pu.parameters():
            p.normal_(0, 0.2 if p.dim() == 4 else 0.05)
    x = torch.randn(N, 1, H, W, device="cuda", requires_grad=True)
    g = torch.randn(N, 1, H, W, device="cuda")
    y = train.predict_update(pu, x)
    assert type(y.grad_fn).__name__.startswith("_PredictUpdate")
    y.backward(g)
    got = [y.detach(), x.grad.clone()] + [p.grad.clone() for p in pu.parameters()]
    x.grad = None
    for p in pu.parameters():
        p.grad = None
    c1 = F.conv2d(x, pu.conv1.weight, pu.conv1.bias, padding=1)
    a = torch.tanh(F.conv2d(torch.tanh(c1), pu.conv2.weight, pu.conv2.bias, padding=1))
    yr = F.conv2d(c1 + F.conv2d(a, pu.conv3.weight, pu.conv3.bias, padding=1), pu.conv4.weight, pu.conv4.bias, padding=1)
    yr.backward(g)
    want = [yr.detach(), x.grad] + [p.grad for p in pu.parameters()]
    tf32 = torch.backends.cudnn.allow_tf32
    for a_, b_, name in zip(got, want, ["y", "dx"] + [n for n, _ in pu.named_parameters()]):
        assert rel(a_, b_) <= (5e-3 if tf32 else 5e-5), (name, rel(a_, b_))
