"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, CPU fp32).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference is imported through four stub packages under oracle/ref_stubs/ (timm, matplotlib,
mpl_toolkits, pytorch_msssim are absent from this image and irrelevant to the hot path;
SURVEY.md section 8c).  Nothing in tests/ or the product imports the reference at run time: the vectors
written here are the only thing that travels.

Model: pMCTF(num_me_stages=4), torch.manual_seed(0), then the following PERTURBATIONS so that the
random-init degeneracies (SURVEY.md section 7 "Random-init degeneracy") do not hide bugs:
  * every hot-path 3x3 conv weight (temporal_filtering.*, *.wavelet_transform.*) is multiplied by 4
    and every hot-path conv bias is drawn from N(0, 0.05)          [seed 1234]
  * the (3,1) skip taps are reset to their bior4.4 init values (pMCTF._init_weights had
    overwritten them, pMCTF_L.py:116-122) plus N(0, 0.01), biases N(0, 0.05)
  * QP = [1/32, 1/2], QP_ll = [1/16, 3/4] for both coders; hp_q_scale[i] = [0.9-0.1 i, 1.3-0.1 i]
Inputs: seeded synthetic frames (smooth noise in 16..235 plus texture), flows N(0, 4^2) px with
an out-of-frame band.  Everything is recorded in the files themselves.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference"]
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from pMCTF.layers import RoundNoGradient  # noqa: E402
from pMCTF.layers.video.video_net import bilineardownsacling, flow_warp  # noqa: E402
from pMCTF.models.video.pMCTF_L import pMCTF  # noqa: E402


def is_hot(k):
    return k.startswith("temporal_filtering.") or ".wavelet_transform.lift_h." in k or k.startswith("hp_q_scale") \
        or k.endswith(".QP") or k.endswith(".QP_ll")


def build_model(lossy=True):
    torch.manual_seed(0)
    m = pMCTF(lossy=lossy, num_me_stages=4).eval()
    g = torch.Generator().manual_seed(1234)
    bior = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971]
    with torch.no_grad():
        for k, p in m.named_parameters():
            if not (k.startswith("temporal_filtering.") or ".wavelet_transform.lift_h." in k):
                continue
            leaf = k.split(".")[-2]
            if leaf.startswith("conv_"):  # skip taps (1,1,3,1)
                i = ["conv_P1", "conv_U1", "conv_P2", "conv_U2"].index(leaf)
                if k.endswith("weight"):
                    base = torch.tensor([0.0, bior[i], bior[i]] if i % 2 == 0 else [bior[i], bior[i], 0.0])
                    p.copy_((base + 0.01 * torch.randn(3, generator=g)).view(1, 1, 3, 1))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif k.endswith("weight"):
                p.mul_(4.0)
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
        for coder in (m.lp_coder, m.hp_coder):
            coder.QP.copy_(torch.tensor([1 / 32, 1 / 2]).view(2, 1, 1, 1))
            coder.QP_ll.copy_(torch.tensor([1 / 16, 3 / 4]).view(2, 1, 1, 1))
        for i, p in enumerate(m.hp_q_scale):
            p.copy_(torch.tensor([0.9 - 0.1 * i, 1.3 - 0.1 * i]).view(2, 1, 1, 1))
    return m


def frames(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 1, h + 16, w + 16, generator=g)
    k = torch.ones(1, 1, 5, 5) / 25
    for _ in range(2):
        x = F.conv2d(x, k, padding=2)
    x = x[:, :, 8:-8, 8:-8]
    x = (x - x.amin()) / (x.amax() - x.amin())
    x = 16 + 219 * x + 6 * torch.randn(n, 1, h, w, generator=g)
    return x.clamp(0, 255).round()


def flows(n, h, w, seed, sigma=4.0):
    g = torch.Generator().manual_seed(seed)
    f = sigma * torch.randn(n, 2, h, w, generator=g)
    f = F.avg_pool2d(F.pad(f, (2, 2, 2, 2), mode="replicate"), 5, stride=1)
    f = f * 2.5
    f[:, :, :2, :] -= 40.0  # points out of the frame at the top
    f[:, :, :, -2:] += 37.5  # and at the right border
    return f


def npy(t):
    return t.detach().cpu().numpy().astype(np.float32)


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    m = build_model()
    sd = {k: npy(v) for k, v in m.state_dict().items() if is_hot(k)}
    np.savez_compressed(os.path.join(OUT, "weights.npz"), **sd)
    meta = f"torch {torch.__version__}; cpu capability {torch.backends.cpu.get_cpu_capability()}; threads 8"

    # ---- a1/a2: warp + chroma MV ------------------------------------------------------------
    c = {}
    for tag, (n, ch, h, w, fn) in {"luma": (1, 1, 40, 72, 1), "chromaN": (2, 1, 20, 36, 2), "tile": (2, 1, 24, 40, 1),
                                   "rgb": (1, 3, 16, 24, 1)}.items():
        g = torch.Generator().manual_seed({"luma": 11, "chromaN": 12, "tile": 13, "rgb": 14}[tag])
        im = torch.rand(n, ch, h, w, generator=g) * 255
        fl = flows(fn, h, w, 20 + n + h)
        fl_t = fl.tile((n // fn, 1, 1, 1)) if fn != n else fl
        c[f"{tag}.im"], c[f"{tag}.flow"] = npy(im), npy(fl)
        c[f"{tag}.out_pos"] = npy(flow_warp(im, fl_t))
        c[f"{tag}.out_neg"] = npy(flow_warp(im, -fl_t))
        c[f"{tag}.lin_x"] = npy(torch.linspace(-1.0, 1.0, w))
        c[f"{tag}.lin_y"] = npy(torch.linspace(-1.0, 1.0, h))
    mv = flows(1, 32, 48, 31)
    c["down.mv"], c["down.out"] = npy(mv), npy(bilineardownsacling(mv) / 2)
    np.savez_compressed(os.path.join(OUT, "warp.npz"), meta=meta, **c)

    # ---- a3/a4: PredictUpdate, temporal filters -----------------------------------------------
    c = {}
    x = frames(2, 24, 40, 41)
    tf0 = m.temporal_filtering[0]
    c["x"] = npy(x)
    c["P_t0"] = npy(tf0.P_t(x))
    c["P_1_norm"] = npy(m.hp_coder.wavelet_transform.lift_h.P_1(x / 256.0))
    c["predict0"] = npy(tf0.predict_filter(x))
    c["update3"] = npy(m.temporal_filtering[3].update_filter(x - 100.0))
    np.savez_compressed(os.path.join(OUT, "pu.npz"), meta=meta, **c)

    # ---- a5/a6: forward/inverse MCTF (luma + chroma with tiled, down-scaled MV) ---------------
    c = {}
    H, W = 64, 96
    fr = frames(2, H, W, 51)
    ch = frames(4, H // 2, W // 2, 52)
    mvh = flows(1, H, W, 53)
    c["lin_x"], c["lin_y"] = npy(torch.linspace(-1.0, 1.0, W)), npy(torch.linspace(-1.0, 1.0, H))
    c["lin_xc"], c["lin_yc"] = npy(torch.linspace(-1.0, 1.0, W // 2)), npy(torch.linspace(-1.0, 1.0, H // 2))
    c["ref"], c["cur"], c["ref_c"], c["cur_c"], c["mv"] = npy(fr[0:1]), npy(fr[1:2]), npy(ch[0:2]), npy(ch[2:4]), npy(mvh)
    for s in (0, 3):
        L, Hh, pred, inv = m.forward_MCTF(fr[0:1], fr[1:2], mvh, stage_idx=s)
        c[f"s{s}.L"], c[f"s{s}.H"], c[f"s{s}.pred"], c[f"s{s}.inv"] = npy(L), npy(Hh), npy(pred), npy(inv)
        r, cu = m.inverse_MCTF(L, Hh, mvh, stage_idx=s)
        c[f"s{s}.ref_rec"], c[f"s{s}.cur_rec"] = npy(r), npy(cu)
        mvc = bilineardownsacling(mvh) / 2
        Lc, Hc, _, _ = m.forward_MCTF(ch[0:2], ch[2:4], mvc, stage_idx=s)
        c[f"s{s}.Lc"], c[f"s{s}.Hc"] = npy(Lc), npy(Hc)
        rc, cc = m.inverse_MCTF(Lc, Hc, mvh, downscale=True, stage_idx=s)
        c[f"s{s}.ref_c_rec"], c[f"s{s}.cur_c_rec"] = npy(rc), npy(cc)
    np.savez_compressed(os.path.join(OUT, "mctf.npz"), meta=meta, **c)

    # ---- a7-a12: spatial lifting, quantise ----------------------------------------------------
    c = {}
    coder = m.hp_coder
    lift = coder.wavelet_transform
    x = frames(2, 32, 48, 61) - 60.0
    l, h = lift.lift_h.forward_lift(x)
    c["x1d"], c["l1d"], c["h1d"] = npy(x), npy(l), npy(h)
    c["x1d_rec"] = npy(lift.lift_h.backward_lift(l, h))
    d = lift.forward_lift_2d(x)
    for k in ("ll", "lh", "hl", "hh"):
        c[f"2d.{k}"] = npy(d[k].contiguous())
    c["x2d_rec"] = npy(lift.backward_lift_2d(d))
    x = frames(1, 64, 96, 62)
    c["x"] = npy(x)
    y = coder.encode(x)
    for lvl in range(4):
        for k in ("ll", "lh", "hl", "hh"):
            c[f"enc.{lvl}.{k}"] = npy(y[lvl][k].contiguous())
    c["dec"] = npy(coder.decode({lvl: dict(y[lvl]) for lvl in range(4)}))
    # spatial_wavelet_dec (pWave.py:314-349) up to, not including, the PostProcess net
    for qi in (0, 4, 8, 12, 16, 20):
        for tag, scale in (("lp", None), ("hp1", m.get_curr_q(m.hp_q_scale[1], qi))):
            q = coder.get_curr_q(coder.QP, qi)
            qll = coder.get_curr_q(coder.QP_ll, qi)
            if scale is not None:
                q, qll = q * scale, qll * scale
            yy = coder.encode(x)
            hat = {lvl: {} for lvl in range(4)}
            hat[3]["ll"] = RoundNoGradient.apply(coder.quantize_subband(yy[3]["ll"], qll))
            for lvl in range(3, -1, -1):
                for b in ("lh", "hl", "hh"):
                    hat[lvl][b] = RoundNoGradient.apply(coder.quantize_subband(yy[lvl][b], q))
            rec = coder.dequantize_subbands(hat, q, qll)
            xh = coder.decode(rec)
            p = f"q{qi}.{tag}."
            c[p + "q"], c[p + "qll"], c[p + "x_hat"] = npy(q), npy(qll), npy(xh)
            for lvl in hat:
                for b, v in hat[lvl].items():
                    c[p + f"sym.{lvl}.{b}"] = npy(v.contiguous()).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "pwave.npz"), meta=meta, **c)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


@torch.no_grad()
def main_lossless():
    """tests/golden/lossless.npz: the lossless variant of the path (pMCTF(lossy=False): the warp output, 0.1 * PU and every
    lifting update are rounded, no subband scaling -- lifting_1d.py:110-148, wavelet_transform_temporal_mctf.py:30-43,
    pMCTF_L.py:302-326).  Integer frames in, integer subbands out, perfect reconstruction.  The model's own hot-path weights
    are stored in the file (prefix `w.`): the lossless model is a separate construction."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    m = build_model(lossy=False)
    c = {"w." + k: npy(v) for k, v in m.state_dict().items() if is_hot(k)}
    meta = f"torch {torch.__version__}; cpu capability {torch.backends.cpu.get_cpu_capability()}; threads 8"
    H, W = 64, 96
    fr = frames(2, H, W, 71)
    ch = frames(4, H // 2, W // 2, 72)
    mvh = flows(1, H, W, 73)
    c["lin_x"], c["lin_y"] = npy(torch.linspace(-1.0, 1.0, W)), npy(torch.linspace(-1.0, 1.0, H))
    c["lin_xc"], c["lin_yc"] = npy(torch.linspace(-1.0, 1.0, W // 2)), npy(torch.linspace(-1.0, 1.0, H // 2))
    c["ref"], c["cur"], c["ref_c"], c["cur_c"], c["mv"] = npy(fr[0:1]), npy(fr[1:2]), npy(ch[0:2]), npy(ch[2:4]), npy(mvh)
    for s in (0, 2):
        L, Hh, pred, inv = m.forward_MCTF(fr[0:1], fr[1:2], mvh, stage_idx=s)
        c[f"s{s}.L"], c[f"s{s}.H"], c[f"s{s}.pred"], c[f"s{s}.inv"] = npy(L), npy(Hh), npy(pred), npy(inv)
        r, cu = m.inverse_MCTF(L, Hh, mvh, stage_idx=s)
        c[f"s{s}.ref_rec"], c[f"s{s}.cur_rec"] = npy(r), npy(cu)
        mvc = bilineardownsacling(mvh) / 2
        Lc, Hc, _, _ = m.forward_MCTF(ch[0:2], ch[2:4], mvc, stage_idx=s)
        c[f"s{s}.Lc"], c[f"s{s}.Hc"] = npy(Lc), npy(Hc)
    coder = m.lp_coder
    lift = coder.wavelet_transform
    x = frames(2, 32, 48, 74) - 60.0
    l, h = lift.lift_h.forward_lift(x)
    c["x1d"], c["l1d"], c["h1d"] = npy(x), npy(l), npy(h)
    c["x1d_rec"] = npy(lift.lift_h.backward_lift(l, h))
    x = frames(1, 64, 96, 75)
    c["x"] = npy(x)
    y = coder.encode(x)
    for lvl in range(4):
        for k in ("ll", "lh", "hl", "hh"):
            c[f"enc.{lvl}.{k}"] = npy(y[lvl][k].contiguous())
    c["dec"] = npy(coder.decode({lvl: dict(y[lvl]) for lvl in range(4)}))
    np.savez_compressed(os.path.join(OUT, "lossless.npz"), meta=meta, **c)
    print("lossless.npz", os.path.getsize(os.path.join(OUT, "lossless.npz")),
          "perfect reconstruction in the reference:", bool((c["dec"] == c["x"]).all()), bool((c["s0.ref_rec"] == c["ref"]).all()))


def main_postprocess():
    """tests/golden/postprocess.npz: the reference's PostProcess module (pMCTF/layers/postprocessing.py:20-44) as pWave applies it
    (pWave.py:300: dequantModule(x_hat / 256) * 256), CPU fp32.  Weights: N(0, 0.05) for the 64 -> 64 layers (gain ~1.2 per layer,
    so the activations stay O(1)..O(10) through the 13 layers instead of vanishing as with the 0.02 init), N(0, 0.3) for 1 -> 64,
    N(0, 1e-4) for 64 -> 1 (a correction of a few grey levels, like a trained filter), biases N(0, 0.05); seed 4321."""
    from pMCTF.layers.postprocessing import PostProcess
    torch.manual_seed(7)
    pp = PostProcess().eval()
    g = torch.Generator().manual_seed(4321)
    with torch.no_grad():
        for k, p in pp.named_parameters():
            if k.endswith("weight"):
                std = 0.3 if p.shape[1] == 1 else (1e-4 if p.shape[0] == 1 else 0.05)
                p.copy_(std * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    out = {"w." + k: npy(v) for k, v in pp.state_dict().items()}
    x = frames(2, 72, 100, 99)
    with torch.no_grad():
        y = pp(x / 256.0) * 256.0
        # one 64 -> 64 layer on a real feature map, for the single-layer tests
        feat = pp.conv1(x[:1, :, :24, :40] / 256.0)      # a small crop keeps the fixture small
        blk = pp.resBlocks[0]
        mid = blk.lrelu(blk.conv1(feat))
        res = blk.conv2(mid) + feat
    out.update({"x": npy(x), "y": npy(y), "feat": npy(feat), "mid": npy(mid), "res0": npy(res)})
    print("postprocess: |y - x| mean %.3f max %.3f (0..255 scale); feature rms %.3f -> %.3f" %
          (float((y - x).abs().mean()), float((y - x).abs().max()), float(feat.pow(2).mean().sqrt()), float(res.pow(2).mean().sqrt())))
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **out)


def main_ctx4():
    """tests/golden/ctx4.npz: the reference's ContextFusionFourStep (pMCTF/layers/context_fusion_4step.py:23-194) exactly as pWave
    builds it (pWave.py:70-78: num_features 112, num_parameters 2, ctx_channels 2 below the top level, 1 at the top level), CPU
    fp32, with the seeded weights of tests/ctx_weights.py (regenerated by the tests, not stored).  Inputs: a quantiser-scaled
    subband (Laplacian, sigma 3, plus a smooth component), an LSTM-like context plane in [-1, 1] and, for ctx_channels 2, the
    coarser level's reconstructed subband at half the size."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import ctx_weights
    from pMCTF.layers.context_fusion_4step import ContextFusionFourStep
    out = {}
    for tag, cc, seed in (("a", 2, 11), ("b", 1, 12)):
        m = ContextFusionFourStep(in_channels=1, num_features=112, num_parameters=2, lossy=True, ctx_channels=cc).eval()
        w = ctx_weights.make(seed, cc)
        missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
        g = torch.Generator().manual_seed(100 + seed)
        N, H, W = 2, 24, 40
        lap = torch.distributions.laplace.Laplace(0.0, 3.0).sample((N, 1, H, W))
        x = lap + 4.0 * frames(N, H, W, 50 + seed) / 255.0
        context = torch.tanh(torch.randn(N, 1, H, W, generator=g))
        prev = torch.round(2.0 * torch.randn(N, 1, H // 2, W // 2, generator=g)) if cc == 2 else None
        with torch.no_grad():
            x_res, x_q, x_hat, s_hat = m(x, context=context, prev_subband=prev)
            comp = m.compress(x, context=context, prev_subband=prev)
        out.update({f"{tag}.x": npy(x), f"{tag}.context": npy(context), f"{tag}.x_res": npy(x_res), f"{tag}.x_q": npy(x_q),
                    f"{tag}.x_hat": npy(x_hat), f"{tag}.s_hat": npy(s_hat), f"{tag}.seed": np.array(seed), f"{tag}.ctx_channels": np.array(cc)})
        if prev is not None:
            out[f"{tag}.prev"] = npy(prev)
        for i in range(4):
            out[f"{tag}.x_q_{i}"], out[f"{tag}.s_w_{i}"] = npy(comp[i]), npy(comp[4 + i])
        print(f"ctx4 {tag}: scales {float(s_hat.min()):.3f}..{float(s_hat.max()):.3f} (mean {float(s_hat.mean()):.3f}), |x_q| mean "
              f"{float(x_q.abs().mean()):.3f}, |x - x_hat| max {float((x - x_hat).abs().max()):.3f}, missing keys: {missing}")
    np.savez_compressed(os.path.join(OUT, "ctx4.npz"), **out)


def main_spynet():
    """tests/golden/spynet.npz: the reference's ME_Spynet (pMCTF/layers/video/video_net.py:94-121) on a synthetic frame pair shifted by
    (1.5, -0.75) px, as pMCTF feeds it (luma / 255 tiled to 3 channels, pMCTF_L.py:253-256), CPU fp32, seeded weights of
    tests/spynet_weights.py."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import spynet_weights
    from pMCTF.layers.video.video_net import ME_Spynet
    m = ME_Spynet(L=6).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in spynet_weights.make(21).items()}, strict=True)
    f = frames(2, 128, 192, 77)
    cur = (f[0:1] / 255.0).tile((1, 3, 1, 1))
    ref = (f[1:2] / 255.0).tile((1, 3, 1, 1))
    with torch.no_grad():
        flow = m(cur, ref)
    print("spynet: flow range %.3f .. %.3f, mean |flow| %.3f" % (float(flow.min()), float(flow.max()), float(flow.abs().mean())))
    np.savez_compressed(os.path.join(OUT, "spynet.npz"), cur=npy(cur[:, :1]), ref=npy(ref[:, :1]), flow=npy(flow), seed=np.array(21))


def main_llar():
    """tests/golden/llar.npz: the reference's LL-band autoregressive model ContextFusionSubband (pMCTF/layers/context_fusion.py:56-204)
    driven coefficient by coefficient exactly as pWave._compress_subband_ar does (pMCTF/models/pWave.py:531-553: forward_sequential on
    the padded band, symbol = round(y - mean), reconstruction round(symbol + mean); the reconstruction is NOT fed back on the encoder
    side -- the padded input is the quantised band itself), CPU fp32, seeded weights of llar_weights()."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from llar_weights import llar_weights
    from pMCTF.layers.context_fusion import ContextFusionSubband
    out = {}
    for tag, seed, (H, W) in (("a", 31, (9, 14)), ("b", 32, (12, 7))):
        m = ContextFusionSubband(num_features=128, num_parameters=2, context=False, in_channels=1).eval()
        sd = llar_weights(seed)
        full = m.state_dict()
        for k, v in sd.items():
            full[k] = torch.from_numpy(v)
        m.load_state_dict(full, strict=True)      # the mask buffers keep their constructed values
        g = torch.Generator().manual_seed(200 + seed)
        y = torch.round(torch.randn(1, 1, H, W, generator=g) * 6)
        pad = F.pad(y, (1, 1, 1, 1))
        scales, means, sym, rec = (np.zeros((H, W), np.float32) for _ in range(4))
        with torch.no_grad():
            for h in range(H):
                for w in range(W):
                    p = m.forward_sequential(pad, h, w)
                    sc, mu = p.chunk(2, dim=1)
                    res = (pad[:, :, h + 1:h + 2, w + 1:w + 2] - mu).round()
                    scales[h, w], means[h, w], sym[h, w], rec[h, w] = float(sc), float(mu), float(res), float((res + mu).round())
            full_plane = m(y)
        out.update({f"{tag}.y": npy(y[0, 0]), f"{tag}.scales": scales, f"{tag}.means": means, f"{tag}.symbols": sym, f"{tag}.recon": rec,
                    f"{tag}.full_plane": npy(full_plane[0]), f"{tag}.seed": np.array(seed)})
        print(f"llar {tag}: mean range {means.min():.3f}..{means.max():.3f}, scale {scales.min():.3f}..{scales.max():.3f}, "
              f"|sequential - full plane| max {np.abs(np.stack([scales, means]) - npy(full_plane[0])).max():.2e}, recon == y: {np.array_equal(rec, npy(y[0, 0]))}")
    np.savez_compressed(os.path.join(OUT, "llar.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "llar":
        main_llar()
    elif len(sys.argv) > 1 and sys.argv[1] == "spynet":
        main_spynet()
    elif len(sys.argv) > 1 and sys.argv[1] == "ctx4":
        main_ctx4()
    elif len(sys.argv) > 1 and sys.argv[1] == "postprocess":
        main_postprocess()
    elif len(sys.argv) > 1 and sys.argv[1] == "lossless":
        main_lossless()
    else:
        main()
