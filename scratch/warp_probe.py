import sys, numpy as np, torch
sys.path[:0] = ["oracle/ref_stubs", "/root/reference"]
from pMCTF.layers.video.video_net import flow_warp, bilineardownsacling
f32 = np.float32
torch.manual_seed(3)
def linspace(n):
    # torch.linspace(-1,1,n) fp32 CPU
    return torch.linspace(-1.0, 1.0, n, dtype=torch.float32).numpy()
def my_linspace(n):
    start, end = f32(-1.0), f32(1.0)
    step = (end - start) / f32(n - 1)
    out = np.empty(n, f32)
    half = n // 2
    for i in range(n):
        if i < half: out[i] = start + step * f32(i)
        else: out[i] = end - step * f32(n - i - 1)
    return out
for n in (7, 96, 1920, 1152, 960, 576):
    print("linspace", n, np.array_equal(linspace(n), my_linspace(n)))

def warp_np(im, flow, variant):
    N, C, H, W = im.shape
    lx, ly = linspace(W), linspace(H)
    sx = f32((W - 1.0) / 2.0); sy = f32((H - 1.0) / 2.0)
    fx = flow[:, 0] / sx; fy = flow[:, 1] / sy
    gx = lx[None, None, :] + fx; gy = ly[None, :, None] + fy
    if variant["unnorm"] == "cpu":   # (g+1) * ((size-1)/2)
        ix = (gx + f32(1)) * f32((W - 1) / 2.0); iy = (gy + f32(1)) * f32((H - 1) / 2.0)
    else:
        ix = ((gx + f32(1)) / f32(2)) * f32(W - 1); iy = ((gy + f32(1)) / f32(2)) * f32(H - 1)
    ix = np.minimum(np.maximum(ix, f32(0)), f32(W - 1)); iy = np.minimum(np.maximum(iy, f32(0)), f32(H - 1))
    x0 = np.floor(ix); y0 = np.floor(iy)
    w = ix - x0; e = f32(1) - w; n_ = iy - y0; s = f32(1) - n_
    nw = s * e; ne = s * w; sw = n_ * e; se = n_ * w
    x0i = x0.astype(np.int64); y0i = y0.astype(np.int64)
    x1i = np.minimum(x0i + 1, W - 1); y1i = np.minimum(y0i + 1, H - 1)
    x1ok = (x0i + 1 <= W - 1); y1ok = (y0i + 1 <= H - 1)
    out = np.empty_like(im)
    bi = np.arange(N)[:, None, None]
    for c in range(C):
        p = im[:, c]
        vnw = p[bi, y0i, x0i]; vne = np.where(x1ok, p[bi, y0i, x1i], f32(0))
        vsw = np.where(y1ok, p[bi, y1i, x0i], f32(0)); vse = np.where(x1ok & y1ok, p[bi, y1i, x1i], f32(0))
        if variant["sum"] == "plain":
            out[:, c] = ((vnw * nw + vne * ne) + vsw * sw) + vse * se
        elif variant["sum"] == "fma":
            # fma chain emulate in float64 (products of fp32 exact in f64; sum rounding approximates fma)
            acc = (vnw.astype(np.float64) * nw).astype(f32)
            acc = (vne.astype(np.float64) * ne + acc).astype(f32)
            acc = (vsw.astype(np.float64) * sw + acc).astype(f32)
            acc = (vse.astype(np.float64) * se + acc).astype(f32)
            out[:, c] = acc
    return out
for (N, C, H, W) in [(1, 1, 64, 96), (2, 1, 36, 60), (1, 3, 40, 72), (1, 1, 1152, 1920)]:
    im = (torch.rand(N, C, H, W) * 255).round()
    flow = torch.randn(N, 2, H, W) * 6
    flow[:, :, :3] *= 20  # some far out of frame
    ref = flow_warp(im, flow).numpy()
    for un in ("cpu", "cuda"):
        for sm in ("plain", "fma"):
            o = warp_np(im.numpy(), flow.numpy(), {"unnorm": un, "sum": sm})
            d = np.abs(o - ref)
            print((N, C, H, W), un, sm, "bitexact" if np.array_equal(o, ref) else f"max {d.max():.3e} nz {(d>0).mean():.4f}")
# chroma mv
mv = torch.randn(1, 2, 64, 96) * 5
ref = (bilineardownsacling(mv) / 2).numpy()
a = mv.numpy()
h1 = (f32(0.5) * a[:, :, :, 0::2] + f32(0.5) * a[:, :, :, 1::2])
v1 = (f32(0.5) * h1[:, :, 0::2] + f32(0.5) * h1[:, :, 1::2]) / f32(2)
v0 = (f32(0.5) * a[:, :, 0::2] + f32(0.5) * a[:, :, 1::2])
h0 = (f32(0.5) * v0[:, :, :, 0::2] + f32(0.5) * v0[:, :, :, 1::2]) / f32(2)
print("chroma h-first", np.array_equal(v1, ref), "v-first", np.array_equal(h0, ref), np.abs(v1-ref).max(), np.abs(h0-ref).max())
