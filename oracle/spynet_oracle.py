"""TEST INFRASTRUCTURE -- fp32 CPU restatement of SpyNet motion estimation (pMCTF/layers/video/video_net.py:74-121: MEBasic,
ME_Spynet; bilinearupsacling :56-61; the warp is oracle.flow_warp, the bit-exact restatement of video_net.py:32-55).  Plain fp32
convolutions (orc_conv2d).  The CUDA path runs the 7x7 layers with bf16 operands, so this is a TOLERANCE oracle for the flow; it is
pinned to the reference's module (tests/golden/spynet.npz, oracle/make_golden.py spynet).  Only tests/ may import this file."""
import numpy as np

from . import oracle as orc
from .ctx_oracle import conv2d


def avg_pool2(x):
    n, c, h, w = x.shape
    v = x.reshape(n, c, h // 2, 2, w // 2, 2)
    return (((v[:, :, :, 0, :, 0] + v[:, :, :, 0, :, 1]) + (v[:, :, :, 1, :, 0] + v[:, :, :, 1, :, 1])) * np.float32(0.25)).astype(np.float32)


def upsample2(x):
    """F.interpolate(scale 2, bilinear, align_corners=False)"""
    n, c, h, w = x.shape

    def axis(size):
        s = np.maximum((np.arange(2 * size, dtype=np.float32) + np.float32(0.5)) * np.float32(0.5) - np.float32(0.5), 0).astype(np.float32)
        i0 = np.floor(s).astype(np.int64)
        i1 = np.minimum(i0 + 1, size - 1)
        return i0, i1, (s - i0).astype(np.float32)
    y0, y1, ay = axis(h)
    x0, x1, ax = axis(w)
    ay, ax = ay[None, None, :, None], ax[None, None, None, :]
    top = x[:, :, y0][:, :, :, x0] * (1 - ax) + x[:, :, y0][:, :, :, x1] * ax
    bot = x[:, :, y1][:, :, :, x0] * (1 - ax) + x[:, :, y1][:, :, :, x1] * ax
    return (top * (1 - ay) + bot * ay).astype(np.float32)


def me_basic(sd, lvl, x):
    for i in range(5):
        x = conv2d(x, sd[f"moduleBasic.{lvl}.conv{i + 1}.weight"], sd[f"moduleBasic.{lvl}.conv{i + 1}.bias"])
        if i < 4:
            x = np.maximum(x, 0)
    return x


def spynet(sd, im1, im2, L=6):
    """im1, im2 [N,3,H,W] -> flow [N,2,H,W]"""
    im1s, im2s = [orc._a(im1)], [orc._a(im2)]
    for _ in range(L - 1):
        im1s.append(avg_pool2(im1s[-1]))
        im2s.append(avg_pool2(im2s[-1]))
    n, _, h, w = im2s[-1].shape
    flow = np.zeros((n, 2, h // 2, w // 2), np.float32)
    for level in range(L):
        up = (upsample2(flow) * np.float32(2.0)).astype(np.float32)
        i = L - 1 - level
        warped = orc.flow_warp(im2s[i], up)
        x = np.concatenate([im1s[i], warped, up], axis=1)
        flow = np.stack([up[b] + me_basic(sd, level, x[b]) for b in range(n)]).astype(np.float32)
    return flow
