"""Flow-based backward warp and chroma motion-vector scaling
(reference: pMCTF/layers/video/video_net.py:32-71)."""
from ... import ops


def torch_warp(feature, flow):
    """grid_sample(bilinear, border, align_corners=True) driven by a pixel-unit flow (video_net.py:32-50)."""
    return flow_warp(feature, flow)


def flow_warp(im, flow):
    """video_net.py:53-55"""
    from ... import train
    if train.needs_grad(im, flow):  # autograd: forward kernel + its adjoint (pmctf_flow_warp_bwd)
        if flow.size(0) != im.size(0) and im.size(0) % flow.size(0) == 0 and flow.size(0) != 1:
            flow = flow.repeat_interleave(im.size(0) // flow.size(0), 0)
        return train.flow_warp(im, flow)
    return ops.flow_warp(im, flow)


def bilineardownsacling(inputfeature, factor=2):
    """2x2 box mean (video_net.py:66-71).  The reference's only use divides the result by 2
    (pMCTF_L.py:317,336,401); the kernel fuses that, so it is undone here for the bare call."""
    if factor != 2 or inputfeature.size(1) != 2:
        raise NotImplementedError("only the factor-2 motion-field case of the hot path is implemented")
    from ... import train
    if train.needs_grad(inputfeature):
        return train.chroma_mv(inputfeature) * 2.0
    return ops.chroma_mv_down(inputfeature) * 2.0


def chroma_mv(mv):
    """bilineardownsacling(mv) / 2 as one kernel."""
    return ops.chroma_mv_down(mv)


# ---- SpyNet motion estimation (video_net.py:74-121) ---------------------------------------------
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from ... import _native as nat  # noqa: E402


def bilinearupsacling(inputfeature, factor=2):
    """video_net.py:56-61"""
    h, w = inputfeature.size(2), inputfeature.size(3)
    return F.interpolate(inputfeature, (h * factor, w * factor), mode="bilinear", align_corners=False)


class MEBasic(nn.Module):
    """One pyramid level's network: five 7x7 convolutions 8 -> 32 -> 64 -> 32 -> 16 -> 2, ReLU between them (video_net.py:74-91).
    On CUDA tensors outside autograd the layers run as tcgen05 CTA-pair implicit GEMMs (csrc/pmctf_pairconv.cu) between
    channel-chunked bf16 maps; the module is driven by ME_Spynet, which owns the layouts."""
    CHANNELS = ((8, 32), (32, 64), (64, 32), (32, 16), (16, 2))

    def __init__(self, in_ch=8):
        super().__init__()
        if in_ch != 8:
            raise NotImplementedError("SpyNet's level network takes [im1 (3), warp(im2) (3), flow (2)] (video_net.py:115-119)")
        self.relu = nn.ReLU()
        for i, (ci, co) in enumerate(self.CHANNELS):
            setattr(self, f"conv{i + 1}", nn.Conv2d(ci, co, 7, 1, padding=3))
        self._key = None
        self._packed = None

    def convs(self):
        return [getattr(self, f"conv{i + 1}") for i in range(5)]

    def forward(self, x, modes=None):
        for i, c in enumerate(self.convs()):
            x = c(x)
            if i < 4:
                x = self.relu(x)
        return x

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_key"], st["_packed"] = None, None
        return st

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._key = None

    @staticmethod
    def _pad16(c):
        return (c + 15) // 16 * 16

    def packed(self):
        """bf16 operand images of the five layers (one buffer), rebuilt when a weight changes"""
        key = tuple((c.weight.data_ptr(), c.weight._version) for c in self.convs())
        if key != self._key:
            lib = nat.lib()
            dev = self.conv1.weight.device
            sizes = [int(lib.pmctf_pair_packed_bytes(7, self._pad16(ci), self._pad16(co))) for ci, co in self.CHANNELS]
            buf = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
            offs, off = [], 0
            for (ci, co), c, sz in zip(self.CHANNELS, self.convs(), sizes):
                w = ops._chk(c.weight.detach().contiguous(), "conv weight")
                ops._launch(dev, "pair_pack_conv", lib.pmctf_pair_pack_conv, w.data_ptr(), co, ci, 7, self._pad16(ci), self._pad16(co),
                            buf.data_ptr() + off)
                offs.append(off)
                off += sz
            self._key, self._packed = key, (buf, offs)
        return self._packed

    def run(self, rec, flow_up, N, H, W):
        """rec: bf16 [N,2,H,W,8] level input records; returns flow_up + conv5(...) as fp32 [N,2,H,W]"""
        buf, offs = self.packed()
        lib, dev = nat.lib(), rec.device
        x = rec
        for i, ((ci, co), c) in enumerate(zip(self.CHANNELS, self.convs())):
            cip, cop = self._pad16(ci), self._pad16(co)
            last = i == 4
            out_b = None if last else torch.empty((N, cop // 8, H, W, 8), dtype=torch.bfloat16, device=dev)
            out_f = torch.empty((N, 2, H, W), dtype=torch.float32, device=dev) if last else None
            ops._launch(dev, "pair_conv", lib.pmctf_pair_conv, x.data_ptr(), buf.data_ptr() + offs[i], c.bias.detach().data_ptr(), 7, cip, co, cop,
                        1.0 if last else 0.0, out_b.data_ptr() if out_b is not None else None, out_f.data_ptr() if last else None,
                        flow_up.data_ptr() if last else None, N, H, W)
            x = out_f if last else out_b
        return x


class ME_Spynet(nn.Module):
    """Coarse-to-fine optical flow over L = 6 pyramid levels (video_net.py:94-121): flow = up(flow) * 2 + MEBasic_l([im1_l,
    warp(im2_l, up(flow) * 2), up(flow) * 2]).  Same module tree and parameter names (moduleBasic.{l}.conv{1..5})."""

    def __init__(self, in_ch=8, L=6):
        super().__init__()
        self.L = L
        self.moduleBasic = nn.ModuleList([MEBasic(in_ch=in_ch) for _ in range(L)])

    def _forward_torch(self, im1, im2, modes=None):
        im1s, im2s = [im1], [im2]
        for _ in range(self.L - 1):
            im1s.append(F.avg_pool2d(im1s[-1], kernel_size=2, stride=2))
            im2s.append(F.avg_pool2d(im2s[-1], kernel_size=2, stride=2))
        n, _, h, w = im2s[-1].shape
        flow = torch.zeros((n, 2, h // 2, w // 2), dtype=im1.dtype, device=im1.device)
        for level in range(self.L):
            up = bilinearupsacling(flow) * 2.0
            i = self.L - 1 - level
            flow = up + self.moduleBasic[level](torch.cat([im1s[i], flow_warp(im2s[i], up), up], 1), modes)
        return flow

    def forward(self, im1, im2, modes=None):
        from ... import train
        if train.needs_grad(im1, im2, self) or not im1.is_cuda:
            if not im1.is_cuda:
                raise RuntimeError("ME_Spynet: CUDA tensors only (there is no CPU path)")
            return self._forward_torch(im1, im2, modes)
        im1, im2 = ops._chk(im1, "im1", 4).contiguous(), ops._chk(im2, "im2", 4).contiguous()
        N, Cc, H, W = im1.shape
        if Cc != 3 or tuple(im2.shape) != (N, 3, H, W) or H % (1 << self.L) or W % (1 << self.L):
            raise RuntimeError(f"ME_Spynet: two [N,3,H,W] images with H, W multiples of {1 << self.L} (got {tuple(im1.shape)}, {tuple(im2.shape)})")
        lib, dev = nat.lib(), im1.device
        im1s, im2s = [im1], [im2]
        for lvl in range(self.L - 1):
            h, w = H >> lvl, W >> lvl
            for src in (im1s, im2s):
                dst = torch.empty((N, 3, h // 2, w // 2), dtype=torch.float32, device=dev)
                ops._launch(dev, "avgpool2", lib.pmctf_avgpool2, src[-1].data_ptr(), dst.data_ptr(), N * 3, h, w)
                src.append(dst)
        flow = None
        for level in range(self.L):
            i = self.L - 1 - level
            h, w = H >> i, W >> i
            flow_up = torch.empty((N, 2, h, w), dtype=torch.float32, device=dev)
            rec = torch.empty((N, 2, h, w, 8), dtype=torch.bfloat16, device=dev)
            ops._launch(dev, "spynet_prep", lib.pmctf_spynet_prep, im1s[i].data_ptr(), im2s[i].data_ptr(), flow.data_ptr() if flow is not None else None,
                        flow_up.data_ptr(), rec.data_ptr(), N, h, w)
            flow = self.moduleBasic[level].run(rec, flow_up, N, h, w)
        return flow
