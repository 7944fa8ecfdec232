"""SpyNet motion estimation on the GPU: the run-time-shaped CTA-pair convolution against a
bf16-operand emulation for every layer shape of the network, and the whole six-level estimator against the fp32 oracle pinned to
the reference's module (flow tolerance: the estimator runs in the encoder only, so bf16 operands cost accuracy of the vectors,
never encoder / decoder agreement)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import spynet_weights  # noqa: E402

import learned_pmctf_b200 as pkg  # noqa: E402
from learned_pmctf_b200 import _native as nat  # noqa: E402
from learned_pmctf_b200.layers.video.video_net import ME_Spynet  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_planar(x, cpad):   # [N,C,H,W] fp32 -> bf16 [N,cpad/8,H,W,8] with zero padding channels
    N, Cc, H, W = x.shape
    xp = torch.zeros(N, cpad, H, W)
    xp[:, :Cc] = x
    return xp.view(N, cpad // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def from_planar(t):
    N, G, H, W, _ = t.shape
    return t.float().permute(0, 1, 4, 2, 3).reshape(N, G * 8, H, W)


@pytest.mark.parametrize("ks,cin,cout,N,H,W,slope,nchw", [(7, 8, 32, 1, 36, 60, 0.0, False), (7, 32, 64, 2, 19, 45, 0.0, False),
                                                           (7, 64, 32, 1, 72, 120, 0.0, False), (7, 32, 16, 1, 9, 27, 0.0, False),
                                                           (7, 16, 2, 2, 18, 30, 1.0, True), (3, 48, 48, 1, 20, 33, 0.2, False),
                                                           (1, 16, 128, 1, 7, 70, 1.0, False), (7, 64, 32, 1, 144, 240, 0.0, False)])
def test_pair_conv_vs_bf16_emulation(ks, cin, cout, N, H, W, slope, nchw, conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    g = torch.Generator().manual_seed(ks * 100 + cin + cout + H)
    cip, cop = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) / (ks * np.sqrt(cin))
    b = torch.randn(cout, generator=g) * 0.1
    add = torch.randn(N, cout, H, W, generator=g) if nchw else None
    want = F.conv2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), b.double(), padding=ks // 2)
    if add is not None:
        want = want + add.double()
    want = torch.where(want >= 0, want, want * slope).float()
    lib = nat.lib()
    st = torch.cuda.current_stream().cuda_stream
    packed = torch.empty(int(lib.pmctf_pair_packed_bytes(ks, cip, cop)), dtype=torch.uint8, device=DEV)
    wd, bd, xd = w.to(DEV).contiguous(), b.to(DEV), to_planar(x, cip).to(DEV)
    nat.check(lib.pmctf_pair_pack_conv(wd.data_ptr(), cout, cin, ks, cip, cop, packed.data_ptr(), st), "pack")
    ob = torch.full((N, cop // 8, H, W, 8), 7.0, dtype=torch.bfloat16, device=DEV)
    on = torch.full((N, cout, H, W), float("nan"), device=DEV) if nchw else None
    addd = add.to(DEV) if add is not None else None
    nat.check(lib.pmctf_pair_conv(xd.data_ptr(), packed.data_ptr(), bd.data_ptr(), ks, cip, cout, cop, slope, ob.data_ptr(),
                                  on.data_ptr() if nchw else None, addd.data_ptr() if nchw else None, N, H, W, st), "pair_conv")
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    got = from_planar(ob).cpu()
    assert torch.equal(got[:, cout:], torch.zeros_like(got[:, cout:]))           # padding channels are written as zeros
    scale = max(1.0, want.abs().max().item())
    assert (got[:, :cout] - want.to(torch.bfloat16).float()).abs().max().item() <= 2 ** -7 * scale     # one bf16 rounding of the output
    if nchw:
        assert (on.cpu() - want).abs().max().item() < 2e-4 * scale


def test_spynet_vs_oracle_and_reference(golden, conv_mode, capsys):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    g = golden("spynet")
    sd = spynet_weights.make(int(g["seed"]))
    m = ME_Spynet(L=6).to(DEV).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    cur = torch.from_numpy(np.repeat(g["cur"], 3, axis=1)).to(DEV)
    ref = torch.from_numpy(np.repeat(g["ref"], 3, axis=1)).to(DEV)
    with torch.no_grad():
        flow = m(cur, ref)
        flow_t = m._forward_torch(cur, ref)          # the same module on stock torch ops (fp32 / TF32 cuDNN)
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    d = np.abs(flow.cpu().numpy() - g["flow"])
    dt = np.abs(flow_t.cpu().numpy() - g["flow"])
    with capsys.disabled():
        print(f"\n[spynet] bf16 tensor-core path vs the reference's fp32 flow: mean |d| {d.mean():.4f} px, max {d.max():.4f} px (flow magnitude "
              f"up to {np.abs(g['flow']).max():.2f} px); stock torch on this GPU (TF32 convolutions): mean {dt.mean():.4f}, max {dt.max():.4f}")
    assert d.mean() < 0.03 and d.max() < 0.5
    # the CPU oracle reproduces the reference (the check that pins it runs without a GPU: tests/test_spynet_oracle.py)


def test_spynet_1080p_runs(conv_mode):
    if conv_mode != "tensor":
        pytest.skip("independent of the lifting arithmetic")
    m = ME_Spynet(L=6).to(DEV).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in spynet_weights.make(3).items()})
    g = torch.Generator().manual_seed(1)
    a = F.avg_pool2d(torch.rand(1, 1, 1156, 1924, generator=g), 5, 1).to(DEV)
    b = torch.roll(a, (2, -3), (2, 3))
    with torch.no_grad():
        flow = m(a.tile(1, 3, 1, 1), b.tile(1, 3, 1, 1))
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    assert tuple(flow.shape) == (1, 2, 1152, 1920) and torch.isfinite(flow).all()
