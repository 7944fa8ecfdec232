"""Timing-only what-if builds (PMCTF_WHATIF): which resource bounds the tensor-core lifting step?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P
m = P.pMCTF(num_me_stages=4).cuda().eval()
x = torch.rand(4, 1, 1152, 1920, device="cuda") * 255
with torch.no_grad():
    for _ in range(2):
        m.hp_coder.wavelet_transform.forward_lift_2d_bands(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.hp_coder.wavelet_transform.forward_lift_2d_bands(x)
    e1.record(); torch.cuda.synchronize()
print(os.environ.get("PMCTF_LIB", "base"), "lift2d fwd 4x1080p: %.2f ms" % (e0.elapsed_time(e1) / 3))
