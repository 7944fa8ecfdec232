"""One spatial + one temporal tensor-core lifting step at 1080p for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P
m = P.pMCTF(num_me_stages=4).cuda().eval()
x = torch.rand(1, 1, 1152, 1920, device="cuda") * 255
mv = torch.randn(1, 2, 1152, 1920, device="cuda") * 3
for _ in range(2):
    m.forward_MCTF(x, x, mv, 0, want_pred=False)
    m.hp_coder.wavelet_transform.forward_lift_2d_bands(x)
torch.cuda.synchronize()
print("done")
