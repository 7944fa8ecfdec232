"""One temporal + one spatial tensor-core lifting step at 1080p for ncu (bench.py's weights, 4 planes per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P
import bench
m = bench.build_model(P, torch.device("cuda"))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.nn.functional.avg_pool2d(torch.rand(4, 1, 1152 + 8, 1920 + 8, device="cuda", generator=g), 9, 1, 0) * 255
x = x.round().contiguous()
mv = torch.randn(4, 2, 1152, 1920, device="cuda", generator=g) * 3
for _ in range(2):
    m.forward_MCTF(x, x.roll(3, -1), mv, 0, want_pred=False)
    m.hp_coder.wavelet_transform.forward_lift_2d_bands(x)
torch.cuda.synchronize()
print("done")
