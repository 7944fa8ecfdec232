"""Kernel-time breakdown of one hot-path training step (torch.profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import learned_pmctf_b200 as P
from learned_pmctf_b200 import gop as G
torch.manual_seed(0)
m = P.pMCTF(num_me_stages=4).cuda().train()
with torch.no_grad():
    for p in m.parameters():
        if p.dim() == 4 and p.shape[-1] == 3:
            p.normal_(0, 0.08)
clips = torch.rand(8, 8, 1, 256, 256, device="cuda") * 255
mvs = [torch.randn((8 * (8 >> (s + 1)), 2, 256, 256), device="cuda") for s in range(3)]
def step():
    for p in m.parameters(): p.grad = None
    loss, _ = G.training_loss_hot_path(m, clips, mvs, q_index=8)
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
