"""How far ahead of the GPU does the host run?  Wall time to ENQUEUE one 96-frame step (no synchronisation) against its GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P
from learned_pmctf_b200 import gop as G
torch.manual_seed(0)
m = P.pMCTF(num_me_stages=4).cuda().eval()
codec = G.GopCodec(m, 16, q_index=12)
y_u8, c_u8 = G.synthetic_sequence(0, 96, 1080, 1920, "cuda")
Y = P.ops.unpack_u8(y_u8, 1152, 1920)
C = P.ops.unpack_u8(c_u8.view(-1, 540, 960), 576, 960).view(96, 2, 1, 576, 960)
mvs = [G.synthetic_motion(0, g, 16, 1152, 1920, "cuda") for g in range(6)]
for _ in range(2):
    codec.code_sequence(Y, C, mvs, y_u8, c_u8)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    codec.code_sequence(Y, C, mvs, y_u8, c_u8)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0):.1f} ms, until GPU done {1e3 * (t2 - t0):.1f} ms, launches {P._native.lib().pmctf_launch_count()}")
