"""Host cost per launch: the same 96-frame step on tiny frames (the GPU work per launch is minimal), enqueue wall time / launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P
from learned_pmctf_b200 import gop as G
torch.manual_seed(0)
m = P.pMCTF(num_me_stages=4).cuda().eval()
lib = P._native.lib()
for (h0, w0) in ((128, 128), (256, 256)):
    codec = G.GopCodec(m, 16, q_index=12)
    y_u8, c_u8 = G.synthetic_sequence(0, 96, h0, w0, "cuda")
    Y = P.ops.unpack_u8(y_u8, h0, w0)
    C = P.ops.unpack_u8(c_u8.view(-1, h0 // 2, w0 // 2), h0 // 2, w0 // 2).view(96, 2, 1, h0 // 2, w0 // 2)
    mvs = [G.synthetic_motion(0, g, 16, h0, w0, "cuda") for g in range(6)]
    for _ in range(2):
        codec.code_sequence(Y, C, mvs, y_u8, c_u8)
    torch.cuda.synchronize()
    for conc in (True, False):
        codec.concurrent_chroma = conc
        l0 = lib.pmctf_launch_count()
        t0 = time.perf_counter()
        codec.code_sequence(Y, C, mvs, y_u8, c_u8)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        n = lib.pmctf_launch_count() - l0
        print(f"{h0}x{w0} two_streams={conc}: enqueue {1e3 * (t1 - t0):.1f} ms, GPU done {1e3 * (t2 - t0):.1f} ms, {n} launches -> {1e6 * (t1 - t0) / n:.1f} us host per launch")
