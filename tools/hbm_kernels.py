"""Achieved HBM bandwidth of the stand-alone (module-API) HBM-bound kernels: algorithmic bytes / CUDA-event time, against the
measured copy bandwidth of MEASURED_PEAKS.json.  Inputs larger than L2 (16 luma frames), 20 timed launches after 5 warm-ups.
    python tools/hbm_kernels.py  ->  one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_pmctf_b200 as P

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
N, H, W = 16, 1152, 1920
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
im = torch.rand((N, 1, H, W), device=dev, generator=g) * 255
mv = torch.nn.functional.avg_pool2d(4.0 * torch.randn((N, 2, H, W), device=dev, generator=g), 5, 1, 2).mul_(5).clamp_(-32, 32).contiguous()
u8 = (torch.rand((N, 1080, 1920), device=dev, generator=g) * 255).to(torch.uint8)
stats = torch.zeros((N, 2), dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()                      # L2 flush between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


px = N * H * W
cases = {
    # name: (callable, algorithmic bytes)
    "flow_warp (video_net.py:32-55): src 4 + flow 8 + out 4 B/px": (lambda: P.ops.flow_warp(im, mv), 16 * px),
    "chroma_mv_down (video_net.py:66-71): 8 in + 2 out B per luma px": (lambda: P.ops.chroma_mv_down(mv), 10 * px),
    "quantize_stats (pWave.py:184-189 + symbol statistics): 4 + 4 B/coeff": (lambda: P.ops.quantize_stats(im, 0.37, stats), 8 * px),
    "dequantize (pWave.py:191-202): 4 + 4 B/coeff": (lambda: P.ops.dequantize(im, 0.37), 8 * px),
    "unpack_u8 (test_pMCTF_flex.py:151-192): 1 B in (un-padded) + 4 B out (padded)": (lambda: P.ops.unpack_u8(u8, H, W), N * 1080 * 1920 + 4 * px),
    "frame_sse (test_pMCTF_flex.py:300-310): 4 + 1 B per un-padded px": (lambda: P.ops.frame_sse(im, u8), 5 * N * 1080 * 1920),
}
out = {}
for name, (fn, nbytes) in cases.items():
    ms = timed(fn)
    out[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "frac_of_measured_copy_peak": round(nbytes / ms / 1e6 / peak, 3)}
print(json.dumps({"hbm_peak_gbs": peak, "planes": [N, H, W], "kernels": out}))
