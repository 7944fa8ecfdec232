"""Host enqueue time vs total time of the pieces of one full-model frame pair at 1080p (diagnostic)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import learned_pmctf_b200 as pkg
from test_pwave_coder import _randomise
dev = torch.device("cuda:0")
m = pkg.pMCTF(num_me_stages=4, entropy_model=True, motion=True)
_randomise(m.lp_coder, 1), _randomise(m.hp_coder, 2)
m = m.to(dev).eval()
y0 = (torch.nn.functional.avg_pool2d(torch.rand((1, 1, 1156, 1924), device=dev), 5, 1) * 255).round().contiguous()
y1 = torch.roll(y0, (1, -2), (2, 3)).contiguous()
c0 = torch.cat([torch.nn.functional.avg_pool2d(y0, 2)] * 2, 0).contiguous()
c1 = torch.cat([torch.nn.functional.avg_pool2d(y1, 2)] * 2, 0).contiguous()
dpb = {"mv_feature": None, "ref_mv_y": None}


def t(name, fn, reps=3):
    with torch.no_grad():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a = time.perf_counter()
        for _ in range(reps):
            fn()
        b = time.perf_counter()
        torch.cuda.synchronize()
        c = time.perf_counter()
    print(f"{name:28s} host {1e3 * (b - a) / reps:7.2f} ms   total {1e3 * (c - a) / reps:7.2f} ms")


t("compute_and_code_motion", lambda: m.compute_and_code_motion(y0, y1, 12, dpb, stage_idx=0))
mv = m.compute_and_code_motion(y0, y1, 12, dpb, stage_idx=0)[0]
t("forward_MCTF luma", lambda: m.forward_MCTF(y0, y1, mv, 0))
t("hp_coder.forward luma", lambda: m.hp_coder.forward(y1, 12))
t("hp_coder.forward chroma x2", lambda: m.hp_coder.forward(c1, 12))
t("encode_one_stage (pair)", lambda: m.encode_one_stage([y0, c0], [y1, c1], False, dpb, stage_idx=0, q_index=12))
