"""Timing of the four-step entropy-parameter network (section 8f row 1) on one B200: the CTA-pair tensor-core layer alone and the
whole module on the 12 subbands of a 4-level decomposition of a 1080p luma plane; stock torch (cuDNN) beside it.
    python tools/bench_ctx.py [--no-stock]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ctx_weights  # noqa: E402

import learned_pmctf_b200 as pkg  # noqa: E402
from learned_pmctf_b200 import _native as nat  # noqa: E402
from learned_pmctf_b200.layers.context_fusion_4step import ContextFusionFourStep  # noqa: E402

LAYER_FLOPS = 2 * 9 * 112 * 112
MODULE_FLOPS = 22 * LAYER_FLOPS + 2 * 112 * 112 + 2 * 9 * 112 * 5 + 2 * 9 * 112 + 2 * 112 * 2 * 5   # per coefficient


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    lib = nat.lib()
    out = {}
    st = torch.cuda.current_stream().cuda_stream
    w = torch.randn(112, 112, 3, 3, device=dev) * 0.03
    b = torch.zeros(112, device=dev)
    packed = torch.empty(int(lib.pmctf_ctx_packed_bytes(9)), dtype=torch.uint8, device=dev)
    nat.check(lib.pmctf_ctx_pack_conv(w.data_ptr(), 9, packed.data_ptr(), st), "pack")
    prof = "--profile" in sys.argv          # under ncu: one level-0 subband, one launch of each layer type, one module pass
    for (N, H, W) in (((1, 576, 960),) if prof else ((1, 576, 960), (4, 576, 960), (1, 144, 240), (1, 72, 120))):
        x = torch.randn(N, 14, H, W, 8, device=dev).to(torch.bfloat16)
        res = torch.randn(N, 28, H, W, 4, device=dev)
        of, ob = torch.empty_like(res), torch.empty_like(x)
        for tag, r, o32 in (("bf16_out", None, None), ("res_f32_bf16_out", res, of)):
            ms = timed(lambda: nat.check(lib.pmctf_ctx_conv112(x.data_ptr(), packed.data_ptr(), 9, b.data_ptr(), r.data_ptr() if r is not None else None,
                                                               None, 1.0, o32.data_ptr() if o32 is not None else None, ob.data_ptr(), N, H, W, st), "conv"),
                       1 if prof else 10, warm=0 if prof else 2)
            out[f"layer_{N}x{H}x{W}_{tag}"] = {"ms": ms, "tflops": LAYER_FLOPS * N * H * W / ms / 1e9}
    assert pkg.ops.tc_error_flag() == 0
    # whole module on the 12 subbands of one 1080p luma plane
    mods = {}
    torch.manual_seed(0)
    inputs = []
    for lvl in range(4):
        h, wd = 576 >> lvl, 960 >> lvl
        for band in ("lh", "hl", "hh"):
            cc = 2 if lvl < 3 else 1
            m = ContextFusionFourStep(ctx_channels=cc).to(dev).eval()
            m.load_state_dict({k: torch.from_numpy(v) for k, v in ctx_weights.make(7, cc).items()})
            mods[(lvl, band)] = m
            x = torch.round(torch.randn(1, 1, h, wd, device=dev) * 4)
            c = torch.tanh(torch.randn(1, 1, h, wd, device=dev))
            p = torch.round(torch.randn(1, 1, h // 2, wd // 2, device=dev) * 2) if cc == 2 else None
            inputs.append((mods[(lvl, band)], x, c, p))
    coeffs = sum(x.numel() for _, x, _, _ in inputs)

    def ours():
        for m, x, c, p in inputs:
            m(x, context=c, prev_subband=p)
    with torch.no_grad():
        ms = timed(ours, 1 if prof else 3, warm=1 if prof else 2)
    out["module_12_subbands_1080p_luma"] = {"ms": ms, "coefficients": coeffs, "tflops": MODULE_FLOPS * coeffs / ms / 1e9}
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        ours()
    out["module_12_subbands_1080p_luma"]["host_enqueue_ms"] = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    assert pkg.ops.tc_error_flag() == 0
    if not prof:   # the whole coder of one plane: transform + LL model + long-term context + 12 four-step models + PostProcess (pWave.forward)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_pwave_coder import _randomise
        pw = _randomise(pkg.pWave(entropy_model=True)).to(dev).eval()
        xin = (torch.nn.functional.avg_pool2d(torch.rand((1, 1, 1156, 1924), device=dev), 5, 1, 0) * 255).round().contiguous()
        with torch.no_grad():
            ms_pw = timed(lambda: pw(xin, q_index=12), 3, warm=2)
        out["pwave_forward_1080p_luma"] = {"ms": ms_pw, "planes_per_s": 1e3 / ms_pw}
    if "--no-stock" not in sys.argv and not prof:
        def stock():
            for m, x, c, p in inputs:
                m._forward_torch(x, c, p, False)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                ms_s = timed(stock, 1, warm=1)
            out["stock_torch_" + ("tf32" if tf32 else "fp32")] = {"ms": ms_s, "tflops": MODULE_FLOPS * coeffs / ms_s / 1e9}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
