"""Phase timing of the tensor-core lifting step (one CTA) on a 1080p luma plane.  Needs the timing variant of the library
(clock64 stamps compiled in): build it in the build container first,
    python -c "import learned_pmctf_b200 as P; P._native.build(True, defines=('PMCTF_TC_TIMING=1',), out=P._native.LIB_PATH.replace('.so', '_timing.so'))"
"""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMCTF_LIB"] = os.path.join(ROOT, "learned-pmctf_b200", "lib", "libpmctf_b200_timing.so")
import torch
import learned_pmctf_b200 as P
m = P.pMCTF(num_me_stages=4).cuda().eval()
x = torch.rand(1, 1, 1152, 1920, device="cuda") * 255
mv = torch.randn(1, 2, 1152, 1920, device="cuda") * 3
lib = P._native.lib()
out = (C.c_longlong * 16)()
for name, fn in (("temporal fwd", lambda: m.forward_MCTF(x, x, mv, 0, want_pred=False)),
                 ("lift2d fwd", lambda: m.hp_coder.wavelet_transform.forward_lift_2d_bands(x))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); lib.pmctf_tc_debug_times(out)
    fn(); torch.cuda.synchronize(); lib.pmctf_tc_debug_times(out)
    t = list(out)
    print(name, "phases (cycles): setup+src %d | conv1 %d | conv2+conv3 %d | final %d | total %d" %
          (t[1]-t[0], t[2]-t[1], t[4]-t[2], t[6]-t[4], t[6]-t[0]))
    print("   MMA issue conv2 %d conv3 %d cycles; epilogue warp0 waited %d / %d cycles; CTA total %d cycles for %d tiles = %d / tile" % (t[9]-t[8], t[11]-t[10], t[12], t[13], t[14], t[15], t[14] // max(t[15], 1)))
o = torch.zeros(3, dtype=torch.int64, device="cuda")
for v, name in ((0, "conv block pattern"), (1, "N=16"), (2, "N=48"), (3, "N=96"), (4, "N=48 disjoint chunks")):
    lib.pmctf_tc_mma_probe(v, 64, o.data_ptr(), None); torch.cuda.synchronize()
    a, b, n = o.tolist()
    print(f"probe {name:24s}: issue {a/n:6.1f} cyc/MMA, complete {b/n:6.1f} cyc/MMA ({n} MMAs)")
