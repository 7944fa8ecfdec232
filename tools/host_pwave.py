import os, sys, time, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, os.path.join(ROOT,'tests')]
import learned_pmctf_b200 as pkg
from test_pwave_coder import _randomise
dev=torch.device('cuda:0')
pw=_randomise(pkg.pWave(entropy_model=True)).to(dev).eval()
x=(torch.nn.functional.avg_pool2d(torch.rand((1,1,1156,1924),device=dev),5,1,0)*255).round().contiguous()
with torch.no_grad():
    for _ in range(3): pw(x,q_index=12)
    torch.cuda.synchronize()
    t0=time.perf_counter(); out=pw(x,q_index=12); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print('host enqueue ms', (t1-t0)*1e3, 'total ms', (t2-t0)*1e3)
    # CUDA graph of the whole forward
    try:
        g=torch.cuda.CUDAGraph()
        s=torch.cuda.Stream()
        with torch.cuda.stream(s):
            pw(x,q_index=12)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            out2=pw(x,q_index=12)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        g.replay(); torch.cuda.synchronize()
        e0.record(); 
        for _ in range(3): g.replay()
        e1.record(); torch.cuda.synchronize()
        print('graph replay ms', e0.elapsed_time(e1)/3, 'x_hat equal', torch.equal(out['x_hat'], out2['x_hat']), 'tc flag', pkg.ops.tc_error_flag())
    except Exception as ex:
        print('graph capture failed:', repr(ex)[:300])
