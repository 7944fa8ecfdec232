import os, sys, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, os.path.join(ROOT,'tests')]
import learned_pmctf_b200 as pkg
from test_pwave_coder import _randomise
dev=torch.device('cuda:0')
pw=_randomise(pkg.pWave(entropy_model=True)).to(dev).eval()
x=(torch.nn.functional.avg_pool2d(torch.rand((1,1,1156,1924),device=dev),5,1,0)*255).round().contiguous()
with torch.no_grad():
    pw(x,q_index=12); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pw(x,q_index=12); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
