import sys, numpy as np, torch
import torch.nn.functional as F
f32=np.float32
torch.manual_seed(3)
print(torch.backends.cpu.get_cpu_capability(), torch.get_num_threads())
mv = torch.randn(1, 2, 64, 96) * 5
ref = F.interpolate(mv, (32, 48), mode='bilinear', align_corners=False).numpy()
a = mv.numpy()
h1 = (f32(0.5) * a[:, :, :, 0::2] + f32(0.5) * a[:, :, :, 1::2])
v1 = (f32(0.5) * h1[:, :, 0::2] + f32(0.5) * h1[:, :, 1::2])
d = np.argwhere(v1 != ref)
print(len(d), v1.size)
n,c,y,x = d[0]
print(a[n,c,2*y:2*y+2,2*x:2*x+2], ref[n,c,y,x], v1[n,c,y,x])
A = a.astype(np.float64)
for name, val in {
 "sum4*0.25 f64": (0.25*(A[n,c,2*y,2*x]+A[n,c,2*y,2*x+1]+A[n,c,2*y+1,2*x]+A[n,c,2*y+1,2*x+1])),
}.items(): print(name, f32(val))
# candidates vectorised
def cand(order):
    v00=a[:,:,0::2,0::2]; v01=a[:,:,0::2,1::2]; v10=a[:,:,1::2,0::2]; v11=a[:,:,1::2,1::2]
    q=f32(0.25)
    if order=="w-products": # each weight product w_y*w_x = 0.25 then sequential sum
        return ((v00*q + v01*q) + v10*q) + v11*q
    if order=="f64": return (0.25*(v00.astype(np.float64)+v01+v10+v11)).astype(f32)
    if order=="pair-v": return (v00*q+v10*q)+(v01*q+v11*q)
for o in ("w-products","f64","pair-v"):
    print(o, np.array_equal(cand(o), ref), np.abs(cand(o)-ref).max())
