import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, psd_b200, psd_rng
h=psd_b200.Handle([0])
S=[k%2 for k in range(10)]
A=psd_rng.gen_uniform(1234,64,10,2); psd_b200.gpschur_batched(A,S,"L",handle=h)
def run(A,S,lr):
    h.set_profiling(True); h.kernel_times()
    out=psd_b200.gpschur_batched(A,S,lr,handle=h)
    kt=h.kernel_times()
    return kt['iterate_ms']/1e3, int((out[5]!=0).sum())
for n,B in ((512,1),(512,148)):
    A=psd_rng.gen_uniform(1234,n,10,B)
    print('C5',n,B,run(A,S,"L"),flush=True)
A=psd_rng.gen_uniform(1234,128,6,296)+1j*psd_rng.gen_uniform(1234,128,6,296,0,1)
print('C3 296',run(A,[1,0,1,1,0,1],"R"))
