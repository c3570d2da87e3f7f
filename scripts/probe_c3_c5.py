import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, psd_b200, psd_rng
h=psd_b200.Handle([0])
S=[k%2 for k in range(10)]
A=psd_rng.gen_uniform(1234,64,10,2); psd_b200.gpschur_batched(A,S,"L",handle=h)
def run(A,S,lr):
    h.set_profiling(True); h.kernel_times()
    t0=time.time()
    out=psd_b200.gpschur_batched(A,S,lr,handle=h)
    t1=time.time()-t0
    kt=h.kernel_times()
    return round(kt['iterate_ms']/1e3,3), round(t1,3), int((out[5]!=0).sum())
A5=psd_rng.gen_uniform(1234,512,10,256)
r=run(A5,S,"L"); print('C5 256',r,'problems/s %.2f'%(256/r[0]),flush=True)
del A5
A3=psd_rng.gen_uniform(1234,128,6,592)+1j*psd_rng.gen_uniform(1234,128,6,592,0,1)
r=run(A3,[1,0,1,1,0,1],"R"); print('C3 592',r,'problems/s %.1f'%(592/r[0]),flush=True)
