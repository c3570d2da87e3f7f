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
B5=int(sys.argv[1]) if len(sys.argv)>1 else 148
A5=psd_rng.gen_uniform(1234,512,10,B5)
A3=psd_rng.gen_uniform(1234,128,6,296)+1j*psd_rng.gen_uniform(1234,128,6,296,0,1)
for env in ({}, {'PSD_NO_DEEP':'1'}):
    os.environ.pop('PSD_NO_DEEP',None); os.environ.update(env)
    print(env,'C5',B5,run(A5,S,"L"),flush=True)
    print(env,'C3 296',run(A3,[1,0,1,1,0,1],"R"),flush=True)
