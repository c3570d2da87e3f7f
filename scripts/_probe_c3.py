import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, psd_b200, psd_rng
h=psd_b200.Handle([0])
A=psd_rng.gen_uniform(1234,64,6,2)+0j; psd_b200.gpschur_batched(A,[1,0,1,1,0,1],"R",handle=h)
def run(A,S,lr):
    h.set_profiling(True); h.kernel_times()
    t0=time.time()
    out=psd_b200.gpschur_batched(A,S,lr,handle=h)
    t1=time.time()-t0
    kt=h.kernel_times()
    return round(kt['iterate_ms']/1e3,3), round(t1,3), int((out[5]!=0).sum())
A3=psd_rng.gen_uniform(1234,128,6,592)+1j*psd_rng.gen_uniform(1234,128,6,592,0,1)
for env in ({}, {'PSD_GEN_WIDE':'0'}, {'PSD_NO_WINDOWED_QZ':'1'}):
    for k in ('PSD_GEN_WIDE','PSD_NO_WINDOWED_QZ'): os.environ.pop(k,None)
    os.environ.update(env)
    B=592 if env.get('PSD_GEN_WIDE')!='0' else 296
    r=run(A3[:B],[1,0,1,1,0,1],"R"); print(env,'C3',B,r,'problems/s %.1f'%(B/r[0]),flush=True)
