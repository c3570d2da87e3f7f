import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, psd_b200, psd_rng
h=psd_b200.Handle([0])
S1=[1]+[k%2 for k in range(1,10)]
A=psd_rng.gen_uniform(1234,64,10,2); psd_b200.gphessenberg_batched(A,S1,True,handle=h)
n,B=512,int(sys.argv[1]) if len(sys.argv)>1 else 148
A=psd_rng.gen_uniform(1234,n,10,B)
def run(tag, env):
    for k in ('PSD_NO_DEEP','PSD_GEN_SKIP','PSD_NO_BLOCKED_STAGE1'): os.environ.pop(k,None)
    os.environ.update(env)
    h.set_profiling(True); h.kernel_times()
    out=psd_b200.gphessenberg_batched(A,S1,True,handle=h)
    kt=h.kernel_times()
    print(tag, 'reduce-only kernel s:', round(kt['iterate_ms']/1e3,3), flush=True)
run('deep', {})
run('nodeep', {'PSD_NO_DEEP':'1'})
run('nodeep skipR', {'PSD_NO_DEEP':'1','PSD_GEN_SKIP':'1'})
run('nodeep skipL', {'PSD_NO_DEEP':'1','PSD_GEN_SKIP':'2'})
run('nodeep skipZ', {'PSD_NO_DEEP':'1','PSD_GEN_SKIP':'4'})
run('nodeep skipall', {'PSD_NO_DEEP':'1','PSD_GEN_SKIP':'7'})
