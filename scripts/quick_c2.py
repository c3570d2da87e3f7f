"""Quick device-resident timing of the C2 shape (p=8, N=32, eigenvalues only)."""
import ctypes as C
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import psd_b200
from oracle import oracle as O

n, p = int(os.environ.get("N", 32)), int(os.environ.get("P", 8))
B = int(os.environ.get("B", 20000))
wantT = int(os.environ.get("WANTT", 0)); wantZ = int(os.environ.get("WANTZ", 0))
A = O.gen_real(1234, n, p, B)
h = psd_b200.Handle([0])
dA0 = torch.from_numpy(A).cuda()
dA = dA0.clone()
dZ = torch.empty_like(dA) if wantZ else None
dE = torch.empty((B, n, 2), dtype=torch.float64, device="cuda")
dI = torch.empty(B, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream()
st = ts.cuda_stream
L = psd_b200.lib()
def run():
    psd_b200.capi.check(L.psd_rpschur_batched_dev(h.ptr, 0, C.c_void_p(st), n, p, B, 0, wantT, wantZ, 30,
        C.c_void_p(dA.data_ptr()), C.c_void_p(dZ.data_ptr()) if wantZ else None, C.c_void_p(dE.data_ptr()), C.c_void_p(dI.data_ptr())))
for it in range(3):
    dA.copy_(dA0); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ts):
        e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n={n} p={p} B={B} wantT={wantT} wantZ={wantZ}: {ms:.2f} ms  {B/ms*1e3:.0f} problems/s  fails={(dI!=0).sum().item()}")
