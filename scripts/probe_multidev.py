#!/usr/bin/env python
"""One process, several devices: the library's in-process multi-GPU path (psd_rpschur_batched on a
handle that owns N devices) against the same call on one device each."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import psd_b200
L = psd_b200.lib()
n, p, B = 32, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nd = torch.cuda.device_count()
hA = torch.empty((nd * B, p, n, n), dtype=torch.float64, pin_memory=True)
psd_b200.capi.check(L.psd_fill_uniform_host(1234, n, p, nd * B, 0, 0, C.c_void_p(hA.data_ptr())))
hE = torch.empty((nd * B, n, 2), dtype=torch.float64, pin_memory=True)
hI = torch.empty(nd * B, dtype=torch.int32, pin_memory=True)


def run(devs, first, count, reps=2):
    h = psd_b200.Handle(devs)
    a = C.c_void_p(hA.data_ptr() + first * p * n * n * 8)
    e = C.c_void_p(hE.data_ptr() + first * n * 16)
    i = C.c_void_p(hI.data_ptr() + first * 4)
    call = lambda: psd_b200.capi.check(L.psd_rpschur_batched(h.ptr, n, p, count, 0, 0, 0, 30, a, None, e, i))
    call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    h.close()
    return count / dt, dt


for d in range(nd):
    r, dt = run([d], d * B, B)
    print(f"Handle([{d}])   {B} problems: {r:10.0f} problems/s  ({dt:.3f} s)", flush=True)
r, dt = run(list(range(nd)), 0, nd * B)
print(f"Handle({list(range(nd))}) {nd * B} problems: {r:10.0f} problems/s  ({dt:.3f} s)", flush=True)
# the same from two host threads, one single-device handle each (what torchrun ranks do, in one process)
import threading
res = [None] * nd
def worker(d):
    res[d] = run([d], d * B, B)
t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(d,)) for d in range(nd)]
[t.start() for t in th]; [t.join() for t in th]
print("threads, one single-device handle each:", [f"{x[0]:.0f}/s" for x in res], flush=True)
