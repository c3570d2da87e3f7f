#!/usr/bin/env python
"""FP64 GEMM measurements: the hand-written DMMA kernel (csrc/psd_dgemm.cuh) on the shapes of the
large-N trailing updates, and the cuBLAS DGEMM denominator (torch.matmul float64, 8192^3, best of
10) that SURVEY.md section 8(d) asks for.  One JSON line per measurement."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import psd_b200  # noqa: E402


def ours(h, ta, tb, M, N, K, reps=10):
    rng = np.random.default_rng(0)
    A = np.asfortranarray(rng.standard_normal((K, M) if ta else (M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if tb else (K, N)))
    Cm = np.zeros((M, N), order="F")
    ms = C.c_double(0.0)
    psd_b200.capi.check(psd_b200.lib().psd_dgemm_host(
        h.ptr, ta, tb, M, N, K, 1.0, C.c_void_p(A.ctypes.data), A.shape[0], C.c_void_p(B.ctypes.data),
        B.shape[0], 0.0, C.c_void_p(Cm.ctypes.data), M, reps, C.byref(ms)))
    return ms.value


def main():
    import torch
    h = psd_b200.Handle([0])
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    print(json.dumps({"kernel": "cuBLAS DGEMM 8192^3 (torch.matmul float64), best of 10", "ms": best, "tflops": peak}))
    del a, b
    shapes = [("NN 4096x4096x4096", 0, 0, 4096, 4096, 4096), ("NT 4096x4096x64 (A -= Y V')", 0, 1, 4096, 4096, 64),
              ("TN 64x4096x4096 (W = V' A, split-K)", 1, 0, 64, 4096, 4096), ("NN 4096x4096x64 (A -= V W)", 0, 0, 4096, 4096, 64),
              ("NN 4096x64x4096 (Q V, split-K)", 0, 0, 4096, 64, 4096), ("NN 8192x8192x8192", 0, 0, 8192, 8192, 8192)]
    for name, ta, tb, M, N, K in shapes:
        ms = ours(h, ta, tb, M, N, K, 5 if M * N * K > 1e11 else 20)
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        print(json.dumps({"kernel": "psd::dgemm_dmma_kernel " + name, "ms": ms, "tflops": tf, "frac_of_cublas": tf / peak}))


if __name__ == "__main__":
    main()
