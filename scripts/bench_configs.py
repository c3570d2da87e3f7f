#!/usr/bin/env python
"""Secondary measurements (not the bench.py headline): BASELINE configs 1, 3 and 5 through the
host C-ABI calls (H2D + kernels + D2H inside the timed region) with the kernel-only device time
from the library's CUDA-event timers.  Bounded batches; prints one JSON line per config.

  python scripts/bench_configs.py [--c3-batch 296] [--c5-batch 148]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psd_b200  # noqa: E402
import psd_rng  # noqa: E402

HBM = 6551.4
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def rand(seed, n, p, B, cplx):
    re = psd_rng.gen_uniform(seed, n, p, B, 0, 0)
    return re + 1j * psd_rng.gen_uniform(seed, n, p, B, 0, 1) if cplx else re


def timed(h, fn, reps):
    fn()  # warm-up (allocations, module load)
    h.set_profiling(True)
    h.kernel_times()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    dt = (time.perf_counter() - t0) / reps
    kt = h.kernel_times()
    h.set_profiling(False)
    return dt, (kt["reduce_ms"] + kt["iterate_ms"]) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-batch", type=int, default=296)
    ap.add_argument("--c5-batch", type=int, default=148)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    h = psd_b200.Handle([0])
    res = []
    # C1: real p=3 N=50 :R with Schur vectors, latency of one problem and a 1184-problem batch
    for B in (1, 1184):
        A = rand(1234, 50, 3, B, False)
        dt, kms, out = timed(h, lambda: psd_b200.pschur_batched(A, "R", handle=h), 3)
        assert (out[3] == 0).all()
        byts = 180800 * B
        res.append({"config": "C1 real pschur! p=3 N=50 :R with Z", "batch": B, "e2e_s": dt,
                    "kernel_ms": kms, "problems_per_s_e2e": B / dt,
                    "problems_per_s_kernel": B / (kms * 1e-3),
                    "hbm_gbs_kernel": byts / (kms * 1e-3) / 1e9, "hbm_frac": byts / (kms * 1e-3) / 1e9 / HBM,
                    "gflops_standard_count": 25 * 3 * 50 ** 3 * B / (kms * 1e-3) / 1e9})
    # C3: complex GPSD p=6 N=128 S=[T,F,T,T,F,T] :R with Z
    B = a.c3_batch
    A = rand(1234, 128, 6, B, True)
    S = [1, 0, 1, 1, 0, 1]
    dt, kms, out = timed(h, lambda: psd_b200.gpschur_batched(A, S, "R", handle=h), a.reps)
    assert (out[5] == 0).all()
    byts = 4723712 * B
    res.append({"config": "C3 complex GPSD p=6 N=128 mixed S :R with Z", "batch": B, "e2e_s": dt,
                "kernel_ms": kms, "problems_per_s_e2e": B / dt, "problems_per_s_kernel": B / (kms * 1e-3),
                "hbm_gbs_kernel": byts / (kms * 1e-3) / 1e9, "hbm_frac": byts / (kms * 1e-3) / 1e9 / HBM,
                "gflops_standard_count": 4 * 33 * 6 * 128 ** 3 * B / (kms * 1e-3) / 1e9})
    # C5: real GPSD p=10 N=512 alternating S :L with Z
    B = a.c5_batch
    A = rand(1234, 512, 10, B, False)
    S = [k % 2 for k in range(10)]
    dt, kms, out = timed(h, lambda: psd_b200.gpschur_batched(A, S, "L", handle=h), 1)
    assert (out[5] == 0).all()
    byts = 62930944 * B
    res.append({"config": "C5 real GPSD p=10 N=512 alternating S :L with Z", "batch": B, "e2e_s": dt,
                "kernel_ms": kms, "problems_per_s_e2e": B / dt, "problems_per_s_kernel": B / (kms * 1e-3),
                "hbm_gbs_kernel": byts / (kms * 1e-3) / 1e9, "hbm_frac": byts / (kms * 1e-3) / 1e9 / HBM,
                "gflops_standard_count": 33 * 10 * 512 ** 3 * B / (kms * 1e-3) / 1e9})
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
