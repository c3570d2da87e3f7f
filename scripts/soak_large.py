"""Soak run of the large-N path: several orders / periods / seeds, status and residual of each
(the shift selection is timing dependent, so repeated runs exercise different schedules).
Usage: python scripts/soak_large.py [repeats]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import psd_b200  # noqa: E402
import psd_checks as K  # noqa: E402

EPS = np.finfo(float).eps
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(7)
worst = 0.0
bad = 0
for n, p in [(192, 4), (333, 3), (512, 8), (777, 2), (1024, 4), (1500, 5), (2048, 3), (2600, 4)]:
    for rep in range(reps):
        A = rng.uniform(-1.0, 1.0, size=(1, p, n, n))
        if rep % 2 == 1:  # graded: rows scaled over 6 orders of magnitude
            A *= np.logspace(-3, 3, n)[None, None, :, None]
        t0 = time.time()
        T, Z, lam, info = psd_b200.pschur_batched(A, "R")
        dt = time.time() - t0
        st = psd_b200.default_handle().large_stats()
        try:
            r = K.pschur_check(A[0], T[0], Z[0], lam[0], tol=1e9, check_lambda=False, baseline_gates=False)
            res = r["residual_eps_a1"] / n
        except AssertionError as ex:
            print("   check failed:", str(ex)[:100])
            res = float("inf")
        worst = max(worst, res)
        ok = info[0] == 0 and st["status"] == 0 and res < 10
        bad += not ok
        print(f"n {n} p {p} rep {rep}: {dt:.2f} s info {info[0]} status {st['status']} rounds {st['rounds']} pairs {st['shift_pairs']} "
              f"exceptional {st['exceptional']} residual/(n eps) {res:.3f} {'' if ok else 'FAIL'}", flush=True)
print("worst residual / (n eps):", worst, "failures:", bad)
sys.exit(1 if bad else 0)
