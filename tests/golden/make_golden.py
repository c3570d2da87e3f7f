"""Generates tests/golden/real_golden.json.

The reference (pure Julia) cannot run in this image, so the golden vectors are
 (1) the reference's own known-answer family expsplit(p, T) with its literal matrix and
     asymptotic eigenvalues (test/testfuncs.jl:412-421, used at test/runtests.jl:68-87), and
 (2) seeded inputs from the repo's counter-based generator together with the eigenvalues of
     the explicitly formed product computed in 60-digit arithmetic (mpmath), which is the
     independent ground truth the reference's pschur_check uses in double precision
     (test/testfuncs.jl:122-141).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import psd_rng  # noqa: E402

mp.mp.dps = 60


def product_eigs_mp(A, left=False):
    """A: [p][col][row] storage.  eigenvalues of A1*...*Ap (or Ap*...*A1) in high precision."""
    p, n, _ = A.shape
    P = mp.eye(n)
    for j in range(p):
        Mj = mp.matrix(A[j].T.tolist())
        P = (Mj * P) if left else (P * Mj)
    ev = mp.eig(P, left=False, right=False)
    return [[float(mp.re(e)), float(mp.im(e))] for e in ev]


def expsplit(p):
    fac = 0.1
    A1 = np.array([[9, 4, 1, 4, 3, 4], [6, 8, 2, 4, 0, 2], [0, 7, 4, 4, 6, 6], [0, 0, 8, 4, 6, 7],
                   [0, 0, 0, 8, 9, 3], [0, 0, 0, 0, 5, 0]], dtype=np.float64)
    Aj = np.diag([fac, fac ** 2, fac ** 3, 1, 1, 1])
    lam = [[15.6284, 0.0], [-1.31418, -3.51424], [-1.31418, 3.51424], [90 * fac ** p, 0.0],
           [(1600 / 3) * fac ** (2 * p), 0.0], [-(71750 / 11) * fac ** (3 * p), 0.0]]
    return A1, Aj, lam


def main():
    out = {"seed": 1234, "cases": [], "expsplit": []}
    for (n, p, nb) in [(5, 1, 2), (5, 2, 2), (5, 3, 2), (5, 5, 2), (6, 4, 1), (16, 6, 1), (32, 8, 1)]:
        A = psd_rng.gen_uniform(1234, n, p, nb)
        for b in range(nb):
            for left in (False, True):
                out["cases"].append({"n": n, "p": p, "b": b, "left": left,
                                     "eig": product_eigs_mp(A[b], left)})
    for p in (5, 20):
        A1, Aj, lam = expsplit(p)
        # exact eigenvalues of the product in high precision as well
        P = mp.matrix(A1.tolist())
        D = mp.diag([mp.mpf("0.1"), mp.mpf("0.01"), mp.mpf("0.001"), 1, 1, 1])
        for _ in range(p - 1):
            P = P * D
        old = mp.mp.dps
        mp.mp.dps = 200
        P = mp.matrix(A1.tolist())
        for _ in range(p - 1):
            P = P * D
        ev = mp.eig(P, left=False, right=False)
        mp.mp.dps = old
        out["expsplit"].append({"p": p, "A1": A1.tolist(), "Aj_diag": [0.1, 0.01, 0.001, 1, 1, 1],
                                "lambda_reference_asymptotic": lam,
                                "lambda_mp": [[float(mp.re(e)), float(mp.im(e))] for e in ev]})
    with open(os.path.join(HERE, "real_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
