"""Generates tests/golden/graded_golden.json: eigenvalues of prod A_j^{s_j} in 120-digit
arithmetic (mpmath) for small REAL generalized periodic problems with strongly graded factors
(tests/gpsd_cases.py:graded_storage).  These pin the relative accuracy of the 2x2-block kernels
of the real periodic QZ (reference: rpschur2x2.jl:9-317 _rpeigvals2x2/_rp2x2ssr!,
rgeneralized.jl:1140-1509 _qzrots/_shift2rot), whose whole point (SLICOT MB03AF/MB03BD) is that
no product - in particular no product involving inverses - is ever formed.
Run:  python tests/golden/make_golden_graded.py
"""
import json
import os
import sys

import mpmath as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gpsd_cases as GCs  # noqa: E402

mp.mp.dps = 120
SEED = 9091
CASES = [
    # (n, p, S, k)
    (2, 3, [1, 1, 1], 0), (2, 3, [1, 0, 1], 20), (2, 4, [1, 0, 1, 0], 60), (2, 6, [1, 0, 1, 1, 0, 1], 40),
    (2, 6, [1, 1, 1, 1, 1, 1], 45), (4, 4, [1, 0, 1, 0], 8), (4, 5, [1, 1, 0, 1, 0], 12),
    (6, 6, [1, 0, 1, 1, 0, 1], 10), (6, 3, [1, 1, 1], 24), (8, 4, [1, 1, 0, 1], 15),
]
NB = 4


def product_eigs_mp(A, S):
    p, n, _ = A.shape
    P = mp.eye(n)
    for j in range(p):
        Mj = mp.matrix([[mp.mpf(float(x)) for x in row] for row in A[j].T.tolist()])
        Fj = Mj if S[j] else mp.inverse(Mj)
        P = P * Fj
    ev = mp.eig(P, left=False, right=False)
    return [[mp.nstr(mp.re(e), 25), mp.nstr(mp.im(e), 25)] for e in ev]


def main():
    out = {"seed": SEED, "cases": []}
    for (n, p, S, k) in CASES:
        A = GCs.graded_storage(SEED, n, p, NB, k, S)
        for b in range(NB):
            out["cases"].append({"n": n, "p": p, "S": S, "k": k, "b": b, "eig": product_eigs_mp(A[b], S)})
    with open(os.path.join(HERE, "graded_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
