"""Generates tests/golden/generalized_golden.json.

The reference's generalized tests (test/generalized.jl, test/testfuncs.jl:155-382) hold no
literal golden vectors: they compare against the eigenvalues of the explicitly formed product
prod A_j^{s_j} in double precision (testfuncs.jl:213-234).  The fixtures here are that same
ground truth computed in 60-digit arithmetic (mpmath) on seeded inputs of the repo's
counter-based generator, so that the oracle and the CUDA path are both pinned to numbers
neither of them produced.  Convention: :R -> A_1^{s_1} A_2^{s_2} ... A_p^{s_p}, :L -> the
reversed product (as tests/psd_checks.py:gproduct_eigvals).
Run:  python tests/golden/make_golden_generalized.py
"""
import json
import os
import sys

import mpmath as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gpsd_cases as GCs  # noqa: E402

mp.mp.dps = 60

CASES = [
    # (n, p, S, complex, left)
    (4, 2, [1, 0], False, False),
    (5, 3, [1, 1, 0], False, False),
    (5, 3, [0, 1, 1], False, True),
    (6, 4, [1, 0, 1, 0], False, False),
    (8, 5, [1, 0, 1, 1, 0], False, False),
    (8, 6, [1, 1, 1, 1, 1, 1], False, False),
    (4, 2, [1, 0], True, False),
    (5, 4, [1, 0, 1, 0], True, False),
    (5, 4, [0, 1, 0, 1], True, True),
    (8, 6, [1, 0, 1, 1, 0, 1], True, False),
    (6, 5, [1, 1, 1, 1, 1], True, True),
]
SEED = 4321


def product_eigs_mp(A, S, left):
    """A: [p][col][row] storage."""
    p, n, _ = A.shape
    P = mp.eye(n)
    for j in range(p):
        Mj = mp.matrix(A[j].T.tolist())
        Fj = Mj if S[j] else mp.inverse(Mj)
        P = (Fj * P) if left else (P * Fj)
    ev = mp.eig(P, left=False, right=False)
    return [[float(mp.re(e)), float(mp.im(e))] for e in ev]


def main():
    out = {"seed": SEED, "cases": []}
    for (n, p, S, cplx, left) in CASES:
        A = GCs.rand_storage(SEED, n, p, 2, cplx)
        for b in range(2):
            out["cases"].append({"n": n, "p": p, "S": S, "complex": cplx, "left": left, "b": b,
                                 "eig": product_eigs_mp(A[b], S, left)})
    with open(os.path.join(HERE, "generalized_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
