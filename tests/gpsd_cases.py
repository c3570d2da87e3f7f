"""Fixture constructors restated from the reference's generalized tests
(test/generalized.jl:42-246; hole cases :79-151) on the shared counter-based generator."""
import numpy as np

import psd_rng


def rand_storage(seed, n, p, batch, cplx):
    re = psd_rng.gen_uniform(seed, n, p, batch, 0, 0)
    if not cplx:
        return re
    return re + 1j * psd_rng.gen_uniform(seed, n, p, batch, 0, 1)


def hessut_storage(seed, n, p, batch, cplx, hole=None):
    """A[1] = triu(rand, -1), A[j>1] = triu(rand); storage is [col][row], so the math-lower
    part is storage [c][r] with r > c + k.  hole = (factor, index) (1-based) plants an exact
    zero on that diagonal entry."""
    A = rand_storage(seed, n, p, batch, cplx)
    r = np.arange(n).reshape(1, n)
    c = np.arange(n).reshape(n, 1)
    for j in range(p):
        k = 1 if j == 0 else 0
        A[:, j][:, (r > c + k)] = 0
    if hole is not None:
        A[:, hole[0] - 1, hole[1] - 1, hole[1] - 1] = 0
    return A


# (S, hole) pairs of test/generalized.jl:79-151 with SINGLE_MINUS_SIG = false
HOLE_CASES = [
    ([1, 1, 0, 1, 0], (2, 3)),
    ([1, 1, 0, 1, 0], (4, 3)),
    ([1, 0, 1, 0, 1], (4, 2)),
    ([1, 0, 1, 0, 1], (4, 4)),
    ([1, 0, 1, 0, 1], (2, 2)),
    ([1, 0, 1, 0, 1], (2, 4)),
]
