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


def graded_storage(seed, n, p, batch, k, S):
    """Real factors graded in row PAIRS: factor j is D_j * B_j with B_j = uniform + 0.25 (the
    2x2 diagonal blocks of the first factor made rotation-like so that complex pairs occur) and D_j = diag(10^(-+e_r)), e_r = round(k * (r // 2) /
    max(1, n/2 - 1)), sign by S[j] (rows decay downwards for factors that enter directly, grow
    downwards for factors that enter inverted); for n = 2 the whole factor is scaled by
    10^(-+round(k (j+1) / p)).  The eigenvalues of prod A_j^{s_j} then come in 1x1 / 2x2 groups
    spread over ~ p*k decades: each 2x2 block is well conditioned in itself but can only be
    resolved factor by factor, never from an explicitly formed product.  Deterministic (IEEE
    multiplications of the counter-based uniforms by correctly rounded powers of ten)."""
    A = rand_storage(seed, n, p, batch, False) + 0.25
    for i in range(0, n - 1, 2):
        A[:, 0, i, i + 1] *= -1.0                       # math entry (i+1, i) of the Schur factor
    for j in range(p):
        sg = -1 if S[j] else 1
        for r in range(n):
            if n == 2:
                e = round(k * (j + 1) / p)
            else:
                e = round(k * (r // 2) / max(1, n // 2 - 1))
            A[:, j, :, r] *= 10.0 ** (sg * e)          # scale math-row r
    return A
