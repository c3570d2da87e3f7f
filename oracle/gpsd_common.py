"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of the elementary transformations used by the generalized periodic Schur
paths of RalphAS/PeriodicSchurDecompositions.jl (complex: src/generalized.jl, real:
src/rgeneralized.jl, src/rpschur2x2.jl).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/; the product (the CUDA
library) never does.

Conventions: matrices are numpy arrays in MATH orientation (row, col); the helpers take the
reference's 1-based inclusive indices so that the restatements in gpsd_complex.py /
gpsd_real.py can be compared with the Julia source line by line.

Parity pinning: the reference is pure Julia and no Julia runtime exists in this image, so these
restatements are pinned by the reference's own acceptance predicates (test/testfuncs.jl:155-382
gpschur_check, test/generalized.jl fixtures incl. the planted-zero "hole" cases), and by
eigenvalues of the explicitly formed product (tests/test_oracle_generalized.py).

The arithmetic of Givens / givensAlgorithm / lmul! / rmul! is Julia stdlib LinearAlgebra
(LAPACK xLARTG semantics; source not under /root/reference): [c s; -conj(s) c] [f; g] = [r; 0]
with c real.  Any valid (c, s, r) gives a correct algorithm.
"""
from __future__ import annotations

import math

import numpy as np

EPS = np.finfo(np.float64).eps
FLOATMIN = np.finfo(np.float64).tiny


def givens(f, g):
    """givensAlgorithm(f, g) -> (c, s, r); works for real and complex scalars."""
    if g == 0:
        return 1.0, 0.0 * g, f
    if f == 0:
        ag = abs(g)
        if isinstance(g, complex) or np.iscomplexobj(g):
            return 0.0, np.conj(g) / ag, ag + 0j
        return 0.0, 1.0, g
    if np.iscomplexobj(f) or np.iscomplexobj(g) or isinstance(f, complex) or isinstance(g, complex):
        f = complex(f)
        g = complex(g)
        f1 = abs(f)
        g1 = abs(g)
        h = math.hypot(f1, g1)
        ph = f / f1
        return f1 / h, ph * np.conj(g) / h, ph * h
    # real dlartg, sign convention c > 0 when |f| > |g|
    r = math.hypot(f, g)
    c = f / r
    s = g / r
    if abs(f) > abs(g) and c < 0:
        c, s, r = -c, -s, -r
    return c, s, r


def lmul_g(A, i1, i2, c, s, c0, c1):
    """lmul!(Givens(i1,i2,c,s), view(A, :, c0:c1))  (1-based, inclusive)."""
    if c1 < c0:
        return
    a1 = A[i1 - 1, c0 - 1:c1].copy()
    a2 = A[i2 - 1, c0 - 1:c1].copy()
    A[i1 - 1, c0 - 1:c1] = c * a1 + s * a2
    A[i2 - 1, c0 - 1:c1] = -np.conj(s) * a1 + c * a2


def rmul_gadj(A, j1, j2, c, s, r0, r1):
    """rmul!(view(A, r0:r1, :), Givens(j1,j2,c,s)')  (1-based, inclusive)."""
    if r1 < r0:
        return
    a1 = A[r0 - 1:r1, j1 - 1].copy()
    a2 = A[r0 - 1:r1, j2 - 1].copy()
    A[r0 - 1:r1, j1 - 1] = a1 * c + a2 * np.conj(s)
    A[r0 - 1:r1, j2 - 1] = -a1 * s + a2 * c


def opnorm1(A, r0, r1, c0, c1, upper=False):
    """opnorm(view(A, r0:r1, c0:c1), 1); upper=True wraps UpperTriangular first."""
    B = A[r0 - 1:r1, c0 - 1:c1]
    if upper:
        B = np.triu(B)
    if B.size == 0:
        return 0.0
    return float(np.max(np.sum(np.abs(B), axis=0)))


def safeprod(S, x0, v):
    """generalized.jl:939-976: x0^s1 * prod v_i^s_i as alpha / beta * 2^scale with
    |alpha| in [1,2) or 0, beta in {0,1}."""
    p = len(v) + 1
    is_c = np.iscomplexobj(x0) or isinstance(x0, complex)
    alpha = (1.0 + 0j) if is_c else 1.0
    beta = 1
    scale = 0
    for i in range(1, p + 1):
        xi = x0 if i == 1 else v[i - 2]
        if S[i - 1]:
            alpha = alpha * xi
        else:
            if xi == 0:
                beta = 0
            else:
                alpha = alpha / xi
        if abs(alpha) == 0:
            alpha = alpha * 0
            scale = 0
            if beta == 0:
                return alpha, beta, scale
        else:
            while abs(alpha) < 1.0:
                alpha *= 2.0
                scale -= 1
            while abs(alpha) >= 2.0:
                alpha /= 2.0
                scale += 1
    return alpha, beta, scale


def reflector(x):
    """householder.jl:66-156 (_xreflector!) on a 1-D numpy view; returns tau, overwrites x
    with (beta, v).  Real or complex (the complex n == 1 case is non-trivial)."""
    n = x.shape[0]
    is_c = np.iscomplexobj(x)
    if n < 1:
        return 0.0
    if not is_c:
        if n == 1:
            return 0.0
        alpha = float(x[0])
        xnorm = float(np.linalg.norm(x[1:]))
        if xnorm == 0:
            return 0.0
        beta = -math.copysign(math.hypot(alpha, xnorm), alpha)
        tau = (beta - alpha) / beta
        x[1:] *= 1.0 / (alpha - beta)
        x[0] = beta
        return tau
    alpha = complex(x[0])
    xnorm = float(np.linalg.norm(x[1:])) if n > 1 else 0.0
    if xnorm == 0 and alpha.imag == 0:
        return 0.0 + 0j
    beta = -math.copysign(math.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm ** 2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    if n > 1:
        x[1:] *= 1.0 / (alpha - beta)
    x[0] = beta
    return tau


def hh_lmul_adj(A, r0, c0, c1, v, tau):
    """lmul!(H', view(A, r0:r0+len(v), c0:c1)), H = I - tau [1;v][1;v]^H (householder.jl:222)."""
    if c1 < c0:
        return
    m = len(v) + 1
    B = A[r0 - 1:r0 - 1 + m, c0 - 1:c1]
    w = np.concatenate(([1.0], v))
    t = np.conj(tau) * (np.conj(w) @ B)
    B -= np.outer(w, t)


def hh_rmul(A, r0, r1, c0, v, tau):
    """rmul!(view(A, r0:r1, c0:c0+len(v)), H) (householder.jl:207)."""
    if r1 < r0:
        return
    m = len(v) + 1
    B = A[r0 - 1:r1, c0 - 1:c0 - 1 + m]
    w = np.concatenate(([1.0], v))
    t = tau * (B @ w)
    B -= np.outer(t, np.conj(w))


def phessenberg_householder(A):
    """phessenberg!(A) + explicit Q (PeriodicSchurDecompositions.jl:213-259, 136-140) for real
    or complex matrices, in place on the list A; returns Q (list) with Q_j' A_j Q_{j+1} = H_j."""
    p = len(A)
    n = A[0].shape[0]
    Q = [np.eye(n, dtype=A[0].dtype) for _ in range(p)]
    for i in range(1, n):
        for j in range(p, 1, -1):
            xi = A[j - 1][i - 1:, i - 1]
            t = reflector(xi)
            v = xi[1:].copy()
            hh_lmul_adj(A[j - 1], i, i + 1, n, v, t)
            hh_rmul(A[j - 2], 1, n, i, v, t)
            hh_rmul(Q[j - 1], 1, n, i, v, t)
            xi[1:] = 0
        xi = A[0][i:, i - 1]
        t = reflector(xi)
        v = xi[1:].copy()
        hh_lmul_adj(A[0], i + 1, i + 1, n, v, t)
        hh_rmul(A[p - 1], 1, n, i + 1, v, t)
        hh_rmul(Q[0], 1, n, i + 1, v, t)
        xi[1:] = 0
    return Q


def gphessenberg(A, S):
    """_phessenberg!(A, S) (generalized.jl:988-1082): Stage 1 QR/RQ of A[p..2], Stage 2 Givens
    Hessenberg reduction of A[1] propagated through every triangular factor.  In place on the
    list A (math orientation, S[0] must be True); returns Qs."""
    import scipy.linalg as sla
    if not S[0]:
        raise ValueError("The first entry in S must be true")
    p = len(A)
    n = A[0].shape[0]
    Qs = [np.eye(n, dtype=A[0].dtype) for _ in range(p)]
    # Stage 1 (:1009-1028)
    for l in range(p, 1, -1):
        if S[l - 1]:
            Qf, R = sla.qr(A[l - 1])
            if S[l - 2]:
                A[l - 2][...] = A[l - 2] @ Qf
            else:
                A[l - 2][...] = Qf.conj().T @ A[l - 2]
            Qs[l - 1][...] = Qs[l - 1] @ Qf
            A[l - 1][...] = np.triu(R)
        else:
            R, Qf = sla.rq(A[l - 1])
            if S[l - 2]:
                A[l - 2][...] = A[l - 2] @ Qf.conj().T
            else:
                A[l - 2][...] = Qf @ A[l - 2]
            Qs[l - 1][...] = Qs[l - 1] @ Qf.conj().T
            A[l - 1][...] = np.triu(R)
    # Stage 2 (:1034-1079)
    A1 = A[0]
    G = [None] * (n + 2)
    for j in range(1, n - 1):
        for i in range(n, j + 1, -1):
            c, s, r = givens(A1[i - 2, j - 1], A1[i - 1, j - 1])
            A1[i - 2, j - 1] = r
            A1[i - 1, j - 1] = 0
            lmul_g(A1, i - 1, i, c, s, j + 1, n)
            rmul_gadj(Qs[0], i - 1, i, c, s, 1, n)
            G[i] = (c, s)
        for l in range(p, 1, -1):
            Al = A[l - 1]
            if S[l - 1]:
                for i in range(n, j + 1, -1):
                    c, s = G[i]
                    rmul_gadj(Al, i - 1, i, c, s, 1, i)
                    c, s, r = givens(Al[i - 2, i - 2], Al[i - 1, i - 2])
                    Al[i - 2, i - 2] = r
                    Al[i - 1, i - 2] = 0
                    lmul_g(Al, i - 1, i, c, s, i, n)
                    G[i] = (c, s)
            else:
                for i in range(n, j + 1, -1):
                    c, s = G[i]
                    lmul_g(Al, i - 1, i, c, s, i - 1, n)
                    c, s, r = givens(Al[i - 1, i - 1], Al[i - 1, i - 2])
                    Al[i - 1, i - 1] = r
                    Al[i - 1, i - 2] = 0
                    rmul_gadj(Al, i, i - 1, c, np.conj(s), 1, i - 1)
                    G[i] = (c, -s)
            for i in range(n, j + 1, -1):
                c, s = G[i]
                rmul_gadj(Qs[l - 1], i - 1, i, c, s, 1, n)
        for i in range(n, j + 1, -1):
            c, s = G[i]
            rmul_gadj(A1, i - 1, i, c, s, 1, n)
    return Qs


def exceptional_fg(k: int):
    """Deterministic stand-in for the reference's rand(T, 2) exceptional shift
    (generalized.jl:782): golden-ratio sequence keyed by the iteration count."""
    g = 0.6180339887498949

    def fr(m):
        x = (4 * k + m + 1) * g
        return x - math.floor(x)
    return complex(fr(0), fr(1)), complex(fr(2), fr(3))
