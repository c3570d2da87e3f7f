"""ORACLE — TEST INFRASTRUCTURE ONLY (see gpsd_common.py header).

numpy restatement of the row-wise periodic Hessenberg reduction of the reference
(src/rhessx.jl:7-50 RHouseholder lmul!/rmul!, :53-109 _rphessenberg!), math orientation,
real or complex.  Pinned by reconstruction / orthogonality / structure predicates in
tests/test_oracle_rowhess.py (the reference itself pins it only indirectly through
test/krylov.jl:101-121)."""
from __future__ import annotations

import numpy as np

from .gpsd_common import reflector


def _rh_lmul(v, tau, A):
    """lmul!(H::RHouseholder, A) (rhessx.jl:20-35): pivot is the LAST row of A."""
    m = A.shape[0]
    va = A[m - 1, :] + np.conj(v[:m - 1]) @ A[:m - 1, :]
    va = np.conj(tau) * va
    A[m - 1, :] -= va
    A[:m - 1, :] -= np.outer(v[:m - 1], va)


def _rh_rmul_adj(A, v, tau):
    """rmul!(A, H') (rhessx.jl:37-50): pivot is the LAST column of A."""
    n = A.shape[1]
    x = A[:, :n - 1] @ v + A[:, n - 1]
    A[:, n - 1] -= tau * x
    A[:, :n - 1] -= tau * np.outer(x, np.conj(v))


def rphessenberg(Ap, A, Q):
    """_rphessenberg!(Ap, A, Q) in place.  Ap m x n (m = n or n+1), A list of p-1 n x n,
    Q list of p (rows x n) or None."""
    p = len(A) + 1
    m, n = Ap.shape
    if m not in (n, n + 1):
        raise ValueError("only implemented for square or 1 extra row")

    def rowstep(X, i, kc, Qm, L):
        xi = np.conj(X[i - 1, kc - 1::-1][:kc]).copy()  # conj.(X[i, kc:-1:1])
        t = reflector(xi)
        xr = xi[:0:-1].copy()                            # xi[kc:-1:2]
        _rh_lmul(xr, t, L[:kc, :])
        _rh_rmul_adj(X[:, :kc], xr, t)
        if Qm is not None:
            _rh_rmul_adj(Qm[:, :kc], xr, t)

    Ax = Ap if p == 1 else A[p - 2]
    if m == n + 1:
        rowstep(Ap, n + 1, n, None if Q is None else Q[p - 1], Ax)
    for i in range(n, 1, -1):
        for l in range(p - 1, 0, -1):
            L = Ap if l == 1 else A[l - 2]
            rowstep(A[l - 1], i, i, None if Q is None else Q[l - 1], L)
        rowstep(Ap, i, i - 1, None if Q is None else Q[p - 1], Ax)
    Ap[...] = np.triu(Ap, -1)
    for l in range(p - 1):
        A[l][...] = np.triu(A[l])
    return Ap, A
