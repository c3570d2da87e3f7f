"""Shared helpers for the row-wise periodic Hessenberg tests (rhessx.jl:53-109)."""
import numpy as np

import psd_checks as K
import psd_rng

EPS = np.finfo(float).eps


def make(seed, n, p, extra, batch):
    """storage arrays: Ap [batch][n][m], A [batch][p-1][n][n] or None, Q [batch][p][n][n]"""
    m = n + (1 if extra else 0)
    big = psd_rng.gen_uniform(seed, m, p, batch)  # [batch][p][m][m] storage
    Ap = np.ascontiguousarray(big[:, 0, :n, :m])
    A = np.ascontiguousarray(big[:, 1:, :n, :n]) if p > 1 else None
    Q = np.zeros((batch, p, n, n))
    Q[:, :, np.arange(n), np.arange(n)] = 1.0
    return Ap, A, Q


def check(Ap0, A0, Ap, A, Q, n, p):
    """structure, orthogonality and reconstruction for one problem (storage arrays):
    Ap Q_p' ... with  Ap_new = Q_1' Ap0 Q_p (top n rows), A_l_new = Q_{l+1}' A_l0 Q_l."""
    Hp = K.M(Ap)
    assert not np.tril(Hp[:n], -2).any()
    if Hp.shape[0] == n + 1:
        assert not Hp[n, :n - 1].any()
    Qm = [K.M(Q[l]) for l in range(p)]
    for l in range(p):
        assert np.linalg.norm(Qm[l] @ Qm[l].T - np.eye(n)) < 10 * EPS * n
    # left transforms: factor l (1..p-1) is hit from the left by the reflectors of factor l+1,
    # Ap (factor p) from the left by those of factor 1
    X0 = K.M(Ap0)
    rec = Qm[0].T @ X0[:n] @ Qm[p - 1] if p > 1 else Qm[0].T @ X0[:n] @ Qm[0]
    assert np.linalg.norm(rec - Hp[:n]) < 30 * EPS * n * max(1.0, np.linalg.norm(X0))
    if Hp.shape[0] == n + 1:
        assert np.linalg.norm(X0[n] @ Qm[p - 1] - Hp[n]) < 30 * EPS * n * max(1.0, np.linalg.norm(X0))
    for l in range(1, p):
        Tl = K.M(A[l - 1])
        assert not np.tril(Tl, -1).any()
        rec = Qm[l].T @ K.M(A0[l - 1]) @ Qm[l - 1]
        assert np.linalg.norm(rec - Tl) < 30 * EPS * n * max(1.0, np.linalg.norm(A0[l - 1]))
