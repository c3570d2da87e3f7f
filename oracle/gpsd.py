"""ORACLE — TEST INFRASTRUCTURE ONLY (see gpsd_common.py header).

Batched front end of the numpy restatements of the generalized periodic Schur paths, on the
same STORAGE layout as the C ABI ([batch][p][col][row], user factor order) so that tests can
hand identical arrays to the oracle and to the CUDA library."""
from __future__ import annotations

import numpy as np

from . import gpsd_complex as GC


def _math(Ab):
    return [np.ascontiguousarray(Ab[j].T) for j in range(Ab.shape[0])]


def _pack(F, p, n, dtype, wantZ):
    """result dict (reference struct fields) -> storage arrays in user factor order"""
    T = np.zeros((p, n, n), dtype=dtype)
    Z = np.zeros((p, n, n), dtype=dtype) if wantZ else None
    js = F["schurindex"]
    jt = 0
    for j in range(1, p + 1):
        if j == js:
            T[j - 1] = F["T1"].T
        else:
            T[j - 1] = F["T"][jt].T
            jt += 1
    if wantZ:
        for j in range(p):
            Z[j] = F["Z"][j].T
    return T, Z


def cpschur_batched(A, S, left=False, wantT=True, wantZ=True, hessut=False, maxitfac=30):
    """Complex path (generalized.jl:108-148 or, with hessut, the inner :166 entry).
    Returns (T, Z, alpha, beta, alphascale, info)."""
    batch, p, n, _ = A.shape
    T = np.zeros_like(A)
    Z = np.zeros_like(A) if wantZ else None
    alpha = np.zeros((batch, n), dtype=np.complex128)
    beta = np.zeros((batch, n), dtype=np.complex128)
    scale = np.zeros((batch, n), dtype=np.int64)
    info = np.zeros(batch, dtype=np.int32)
    for b in range(batch):
        Am = _math(A[b])
        if hessut:
            F = GC.cpqz(Am[0], Am[1:], [bool(x) for x in S], wantZ=wantZ, wantT=wantT,
                        maxitfac=maxitfac)
        else:
            F = GC.cpschur(Am, S, "L" if left else "R", wantZ=wantZ, wantT=wantT, maxitfac=maxitfac)
        Tb, Zb = _pack(F, p, n, np.complex128, wantZ)
        T[b] = Tb
        if wantZ:
            Z[b] = Zb
        alpha[b], beta[b], scale[b], info[b] = F["alpha"], F["beta"], F["alphascale"], F["info"]
    return T, Z, alpha, beta, scale, info


def rgpschur_batched(A, S, left=False, wantT=True, wantZ=True, hessut=False, maxitfac=120):
    """Real generalized path (rgeneralized.jl:3-45 or, with hessut, the inner :49 entry).
    Returns (T, Z, alpha, beta, alphascale, info)."""
    from . import gpsd_real as GR
    batch, p, n, _ = A.shape
    T = np.zeros_like(A)
    Z = np.zeros_like(A) if wantZ else None
    alpha = np.zeros((batch, n), dtype=np.complex128)
    beta = np.zeros((batch, n), dtype=np.float64)
    scale = np.zeros((batch, n), dtype=np.int64)
    info = np.zeros(batch, dtype=np.int32)
    for b in range(batch):
        Am = _math(A[b])
        if hessut:
            F = GR.rpqz(Am[0], Am[1:], [bool(x) for x in S], wantZ=wantZ, wantT=wantT,
                        maxitfac=maxitfac)
        else:
            F = GR.rgpschur(Am, S, "L" if left else "R", wantZ=wantZ, wantT=wantT, maxitfac=maxitfac)
        Tb, Zb = _pack(F, p, n, np.float64, wantZ)
        T[b] = Tb
        if wantZ:
            Z[b] = Zb
        alpha[b], beta[b], scale[b], info[b] = F["alpha"], F["beta"], F["alphascale"], F["info"]
    return T, Z, alpha, beta, scale, info
