"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libpsdo.so, the CPU restatement (C++17 + OpenMP) of the
dense pschur! hot path of RalphAS/PeriodicSchurDecompositions.jl.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package never does.

The reference is pure Julia and no Julia runtime exists in this image, so parity of this
restatement is pinned by the reference's own test predicates and its known-answer family
(tests/test_oracle_*.py), not by outputs of the reference run here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpsdo.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -fopenmp)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in os.listdir(_HERE)
        if f.endswith((".hpp", ".cpp"))
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.psdo_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def max_threads() -> int:
    return int(lib().psdo_max_threads())


def gen_real(seed: int, n: int, p: int, batch: int, first_b: int = 0) -> np.ndarray:
    """A[batch][p][col][row] (each factor column-major), uniform [0,1)."""
    A = np.empty((batch, p, n, n), dtype=np.float64)
    lib().psdo_gen_real(C.c_uint64(seed), n, p, C.c_int64(batch), C.c_int64(first_b), _p(A))
    return A


def gen_complex(seed: int, n: int, p: int, batch: int, first_b: int = 0) -> np.ndarray:
    A = np.empty((batch, p, n, n), dtype=np.complex128)
    lib().psdo_gen_complex(C.c_uint64(seed), n, p, C.c_int64(batch), C.c_int64(first_b),
                           A.ctypes.data_as(C.POINTER(C.c_double)))
    return A


def rpschur_batched(A: np.ndarray, left: bool = False, wantT: bool = True, wantZ: bool = True,
                    maxitfac: int = 30, nthreads: int = 0):
    """Real standard periodic Schur of a batch.  A is [batch][p][n][n] with each factor stored
    column-major (i.e. A[b, j] is the TRANSPOSE of the math matrix as numpy sees it).
    Returns (T, Z, eig, info, iters); T overwrites a copy of A."""
    assert A.dtype == np.float64 and A.ndim == 4 and A.shape[2] == A.shape[3]
    batch, p, n, _ = A.shape
    T = np.ascontiguousarray(A).copy()
    Z = np.zeros_like(T) if wantZ else None
    eig = np.zeros((batch, n), dtype=np.complex128)
    info = np.zeros(batch, dtype=np.int32)
    iters = np.zeros(batch, dtype=np.int32)
    rc = lib().psdo_rpschur_batched(n, p, C.c_int64(batch), int(left), int(wantT), int(wantZ),
                                    maxitfac, _p(T), _p(Z),
                                    eig.ctypes.data_as(C.POINTER(C.c_double)),
                                    _p(info, C.c_int32), _p(iters, C.c_int32), nthreads)
    assert rc == 0
    return T, Z, eig, info, iters


def rphess_batched(A: np.ndarray, wantQ: bool = True, nthreads: int = 0):
    batch, p, n, _ = A.shape
    H = np.ascontiguousarray(A).copy()
    Q = np.zeros_like(H) if wantQ else None
    rc = lib().psdo_rphess_batched(n, p, C.c_int64(batch), _p(H), _p(Q), nthreads)
    assert rc == 0
    return H, Q


def rpschur_hessut(H: np.ndarray, wantT: bool = True, wantZ: bool = True, maxitfac: int = 30):
    """Inner solver on Hessenberg/triangular input H[p][n][n] (column-major factors)."""
    p, n, _ = H.shape
    T = np.ascontiguousarray(H).copy()
    Z = np.zeros_like(T)
    eig = np.zeros(n, dtype=np.complex128)
    info = lib().psdo_rpschur_hessut(n, p, _p(T), _p(Z), eig.ctypes.data_as(C.POINTER(C.c_double)),
                                     int(wantT), int(wantZ), maxitfac)
    return T, (Z if wantZ else None), eig, int(info)
