"""GPU parity tests for the real generalized periodic Schur path (periodic QZ): CUDA library
through the C ABI against the numpy restatement of the reference (oracle/gpsd_real.py) on
identical seeded inputs, with the reference's predicates (test/testfuncs.jl:238-382) and
BASELINE.json's gates."""
import numpy as np
import pytest

import gpsd_cases as GCs
import psd_checks as K
from oracle import gpsd as OG

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _vals(a, b, s):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return a / b.astype(np.complex128) * np.exp2(s.astype(float))


def _pairs_ok(lam):
    j = 0
    while j < len(lam):
        if np.isfinite(lam[j]) and lam[j].imag != 0:
            assert lam[j].imag > 0 and j + 1 < len(lam)
            assert abs(lam[j + 1] - np.conj(lam[j])) <= 1e-12 * abs(lam[j])
            j += 2
        else:
            j += 1


def _compare(psd, A, S, left=False, hessut=False, oracle_check=True):
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L" if left else "R", hessut=hessut)
    assert (info == 0).all(), info
    assert be.dtype == np.float64
    n = A.shape[-1]
    if oracle_check:
        To, Zo, alo, beo, sco, io = OG.rgpschur_batched(A, S, left=left, hessut=hessut)
        assert (io == 0).all()
    for b in range(A.shape[0]):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left, real_path=True)
        _pairs_ok(r["values"])
        if oracle_check:
            lo = _vals(alo[b], beo[b], sco[b])
            worst, scale = K.match_eigs_finite(lo, r["values"])
            assert worst <= 100 * n * EPS * scale, f"eigenvalue sets differ: {worst / scale:.3e}"
            # identical quasi-triangular structure: same number of complex pairs / zero eigenvalues
            fin = np.isfinite(lo)
            assert np.count_nonzero(lo[fin].imag > 0) == np.count_nonzero(r["values"][fin].imag > 0)
            assert np.count_nonzero(r["values"] == 0) == np.count_nonzero(lo == 0)
    return T, Z, al, be, sc


@pytest.mark.parametrize("p,S,left", [
    (4, [1, 0, 1, 0], False), (4, [0, 1, 0, 1], True), (1, [1], False), (3, [1, 1, 1], False),
    (5, [1, 0, 1, 1, 0], False), (3, [0, 1, 1], True), (2, [1, 0], False),
])
def test_real_full_n5(psd, p, S, left):
    A = GCs.rand_storage(1234, 5, p, 8, False)
    _compare(psd, A, S, left)


@pytest.mark.parametrize("p", [2, 3, 5])
def test_real_hessut_one_minus(psd, p):
    S = [1, 0] + [1] * (p - 2)
    _compare(psd, GCs.hessut_storage(79, 5, p, 6, False), S, hessut=True)


@pytest.mark.parametrize("S,hole", GCs.HOLE_CASES)
def test_real_hole_cases(psd, S, hole):
    A = GCs.hessut_storage(80 + hole[0] * 10 + hole[1], 5, 5, 6, False, hole=hole)
    T, Z, al, be, sc = _compare(psd, A, S, hessut=True)
    lam = _vals(al[0], be[0], sc[0])
    if S[hole[0] - 1]:
        assert np.count_nonzero(lam == 0) >= 1
    else:
        assert np.count_nonzero(~np.isfinite(lam)) >= 1


@pytest.mark.parametrize("n,p", [(1, 1), (1, 3), (2, 2), (3, 4), (4, 3), (16, 3), (33, 2), (24, 6)])
def test_real_shapes(psd, n, p):
    S = [1] + [(k % 2) for k in range(1, p)]
    A = GCs.rand_storage(7, n, p, 3, False)
    _compare(psd, A, S)
    _compare(psd, A, S[::-1] if S[-1] else [1] * p, left=True)


def test_real_fast_paths(psd):
    for p in (1, 4):
        A = GCs.rand_storage(5, 7, p, 4, False)
        S = [1, 0, 1, 0][:p]
        full = psd.gpschur_batched(A, S)
        fast = psd.gpschur_batched(A, S, wantT=False, wantZ=False)
        assert fast[1] is None
        for b in range(4):
            lf, lq = (_vals(x[2][b], x[3][b], x[4][b]) for x in (full, fast))
            assert K.match_eigs(lf, lq) <= 1e-9 * np.max(np.abs(lf))


# BASELINE configs[4] family: p=10, alternating S with S[p] = true, :L orientation, with Z.
# N = 64 is checked against the oracle; N = 512 (the benchmark size, global-memory path)
# through the size-independent predicates only (the numpy oracle would take minutes).
def test_config5_family(psd):
    S = [k % 2 for k in range(10)]  # [F,T,...,F,T]
    A = GCs.rand_storage(1234, 64, 10, 2, False)
    _compare(psd, A, S, left=True)


def test_config5_full_size(psd):
    S = [k % 2 for k in range(10)]
    A = GCs.rand_storage(1234, 512, 10, 1, False)
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L")
    assert (info == 0).all()
    r = K.gpschur_check(A[0], S, T[0], Z[0], al[0], be[0], sc[0], left=True, real_path=True,
                        tol=200)
    _pairs_ok(r["values"])


def test_real_wrapper(psd):
    A = GCs.rand_storage(3, 6, 3, 1, False)[0]
    mats = [np.ascontiguousarray(A[j].T) for j in range(3)]
    G = psd.gpschur(mats, [True, False, True], "R")
    assert isinstance(G, psd.GeneralizedPeriodicSchur) and G.beta.dtype == np.float64
    ref = np.linalg.eigvals(mats[0] @ np.linalg.inv(mats[1]) @ mats[2])
    assert K.match_eigs(ref, G.values) <= 1e-9 * np.max(np.abs(ref))
    with pytest.raises(ValueError):
        psd.gpschur(mats, [False, True, True], "R")


# committed high-precision fixtures (tests/golden/generalized_golden.json), real and complex cases
from test_oracle_generalized import GGOLD, golden_inputs, golden_tol  # noqa: E402


@pytest.mark.parametrize("ci", range(len(GGOLD["cases"])))
def test_golden_generalized_gpu(psd, ci):
    case = GGOLD["cases"][ci]
    A, ref = golden_inputs(case)
    S = case["S"]
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L" if case["left"] else "R")
    assert info[0] == 0
    lam = _vals(al[0], be[0], sc[0])
    assert K.match_eigs(ref, lam) <= golden_tol(A, S, ref)
    K.gpschur_check(A[0], S, T[0], Z[0], al[0], be[0], sc[0], left=case["left"],
                    real_path=not case["complex"])


# Windowed Stage 2 / windowed double-shift sweeps at the small periods (factors in global memory)
@pytest.mark.parametrize("p,S,left", [(1, [1], False), (2, [1, 0], False), (2, [0, 1], True), (3, [1, 1, 1], True)])
def test_windowed_small_periods_real(psd, p, S, left):
    n = 200
    A = GCs.rand_storage(4343 + p, n, p, 2, False)
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L" if left else "R")
    assert (info == 0).all()
    for b in range(2):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left, real_path=True)
        _pairs_ok(r["values"])
        ref = K.gproduct_eigvals(A[b], S, left)
        assert K.match_eigs(ref, r["values"]) <= 1e-7 * np.max(np.abs(ref))
    _, _, al2, be2, sc2, info2 = psd.gpschur_batched(A, S, "L" if left else "R", wantT=False, wantZ=False)
    assert (info2 == 0).all()
    worst, scale = K.match_eigs_finite(_vals(al[0], be[0], sc[0]), _vals(al2[0], be2[0], sc2[0]))
    assert worst <= 1e-7 * scale
