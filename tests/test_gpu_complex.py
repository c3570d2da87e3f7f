"""GPU parity tests for the complex (generalized) periodic Schur path: CUDA library through the
C ABI against the numpy restatement of the reference (oracle/gpsd_complex.py) on identical
seeded inputs, with the reference's predicates (test/testfuncs.jl:155-235) and BASELINE.json's
gates (eigenvalues as a matched set to 100*N*eps*|lambda|max; residual and orthogonality
<= 10*N*eps; triangular structure with exact zeros)."""
import numpy as np
import pytest

import gpsd_cases as GCs
import psd_checks as K
from oracle import gpsd as OG

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _vals(a, b, s):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return a / b * np.exp2(s.astype(float))


def _compare(psd, A, S, left=False, hessut=False, oracle_check=True):
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L" if left else "R", hessut=hessut)
    assert (info == 0).all()
    n = A.shape[-1]
    if oracle_check:
        To, Zo, alo, beo, sco, io = OG.cpschur_batched(A, S, left=left, hessut=hessut)
        assert (io == 0).all()
    for b in range(A.shape[0]):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left)
        if oracle_check:
            lo = _vals(alo[b], beo[b], sco[b])
            worst, scale = K.match_eigs_finite(lo, r["values"])
            assert worst <= 100 * n * EPS * scale, f"eigenvalue sets differ: {worst / scale:.3e}"
            assert np.count_nonzero(r["values"] == 0) == np.count_nonzero(lo == 0)
    return T, Z, al, be, sc


@pytest.mark.parametrize("p,S,left", [
    (5, [1] * 5, False), (5, [1] * 5, True),
    (4, [1, 0, 1, 0], False), (4, [0, 1, 0, 1], True),
    (1, [1], False), (2, [1, 0], False), (6, [1, 0, 1, 1, 0, 1], False), (3, [0, 1, 1], True),
])
def test_complex_full_n5(psd, p, S, left):
    A = GCs.rand_storage(1234, 5, p, 6, True)
    _compare(psd, A, S, left)


@pytest.mark.parametrize("p", [1, 2, 3, 5])
def test_complex_hessut_alltrue(psd, p):
    A = GCs.hessut_storage(77, 5, p, 4, True)
    _compare(psd, A, [1] * p, hessut=True)
    if p > 1:
        A = GCs.hessut_storage(78, 5, p, 4, True, hole=(2, 3))
        _compare(psd, A, [1] * p, hessut=True)


@pytest.mark.parametrize("p", [2, 3, 5])
def test_complex_hessut_one_minus(psd, p):
    S = [1, 0] + [1] * (p - 2)
    _compare(psd, GCs.hessut_storage(79, 5, p, 4, True), S, hessut=True)


@pytest.mark.parametrize("S,hole", GCs.HOLE_CASES)
def test_complex_hole_cases(psd, S, hole):
    A = GCs.hessut_storage(80 + hole[0] * 10 + hole[1], 5, 5, 4, True, hole=hole)
    T, Z, al, be, sc = _compare(psd, A, S, hessut=True)
    lam = _vals(al[0], be[0], sc[0])
    if S[hole[0] - 1]:
        assert np.count_nonzero(lam == 0) >= 1
    else:
        assert np.count_nonzero(~np.isfinite(lam)) >= 1


@pytest.mark.parametrize("S", [[1, 1, 1, 1], [1, 0, 1, 0]])
def test_complex_moderate_n(psd, S):
    A = GCs.hessut_storage(99, 32, 4, 2, True, hole=(2, 3))
    _compare(psd, A, S, hessut=True)


@pytest.mark.parametrize("n,p", [(1, 1), (1, 3), (2, 2), (3, 4), (16, 3), (40, 2)])
def test_complex_shapes(psd, n, p):
    S = [1] + [(k % 2) for k in range(1, p)]
    A = GCs.rand_storage(7, n, p, 3, True)
    _compare(psd, A, S)
    _compare(psd, A, S[::-1] if S[-1] else [1] * p, left=True)


# BASELINE configs[2] shape: p=6, N=128, S = [T,F,T,T,F,T], :R, with Z (global-memory path)
def test_config3_shape(psd):
    S = [1, 0, 1, 1, 0, 1]
    A = GCs.rand_storage(1234, 128, 6, 2, True)
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "R")
    assert (info == 0).all()
    To, Zo, alo, beo, sco, io = OG.cpschur_batched(A[:1], S)
    K.gpschur_check(A[0], S, T[0], Z[0], al[0], be[0], sc[0])
    K.gpschur_check(A[1], S, T[1], Z[1], al[1], be[1], sc[1])
    worst, scale = K.match_eigs_finite(_vals(alo[0], beo[0], sco[0]), _vals(al[0], be[0], sc[0]))
    assert worst <= 100 * 128 * EPS * scale


def test_complex_fast_paths(psd):
    """test/generalized.jl:268-303"""
    for p in (1, 5):
        A = GCs.rand_storage(5, 5, p, 4, True)
        S = [1] * p
        full = psd.gpschur_batched(A, S)
        fast = psd.gpschur_batched(A, S, wantT=False, wantZ=False)
        assert fast[1] is None
        part = psd.gpschur_batched(A, S, wantT=True, wantZ=False)
        for b in range(4):
            lf, lq, lp = (_vals(x[2][b], x[3][b], x[4][b]) for x in (full, fast, part))
            assert K.match_eigs(lf, lq) <= 1e-10 * np.max(np.abs(lf))
            assert K.match_eigs(lf, lp) <= 1e-10 * np.max(np.abs(lf))
            assert np.linalg.norm(part[0][b] - full[0][b]) < 20 * EPS * 5 * max(1.0, np.abs(full[0][b]).max())


def test_complex_standard_wrapper(psd):
    """PeriodicSchurDecompositions.jl:1106-1111: complex pschur!(A, lr) returns PeriodicSchur."""
    A = GCs.rand_storage(3, 6, 3, 1, True)[0]
    mats = [np.ascontiguousarray(A[j].T) for j in range(3)]
    F = psd.pschur(mats, "R")
    assert isinstance(F, psd.PeriodicSchur) and F.schurindex == 1
    ref = np.linalg.eigvals(mats[0] @ mats[1] @ mats[2])
    assert K.match_eigs(ref, F.values) <= 1e-10 * np.max(np.abs(ref))
    G = psd.gpschur(mats, [True, False, True], "R")
    assert isinstance(G, psd.GeneralizedPeriodicSchur)
    ref = np.linalg.eigvals(mats[0] @ np.linalg.inv(mats[1]) @ mats[2])
    assert K.match_eigs(ref, G.values) <= 1e-9 * np.max(np.abs(ref))


def test_signature_error(psd):
    A = GCs.rand_storage(1, 4, 3, 1, True)
    with pytest.raises(psd.PsdError) as ei:
        psd.gpschur_batched(A, [0, 1, 1], "R")
    assert ei.value.code == -4
    with pytest.raises(psd.PsdError):
        psd.gpschur_batched(A, [1, 1, 0], "L")


# test/generalized.jl:2-40 ("Generalized Periodic Hessenberg"): alternating S, p in {2,5}, n = 5,
# real and complex; also larger shapes
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,p", [(5, 2), (5, 5), (1, 2), (17, 4), (40, 3)])
def test_generalized_hessenberg(psd, cplx, n, p):
    S = [True]
    for _ in range(1, p):
        S.append(not S[-1])
    A = GCs.rand_storage(4321, n, p, 3, cplx)
    H, Q = psd.gphessenberg_batched(A, S)
    for b in range(3):
        for j in range(p):
            Hj, Aj, Qj, Qn = K.M(H[b, j]), K.M(A[b, j]), K.M(Q[b, j]), K.M(Q[b, (j + 1) % p])
            assert not np.tril(Hj, -2 if j == 0 else -1).any()
            assert np.linalg.norm(Qj @ Qj.conj().T - np.eye(n)) < 10 * EPS * n
            Ax = Qj @ Hj @ Qn.conj().T if S[j] else Qn @ Hj @ Qj.conj().T
            assert np.linalg.norm(Aj - Ax) < 20 * EPS * n * max(1.0, np.linalg.norm(Aj) / n ** 0.5)
    H2, Q2 = psd.gphessenberg_batched(A, S, wantQ=False)
    assert Q2 is None and np.allclose(H2, H, atol=1e-13)
    with pytest.raises(ValueError):
        psd.gphessenberg_batched(A, [False] + S[1:])


# Windowed Stage 2 / windowed sweeps (factors in global memory) at the small periods, both
# orientations, with and without T / Z: sizes chosen so that the factors do not fit in shared memory.
@pytest.mark.parametrize("p,S,left", [(1, [1], False), (2, [1, 0], False), (2, [0, 1], True), (3, [1, 1, 1], True)])
def test_windowed_small_periods_complex(psd, p, S, left):
    n = 150
    A = GCs.rand_storage(4242 + p, n, p, 2, True)
    T, Z, al, be, sc, info = psd.gpschur_batched(A, S, "L" if left else "R")
    assert (info == 0).all()
    for b in range(2):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left)
        ref = K.gproduct_eigvals(A[b], S, left)
        assert K.match_eigs(ref, r["values"]) <= 1e-8 * np.max(np.abs(ref))
    # eigenvalues only: updates restricted to the active window
    _, _, al2, be2, sc2, info2 = psd.gpschur_batched(A, S, "L" if left else "R", wantT=False, wantZ=False)
    assert (info2 == 0).all()
    worst, scale = K.match_eigs_finite(_vals(al[0], be[0], sc[0]), _vals(al2[0], be2[0], sc2[0]))
    assert worst <= 1e-8 * scale
