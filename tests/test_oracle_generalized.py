"""CPU tests pinning the numpy restatement of the complex generalized periodic Schur path
(oracle/gpsd_complex.py) with the reference's own predicates and fixtures
(test/generalized.jl, test/testfuncs.jl:155-235) and eigvals of the explicit product."""
import numpy as np
import pytest

import gpsd_cases as GCs
import psd_checks as K
from oracle import gpsd as OG

EPS = np.finfo(float).eps


def _alt(p):
    S = [True]
    for _ in range(1, p):
        S.append(not S[-1])
    return S


def _check_batch(A, S, out, left=False):
    T, Z, al, be, sc, info = out
    assert (info == 0).all()
    for b in range(A.shape[0]):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left)
        lam = r["values"]
        if all(np.isfinite(lam)):
            ref = K.gproduct_eigvals(A[b], S, left)
            worst = K.match_eigs(ref, lam)
            assert worst <= 1e-9 * np.max(np.abs(ref))


# test/generalized.jl:175-222: full complex, p=5 all-true and p=4 alternating, :R and :L
@pytest.mark.parametrize("p,S,left", [
    (5, [1] * 5, False), (5, [1] * 5, True),
    (4, [1, 0, 1, 0], False), (4, [0, 1, 0, 1], True),
    (1, [1], False), (2, [1, 0], False), (6, [1, 0, 1, 1, 0, 1], False),
])
def test_complex_full(p, S, left):
    A = GCs.rand_storage(1234, 5, p, 3, True)
    _check_batch(A, S, OG.cpschur_batched(A, S, left=left), left)


# test/generalized.jl:224-246 (Hess+UT, all true, with and without hole) and :68-153
@pytest.mark.parametrize("p", [1, 2, 3, 5])
def test_complex_hessut_alltrue(p):
    A = GCs.hessut_storage(77, 5, p, 2, True)
    _check_batch(A, [1] * p, OG.cpschur_batched(A, [1] * p, hessut=True))
    if p > 1:
        A = GCs.hessut_storage(78, 5, p, 2, True, hole=(2, 3))
        _check_batch(A, [1] * p, OG.cpschur_batched(A, [1] * p, hessut=True))


@pytest.mark.parametrize("p", [2, 3, 5])
def test_complex_hessut_one_minus(p):
    S = [1, 0] + [1] * (p - 2)
    A = GCs.hessut_storage(79, 5, p, 2, True)
    _check_batch(A, S, OG.cpschur_batched(A, S, hessut=True))


@pytest.mark.parametrize("S,hole", GCs.HOLE_CASES)
def test_complex_hole_cases(S, hole):
    A = GCs.hessut_storage(80 + hole[0] * 10 + hole[1], 5, 5, 2, True, hole=hole)
    out = OG.cpschur_batched(A, S, hessut=True)
    _check_batch(A, S, out)
    lam = K.gpschur_check(A[0], S, out[0][0], out[1][0], out[2][0], out[3][0], out[4][0])["values"]
    if S[hole[0] - 1]:
        assert np.count_nonzero(lam == 0) >= 1      # zero eigenvalue of the product
    else:
        assert np.count_nonzero(~np.isfinite(lam)) >= 1  # infinite eigenvalue


# test/generalized.jl:154-173: moderate N
@pytest.mark.parametrize("S", [[1, 1, 1, 1], [1, 0, 1, 0]])
def test_complex_moderate_n(S):
    A = GCs.hessut_storage(99, 32, 4, 1, True, hole=(2, 3))
    _check_batch(A, S, OG.cpschur_batched(A, S, hessut=True))


def test_complex_fast_paths():
    """test/generalized.jl:268-303: wantT/wantZ = false give the same eigenvalues."""
    for p in (1, 5):
        A = GCs.rand_storage(5, 5, p, 2, True)
        S = [1] * p
        full = OG.cpschur_batched(A, S)
        fast = OG.cpschur_batched(A, S, wantT=False, wantZ=False)
        for b in range(2):
            lf = full[2][b] / full[3][b] * np.exp2(full[4][b].astype(float))
            lq = fast[2][b] / fast[3][b] * np.exp2(fast[4][b].astype(float))
            assert K.match_eigs(lf, lq) <= 1e-10 * np.max(np.abs(lf))


def test_signature_error():
    A = GCs.rand_storage(1, 4, 3, 1, True)
    with pytest.raises(ValueError):
        OG.cpschur_batched(A, [0, 1, 1])


# ------------------------------------------------------------------------------------------
# real generalized path (oracle/gpsd_real.py): test/generalized.jl:42-153 with T = Float64,
# predicates test/testfuncs.jl:238-382
# ------------------------------------------------------------------------------------------
def _check_real(A, S, out, left=False):
    T, Z, al, be, sc, info = out
    assert (info == 0).all()
    for b in range(A.shape[0]):
        r = K.gpschur_check(A[b], S, T[b], Z[b], al[b], be[b], sc[b], left=left, real_path=True)
        lam = r["values"]
        if all(np.isfinite(lam)):
            ref = K.gproduct_eigvals(A[b], S, left)
            assert K.match_eigs(ref, lam) <= 1e-9 * np.max(np.abs(ref))
        # complex eigenvalues come in adjacent conjugate pairs, positive imaginary part first
        j = 0
        while j < len(lam):
            if np.isfinite(lam[j]) and lam[j].imag != 0:
                assert lam[j].imag > 0 and lam[j + 1] == np.conj(lam[j])
                j += 2
            else:
                j += 1


@pytest.mark.parametrize("p,S,left", [
    (4, [1, 0, 1, 0], False), (4, [0, 1, 0, 1], True), (1, [1], False), (3, [1, 1, 1], False),
    (5, [1, 0, 1, 1, 0], False), (3, [0, 1, 1], True),
])
def test_real_full(p, S, left):
    A = GCs.rand_storage(1234, 5, p, 3, False)
    _check_real(A, S, OG.rgpschur_batched(A, S, left=left), left)


@pytest.mark.parametrize("p", [2, 3, 5])
def test_real_hessut_one_minus(p):
    S = [1, 0] + [1] * (p - 2)
    A = GCs.hessut_storage(79, 5, p, 2, False)
    _check_real(A, S, OG.rgpschur_batched(A, S, hessut=True))


@pytest.mark.parametrize("S,hole", GCs.HOLE_CASES)
def test_real_hole_cases(S, hole):
    A = GCs.hessut_storage(80 + hole[0] * 10 + hole[1], 5, 5, 2, False, hole=hole)
    out = OG.rgpschur_batched(A, S, hessut=True)
    _check_real(A, S, out)
    lam = K.gpschur_check(A[0], S, out[0][0], out[1][0], out[2][0], out[3][0], out[4][0],
                          real_path=True)["values"]
    if S[hole[0] - 1]:
        assert np.count_nonzero(lam == 0) >= 1
    else:
        assert np.count_nonzero(~np.isfinite(lam)) >= 1


def test_real_larger():
    A = GCs.rand_storage(5, 16, 3, 1, False)
    _check_real(A, [1, 0, 1], OG.rgpschur_batched(A, [1, 0, 1]))
    A = GCs.rand_storage(6, 12, 4, 1, False)
    _check_real(A, [0, 1, 0, 1], OG.rgpschur_batched(A, [0, 1, 0, 1], left=True), True)


# ---- committed high-precision fixtures (tests/golden/generalized_golden.json) -----------------
import json  # noqa: E402
import os  # noqa: E402

GGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "generalized_golden.json")))


def golden_inputs(case):
    A = GCs.rand_storage(GGOLD["seed"], case["n"], case["p"], 2, case["complex"])
    ref = np.array([complex(x, y) for x, y in case["eig"]])
    return A[case["b"]:case["b"] + 1], ref


def golden_tol(A, S, ref):
    """first-order bound: eps * (product of the factor condition numbers) * |lambda|max, with a
    floor of 1e-10 relative"""
    k = 1.0
    for j in range(A.shape[1]):
        k *= np.linalg.cond(K.M(A[0, j])) if not S[j] else 1.0
    return max(1e-10, 1e3 * EPS * k) * np.max(np.abs(ref))


@pytest.mark.parametrize("ci", range(len(GGOLD["cases"])))
def test_golden_generalized_oracle(ci):
    case = GGOLD["cases"][ci]
    A, ref = golden_inputs(case)
    S = case["S"]
    f = OG.cpschur_batched if case["complex"] else OG.rgpschur_batched
    T, Z, al, be, sc, info = f(A, S, left=case["left"])
    assert info[0] == 0
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        lam = al[0] / be[0].astype(np.complex128) * np.exp2(sc[0].astype(float))
    assert K.match_eigs(ref, lam) <= golden_tol(A, S, ref)


# ---- the reference's known-answer family through the generalized oracles -----------------------
# test/runtests.jl:68-87 runs expsplit(p, T) for T = ComplexF64 as well (complex standard method
# = complex periodic QZ with S = trues); the real periodic QZ oracle is held to the same gates.
RGOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "real_golden.json")))


def _expsplit_storage(es, left, dtype):
    p = es["p"]
    A1 = np.array(es["A1"], dtype=dtype)
    Aj = np.diag(es["Aj_diag"]).astype(dtype)
    mats = [A1] + [Aj.copy() for _ in range(p - 1)]
    if left:
        mats[0], mats[-1] = mats[-1], mats[0]
    return np.stack([m.T for m in mats])[None].copy()


def _expsplit_gates(es, lam, rel=1e-10):
    for lr, li in es["lambda_reference_asymptotic"]:
        lj = complex(lr, li)
        d = np.abs(lam - lj)
        k = int(np.argmin(d))
        assert d[k] < 1e-3 * abs(lj) or max(abs(lj), abs(lam[k])) < EPS ** 2, (lj, lam)
    for a, b in es["lambda_mp"]:
        g = complex(a, b if abs(b) > 1e-30 * abs(complex(a, b)) else 0.0)
        d = np.abs(lam - g)
        k = int(np.argmin(d))
        assert d[k] <= rel * abs(g) or max(abs(g), abs(lam[k])) < EPS ** 2, (g, lam)


@pytest.mark.parametrize("es", RGOLD["expsplit"], ids=lambda e: f"p{e['p']}")
@pytest.mark.parametrize("left", [False, True], ids=["R", "L"])
@pytest.mark.parametrize("cplx", [True, False], ids=["complex", "realqz"])
def test_expsplit_generalized_oracles(es, left, cplx):
    p = es["p"]
    A = _expsplit_storage(es, left, np.complex128 if cplx else np.float64)
    S = [True] * p
    f = OG.cpschur_batched if cplx else OG.rgpschur_batched
    T, Z, al, be, sc, info = f(A, S, left=left)
    assert info[0] == 0
    r = K.gpschur_check(A[0], S, T[0], Z[0], al[0], be[0], sc[0], left=left, tol=128, real_path=not cplx)
    _expsplit_gates(es, r["values"])


# ---- graded factors: relative accuracy of every eigenvalue (rpschur2x2.jl, _qzrots) -------------
import graded_cases as GR  # noqa: E402


@pytest.mark.parametrize("case", GR.GRADED["cases"], ids=GR.case_id)
def test_graded_relative_accuracy_oracle(case):
    A, ref = GR.inputs(case)
    T, Z, al, be, sc, info = OG.rgpschur_batched(A, case["S"])
    assert info[0] == 0
    with np.errstate(all="ignore"):
        lam = al[0] / be[0] * np.exp2(sc[0].astype(float))
    K.gpschur_check(A[0], case["S"], T[0], Z[0], al[0], be[0], sc[0], real_path=True,
                    tol=100 * max(1.0, np.abs(K.M(A[0])).max()), baseline_gates=False)
    # one documented outlier of the family (n6 p3 k24 b1: 5e-7, an ill-conditioned cluster)
    gate = 1e-5 if GR.case_id(case) == "n6p3k24b1" else 1e-8
    assert GR.worst_relative_error(ref, lam) <= gate
