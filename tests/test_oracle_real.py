"""CPU tests pinning the oracle (oracle/psdo_real.hpp) for the real standard path against
 - the committed golden vectors (tests/golden/real_golden.json: high-precision eigenvalues of
   the explicit product for seeded inputs, and the reference's expsplit known answers,
   /root/reference test/testfuncs.jl:412-421, test/runtests.jl:68-87),
 - the reference's own acceptance predicates on the reference's own shapes
   (test/runtests.jl:14-132 via test/testfuncs.jl:56-145)."""
import json
import os

import numpy as np
import pytest

import psd_checks as K
import psd_rng

EPS = np.finfo(np.float64).eps
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "real_golden.json")))


def _c(v):
    # 60-digit arithmetic leaves ~1e-60 relative dust in the imaginary parts of real eigenvalues
    return np.array([complex(a, b if abs(b) > 1e-30 * abs(complex(a, b)) else 0.0) for a, b in v])


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"n{c['n']}p{c['p']}b{c['b']}{'L' if c['left'] else 'R'}")
def test_golden_eigenvalues(oracle, case):
    n, p, b, left = case["n"], case["p"], case["b"], case["left"]
    A = psd_rng.gen_uniform(GOLD["seed"], n, p, b + 1)[b:b + 1]
    T, Z, lam, info, _ = oracle.rpschur_batched(A, left=left)
    assert info[0] == 0
    gold = _c(case["eig"])
    scale = np.max(np.abs(gold))
    assert K.match_eigs(gold, lam[0]) <= 1000 * EPS * scale
    K.pschur_check(A[0], T[0], Z[0], lam[0], left=left, lam_ref=gold)
    # eigenvalue-only fast path agrees (test/runtests.jl:103-132)
    _, _, lam0, info0, _ = oracle.rpschur_batched(A, left=left, wantT=False, wantZ=False)
    assert info0[0] == 0
    assert K.match_eigs(gold, lam0[0]) <= 1000 * EPS * scale


@pytest.mark.parametrize("es", GOLD["expsplit"], ids=lambda e: f"p{e['p']}")
@pytest.mark.parametrize("left", [False, True])
def test_expsplit_known_answer(oracle, es, left):
    p = es["p"]
    A1 = np.array(es["A1"], dtype=np.float64)
    Aj = np.diag(es["Aj_diag"])
    mats = [A1] + [Aj.copy() for _ in range(p - 1)]
    if left:
        mats[0], mats[-1] = mats[-1], mats[0]  # test/runtests.jl:80
    A = np.stack([m.T for m in mats])[None].copy()
    T, Z, lam, info, _ = oracle.rpschur_batched(A, left=left)
    assert info[0] == 0
    K.pschur_check(A[0], T[0], Z[0], lam[0], left=left, tol=128, check_lambda=False)
    if not left:
        # reference's assertion: within 1e-3 of the asymptotic values (or both < eps^2)
        for lr, li in es["lambda_reference_asymptotic"]:
            lj = complex(lr, li)
            d = np.abs(lam[0] - lj)
            k = int(np.argmin(d))
            assert d[k] < 1e-3 * abs(lj) or max(abs(lj), abs(lam[0][k])) < EPS ** 2
        # stronger: relative accuracy of every eigenvalue, including the 1e-57 one, against the
        # high-precision values (what the periodic algorithm buys over forming the product)
        for g in _c(es["lambda_mp"]):
            d = np.abs(lam[0] - g)
            k = int(np.argmin(d))
            assert d[k] <= 1e-10 * abs(g) or max(abs(g), abs(lam[0][k])) < EPS ** 2, (g, lam[0])


@pytest.mark.parametrize("p", [1, 2, 3, 5])
def test_hess_ut_input(oracle, p):
    """test/runtests.jl:53-66: already Hessenberg/triangular input, :R and :L."""
    n = 5
    A = psd_rng.gen_uniform(77, n, p, 6)
    for b in range(6):
        for j in range(p):
            A[b, j] = np.triu(K.M(A[b, j]), -1 if j == 0 else 0).T
    T, Z, lam, info, _ = oracle.rpschur_batched(A)
    assert (info == 0).all()
    for b in range(6):
        K.pschur_check(A[b], T[b], Z[b], lam[b])
        Th, Zh, lh, ih = oracle.rpschur_hessut(A[b])
        assert ih == 0
        K.pschur_check(A[b], Th, Zh, lh)
    if p > 1:
        A2 = A.copy()
        A2[:, 0], A2[:, p - 1] = A[:, p - 1], A[:, 0]
        T, Z, lam, info, _ = oracle.rpschur_batched(A2, left=True)
        for b in range(6):
            K.pschur_check(A2[b], T[b], Z[b], lam[b], left=True)


@pytest.mark.parametrize("p", [1, 2, 5])
def test_periodic_hessenberg(oracle, p):
    """test/runtests.jl:14-50."""
    n = 5
    A = psd_rng.gen_uniform(5, n, p, 4)
    H, Q = oracle.rphess_batched(A)
    for b in range(4):
        for j in range(p):
            Hj, Qj, Qn, Aj = K.M(H[b, j]), K.M(Q[b, j]), K.M(Q[b, (j + 1) % p]), K.M(A[b, j])
            assert not np.tril(Hj, -2 if j == 0 else -1).any()
            assert np.linalg.norm(Qj @ Qj.T - np.eye(n)) < 10 * EPS * n
            assert np.linalg.norm(Aj - Qj @ Hj @ Qn.T) < 20 * EPS * n


def test_moderate_and_config_shapes(oracle):
    for (n, p, nb) in [(32, 8, 6), (50, 3, 3), (64, 4, 2), (7, 12, 4), (1, 3, 2), (2, 1, 4)]:
        for left in (False, True):
            A = psd_rng.gen_uniform(1234, n, p, nb)
            T, Z, lam, info, _ = oracle.rpschur_batched(A, left=left)
            assert (info == 0).all()
            for b in range(nb):
                K.pschur_check(A[b], T[b], Z[b], lam[b], left=left)


def test_nonconvergence_is_reported_per_problem(oracle):
    """maxitfac=1 starves the iteration: info = level (PeriodicSchurDecompositions.jl:891-893)
    for the problems that fail, without disturbing the others."""
    A = psd_rng.gen_uniform(1234, 12, 3, 8)
    _, _, lam, info, _ = oracle.rpschur_batched(A, wantT=False, wantZ=False, maxitfac=1)
    assert (info > 0).any()
    _, _, lam2, info2, _ = oracle.rpschur_batched(A, wantT=False, wantZ=False)
    assert (info2 == 0).all()
