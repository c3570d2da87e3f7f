"""GPU parity tests for the real standard periodic Schur path: CUDA library (through the C ABI)
against the CPU oracle on identical seeded inputs, using the reference's own acceptance
predicates (test/testfuncs.jl:56-145) and BASELINE.json's gates:
  eigenvalues matched as a set, rel err <= 100*N*eps*|lambda|max;
  residual ||Q'AQ - T||/||A|| <= 10*N*eps;  ||Q'Q - I|| <= 10*N*eps;  identical structure.
"""
import numpy as np
import pytest

import psd_checks as K

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps


def _eig_gate(lam_oracle, lam_gpu, n):
    scale = np.max(np.abs(lam_oracle))
    worst = K.match_eigs(lam_oracle, lam_gpu)
    assert worst <= 100 * n * EPS * scale, f"eigenvalue sets differ: {worst / scale:.3e} rel"
    # same number of complex pairs (quasi-triangular structure identical)
    assert np.count_nonzero(lam_oracle.imag > 0) == np.count_nonzero(lam_gpu.imag > 0)


def _run_case(psd, oracle, n, p, batch, left, seed=1234, tol=32):
    A = oracle.gen_real(seed, n, p, batch)
    To, Zo, lo, io, _ = oracle.rpschur_batched(A, left=left)
    T, Z, lam, info = psd.pschur_batched(A, "L" if left else "R")
    assert (info == 0).all() and (io == 0).all()
    for b in range(batch):
        K.pschur_check(A[b], T[b], Z[b], lam[b], left=left, tol=tol)
        _eig_gate(lo[b], lam[b], n)
    return A, T, Z, lam


# reference shapes: test/runtests.jl:89-100 ("Periodic Schur full"), n=5, p in 1,2,3,5, :R and :L
@pytest.mark.parametrize("p", [1, 2, 3, 5])
@pytest.mark.parametrize("left", [False, True])
def test_full_n5(psd, oracle, p, left):
    _run_case(psd, oracle, 5, p, 16, left)


@pytest.mark.parametrize("n,p", [(1, 1), (1, 4), (2, 1), (3, 2), (3, 5), (4, 3), (7, 12), (16, 6)])
def test_small_shapes(psd, oracle, n, p):
    _run_case(psd, oracle, n, p, 8, False)
    _run_case(psd, oracle, n, p, 8, True)


# BASELINE configs[0]: real p=3 N=50 :R with Schur vectors
def test_config1_p3_n50(psd, oracle):
    _run_case(psd, oracle, 50, 3, 4, False)
    _run_case(psd, oracle, 50, 3, 2, True)


# BASELINE configs[1] shape (p=8, N=32), with T and Z on a sample ...
def test_config2_shape_with_vectors(psd, oracle):
    _run_case(psd, oracle, 32, 8, 8, False)


# ... and eigenvalues only (the benchmarked mode) on a larger sample, against the oracle
def test_config2_eigs_only(psd, oracle):
    n, p, batch = 32, 8, 512
    A = oracle.gen_real(1234, n, p, batch)
    _, _, lo, io, _ = oracle.rpschur_batched(A, wantT=False, wantZ=False)
    _, Z, lam, info = psd.pschur_batched(A, "R", wantZ=False, wantT=False)
    assert Z is None
    assert (info == 0).all() and (io == 0).all()
    for b in range(batch):
        _eig_gate(lo[b], lam[b], n)
    # size-independent property: sum of eigenvalues == trace of the product
    for b in range(0, batch, 37):
        P = K.M(A[b, 0]).copy()
        for j in range(1, p):
            P = P @ K.M(A[b, j])
        tr = np.trace(P)
        assert abs(lam[b].sum() - tr) <= 1e-9 * abs(tr)


# reference "fast paths" test/runtests.jl:103-132
@pytest.mark.parametrize("p", [1, 5])
def test_fast_paths(psd, oracle, p):
    n = 5
    A = oracle.gen_real(7, n, p, 8)
    T2, Z2, lam2, info2 = psd.pschur_batched(A, "R", wantZ=True, wantT=True)
    T0, Z0, lam0, info0 = psd.pschur_batched(A, "R", wantZ=False, wantT=False)
    T1, Z1, lam1, info1 = psd.pschur_batched(A, "R", wantZ=False, wantT=True)
    assert Z0 is None and Z1 is None
    for b in range(A.shape[0]):
        K.compare_reigvals(lam2[b], lam0[b], 1000 * EPS)
        K.compare_reigvals(lam2[b], lam1[b], 1000 * EPS)
        assert np.linalg.norm(T1[b, 0] - T2[b, 0]) < 20 * EPS * n * max(1.0, np.linalg.norm(T2[b, 0]))


# global-memory (L2-resident) variant of the same kernel: factors do not fit in shared memory
def test_global_mode_n64_p8(psd, oracle):
    _run_case(psd, oracle, 64, 8, 2, False)
    st = psd.default_handle().stats()
    assert st["problems_global"] == 2 and st["problems_smem"] == 0


def test_hessut_entry(psd, oracle):
    # test/runtests.jl:53-66: Hessenberg + upper-triangular input
    n = 5
    for p in (1, 2, 3, 5):
        A = oracle.gen_real(11, n, p, 4)
        for b in range(4):
            for j in range(p):
                Mj = np.triu(K.M(A[b, j]), -1 if j == 0 else 0)
                A[b, j] = Mj.T
        T, Z, lam, info = psd.pschur_hessut_batched(A)
        assert (info == 0).all()
        for b in range(4):
            K.pschur_check(A[b], T[b], Z[b], lam[b], left=False)
            To, Zo, lo, io = oracle.rpschur_hessut(A[b])
            _eig_gate(lo, lam[b], n)


def test_hessut_entry_q_preset(psd, oracle):
    """the `Q` keyword of the inner method (PeriodicSchurDecompositions.jl:326; krylov.jl:583-591):
    Schur vectors accumulated onto the caller's orthogonal Q_j.  Check: the decomposition holds for
    B_j = Q_j H_j Q_{j+1}' with the returned vectors, and they equal Q_j Z_j of the plain call."""
    rng = np.random.default_rng(5)
    for (n, p) in [(5, 1), (7, 3), (33, 2), (50, 3)]:
        A = oracle.gen_real(21, n, p, 3)
        for b in range(3):
            for j in range(p):
                A[b, j] = np.triu(K.M(A[b, j]), -1 if j == 0 else 0).T
        Q = np.empty_like(A)
        for b in range(3):
            for j in range(p):
                Q[b, j] = np.linalg.qr(rng.standard_normal((n, n)))[0].T  # storage = transpose of the matrix
        T0, Z0, lam0, info0 = psd.pschur_hessut_batched(A)
        T, Z, lam, info = psd.pschur_hessut_batched(A, Q=Q)
        assert (info == 0).all() and (info0 == 0).all()
        for b in range(3):
            B = np.empty_like(A[b])
            for j in range(p):
                B[j] = (K.M(Q[b, j]) @ K.M(A[b, j]) @ K.M(Q[b, (j + 1) % p]).T).T
            K.pschur_check(B, T[b], Z[b], lam[b], left=False, check_lambda=False)
            for j in range(p):
                assert np.allclose(K.M(Z[b, j]), K.M(Q[b, j]) @ K.M(Z0[b, j]), atol=200 * n * EPS)
            assert np.array_equal(T[b], T0[b]) and np.array_equal(lam[b], lam0[b])


def test_iteration_counts(psd, oracle):
    """psd_set_iters_output: QR iterations per problem (the reference's niter, :458-459, 1077),
    from the one-CTA kernel (with T / Z) and from the warp-per-problem eigenvalue kernel with its
    occupancy phases; both run the same iteration, so the counts agree problem by problem with the
    CPU restatement wherever that converges the same way (a sanity band is asserted, not equality:
    the eigenvalue kernel's reflectors round differently)."""
    for (n, p, batch) in [(5, 3, 8), (32, 8, 40), (20, 4, 16)]:
        A = oracle.gen_real(77, n, p, batch)
        T, Z, lam, info, it_full = psd.pschur_batched(A, "R", return_iters=True)
        _, _, lam2, info2, it_eig = psd.pschur_batched(A, "R", wantT=False, wantZ=False, return_iters=True)
        assert (info == 0).all() and (info2 == 0).all()
        assert (it_full > 0).all() and (it_eig > 0).all()
        # at least one iteration per two eigenvalues, at most the reference's budget maxitfac * n
        assert (it_full >= n // 2).all() and (it_full <= 30 * n).all()
        assert (it_eig >= n // 2).all() and (it_eig <= 30 * n).all()
        assert np.abs(it_full.astype(float) - it_eig).max() <= 0.5 * it_full.max()
        # switched off again: a later call must not touch the old buffer
        keep = it_full.copy()
        psd.pschur_batched(A, "R")
        assert np.array_equal(keep, it_full)


def test_device_checkpsd(psd, oracle):
    """psd_rcheckpsd_batched against the same norms formed with numpy (checkpsd,
    diagnostics.jl:190-263): good decompositions pass with the reference's thresholds, a perturbed T,
    a perturbed Z and a non-zero entry below the triangle are each reported in the right place."""
    for (n, p, batch, lr) in [(5, 3, 6, "R"), (33, 2, 3, "L"), (50, 3, 4, "R"), (200, 2, 1, "R"), (1100, 1, 1, "R")]:  # the last one: one right-hand side per pass
        A = oracle.gen_real(91, n, p, batch)
        T, Z, lam, info = psd.pschur_batched(A, lr)
        ok, err, tri, orth = psd.checkpsd_batched(A, T, Z, lr)
        assert ok.all(), (err.max(), tri.max(), orth.max())
        left = lr == "L"
        for b in range(batch):
            for j in range(p):
                Tj, Aj, Zj, Zn = K.M(T[b, j]), K.M(A[b, j]), K.M(Z[b, j]), K.M(Z[b, (j + 1) % p])
                Ax = (Zn @ Tj @ Zj.T) if left else (Zj @ Tj @ Zn.T)
                ref = np.linalg.norm(Ax - Aj) / EPS / np.linalg.norm(Aj, 1)
                assert abs(err[b, j] - ref) <= 0.05 * ref + 0.5, (err[b, j], ref)
                ro = np.linalg.norm(Zj @ Zj.T - np.eye(n))
                assert abs(orth[b, j] - ro) <= 0.05 * ro + 2 * EPS
        assert (tri == 0).all()
        # faults
        T2 = T.copy(); T2[0, p - 1, 0, 0] += 1e-6          # entry (0, 0) of the last factor
        ok2, err2, _, _ = psd.checkpsd_batched(A, T2, Z, lr)
        assert not ok2[0] and err2[0, p - 1] > 1e6 and ok2[1:].all()
        Z2 = Z.copy(); Z2[batch - 1, 0, 1, 0] += 1e-8
        ok3, _, _, orth3 = psd.checkpsd_batched(A, T, Z2, lr)
        assert not ok3[batch - 1] and orth3[batch - 1, 0] > 1e-9
        T3 = T.copy(); T3[0, 0, 0, n - 1] = 1e-30           # storage [col][row]: row n-1, column 0
        ok4, _, tri4, _ = psd.checkpsd_batched(A, T3, Z, lr)
        assert not ok4[0] and tri4[0, 0] > 0
        ok5, _, _, _ = psd.checkpsd_batched(A, T3, Z, lr, strict=False)
        assert ok5[0]


def test_reduction_packed_output(psd, oracle):
    """psd_rphess_packed_batched: (A[j], tau[j]) as phessenberg! leaves them
    (PeriodicSchurDecompositions.jl:229-253).  Q_j is rebuilt from the packed reflectors the way
    LAPACK's orghr / orgqr (and Julia's Hessenberg / QR Q factors) do, and must reproduce the
    explicit reduction: Q_j' A_j Q_{j+1} = H_j with the same H_j as psd_rphess_batched."""
    for (n, p, batch) in [(5, 1, 3), (6, 3, 4), (33, 4, 2), (70, 2, 1), (200, 3, 1)]:  # the last one works in global memory
        A = oracle.gen_real(61, n, p, batch)
        F, tau = psd.phessenberg_packed_batched(A)
        H, Q = psd.phessenberg_batched(A)
        for b in range(batch):
            Qs = []
            for j in range(p):
                Fj = K.M(F[b, j])
                first = 1 if j == 0 else 0           # H_1: reflector i acts on rows i+1.., others on rows i..
                Qj = np.eye(n)
                for i in range(n - 1):
                    r0 = i + first
                    if r0 >= n:
                        continue
                    v = np.zeros(n)
                    v[r0] = 1.0
                    v[r0 + 1:] = Fj[r0 + 1:, i]
                    Qj = Qj @ (np.eye(n) - tau[b, j, i] * np.outer(v, v))
                Qs.append(Qj)
                assert np.linalg.norm(Qj @ Qj.T - np.eye(n)) < 50 * n * EPS
            for j in range(p):
                Hj = np.triu(K.M(F[b, j]), -1 if j == 0 else 0)
                Aj = K.M(A[b, j])
                res = np.linalg.norm(Qs[j].T @ Aj @ Qs[(j + 1) % p] - Hj)
                assert res < 50 * n * EPS * max(1.0, np.linalg.norm(Aj)), (n, p, j, res)
                assert np.allclose(Hj, K.M(H[b, j]), atol=200 * n * EPS * max(1.0, np.abs(Hj).max()))


def test_reduction_only(psd, oracle):
    # test/runtests.jl:14-50 "Periodic Hessenberg"
    for (n, p) in [(5, 1), (5, 2), (5, 5), (32, 8)]:
        A = oracle.gen_real(5, n, p, 3)
        H, Q = psd.phessenberg_batched(A)
        for b in range(3):
            for j in range(p):
                Hj, Qj, Qn, Aj = K.M(H[b, j]), K.M(Q[b, j]), K.M(Q[b, (j + 1) % p]), K.M(A[b, j])
                assert not np.tril(Hj, -2 if j == 0 else -1).any()
                assert np.linalg.norm(Qj @ Qj.T - np.eye(n)) < 10 * EPS * n
                assert np.linalg.norm(Aj - Qj @ Hj @ Qn.T) < 20 * EPS * n * max(1.0, np.linalg.norm(Aj, 1))


def test_empty_batch_and_errors(psd):
    A = np.zeros((0, 3, 4, 4))
    T, Z, lam, info = psd.pschur_batched(A)
    assert lam.shape == (0, 4)
    with pytest.raises(ValueError):
        psd.pschur_batched(np.zeros((1, 2, 3, 3)), "X")
    with pytest.raises(ValueError):
        psd.pschur_batched(np.zeros((1, 2, 3, 4)))


def test_struct_api(psd, oracle):
    # pschur / pschur_ mirror of the reference API incl. aliasing of T1
    n, p = 6, 3
    S = oracle.gen_real(3, n, p, 1)[0]
    A = [np.ascontiguousarray(K.M(S[j])) for j in range(p)]
    A0 = [a.copy() for a in A]
    F = psd.pschur_(A, "R")
    assert F.schurindex == 1 and F.orientation == "R" and F.period == p
    assert F.T1 is A[0]
    for j in range(p):
        Tj = F.T1 if j == 0 else F.T[j - 1]
        R = F.Z[j] @ Tj @ F.Z[(j + 1) % p].T - A0[j]
        assert np.linalg.norm(R) < 32 * EPS * np.linalg.norm(A0[j], 1)
    G = psd.pschur(A0, "L")
    assert G.schurindex == p and G.orientation == "L"
    for j in range(p):
        Tj = G.T1 if j == p - 1 else G.T[j]
        R = G.Z[(j + 1) % p] @ Tj @ G.Z[j].T - A0[j]
        assert np.linalg.norm(R) < 32 * EPS * np.linalg.norm(A0[j], 1)
