"""GPU tests of the large-N path: blocked periodic Hessenberg-triangular reduction
(panel kernel + FP64 tensor-core GEMM updates, csrc/psd_large_hess.cuh) against the reference's
"Periodic Hessenberg" predicates (test/runtests.jl:14-50) and the CPU oracle, and the full
pschur! on top of it."""
import numpy as np
import pytest

import psd_checks as K

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _check_hess(A, H, Q):
    p, n, _ = A.shape
    for j in range(p):
        Hj, Aj, Qj, Qn = K.M(H[j]), K.M(A[j]), K.M(Q[j]), K.M(Q[(j + 1) % p])
        assert not np.tril(Hj, -2 if j == 0 else -1).any(), f"factor {j} structure"
        assert np.linalg.norm(Qj @ Qj.T - np.eye(n)) < 10 * EPS * n
        res = np.linalg.norm(Aj - Qj @ Hj @ Qn.T)
        assert res < 20 * EPS * n * max(1.0, np.linalg.norm(Aj) / n ** 0.5), (j, res / (EPS * n))


@pytest.mark.parametrize("n,p", [(192, 1), (200, 3), (256, 4), (333, 2)])
def test_large_reduction(psd, oracle, n, p):
    A = oracle.gen_real(1234, n, p, 1)
    H, Q = psd.phessenberg_batched(A)
    _check_hess(A[0], H[0], Q[0])
    # same factors as the unblocked oracle up to the signs of rows/columns
    Ho, _ = oracle.rphess_batched(A)
    for j in range(p):
        assert np.allclose(np.abs(H[0, j]), np.abs(Ho[0, j]), atol=1e-9 * max(1.0, np.abs(Ho[0, j]).max()))
    # without Q
    H2, Q2 = psd.phessenberg_batched(A, wantQ=False)
    assert Q2 is None and np.allclose(H2, H, atol=1e-12 * np.abs(H).max())


def test_large_pschur(psd, oracle):
    """full decomposition at N = 256 (blocked reduction, then the periodic QR iteration with Z
    preset to the accumulated Q)"""
    n, p = 256, 3
    A = oracle.gen_real(77, n, p, 1)
    for lr in ("R", "L"):
        T, Z, lam, info = psd.pschur_batched(A, lr)
        assert info[0] == 0
        K.pschur_check(A[0], T[0], Z[0], lam[0], left=(lr == "L"), tol=64, check_lambda=False)
        tr = np.trace(np.linalg.multi_dot([K.M(A[0, j]) for j in (range(p) if lr == "R" else range(p - 1, -1, -1))])
                      ) if p > 1 else np.trace(K.M(A[0, 0]))
        assert abs(lam[0].sum() - tr) <= 1e-8 * abs(tr)


def test_large_fast_paths_and_batch(psd, oracle):
    """team-mode iteration: eigenvalues-only and T-only variants agree with the full run
    (test/runtests.jl:103-132 at a large-N shape), and a small batch of large problems is solved
    one after the other by the team"""
    n, p = 200, 2
    A = oracle.gen_real(31, n, p, 3)
    T, Z, lam, info = psd.pschur_batched(A, "R")
    assert (info == 0).all()
    _, Z0, lam0, info0 = psd.pschur_batched(A, "R", wantT=False, wantZ=False)
    T1, Z1, lam1, info1 = psd.pschur_batched(A, "L", wantT=True, wantZ=False)
    assert Z0 is None and Z1 is None and (info0 == 0).all() and (info1 == 0).all()
    for b in range(3):
        K.pschur_check(A[b], T[b], Z[b], lam[b], tol=64, check_lambda=False)
        scale = np.max(np.abs(lam[b]))
        assert K.match_eigs(lam[b], lam0[b]) <= 1000 * n * EPS * scale
        ref = np.linalg.eigvals(K.M(A[b, 0]) @ K.M(A[b, 1]))
        assert K.match_eigs(ref, lam[b]) <= 1e-8 * scale
        refl = np.linalg.eigvals(K.M(A[b, 1]) @ K.M(A[b, 0]))
        assert K.match_eigs(refl, lam1[b]) <= 1e-8 * scale


# Residuals of the CPU restatement of the reference algorithm (oracle.rpschur_batched) on the
# inputs below, in units of eps * ||A_j||_1 (worst factor), recorded with
#   python -c "from oracle import oracle as O; ..."  (74 s for N = 1024 on one core; see
#   profiles/r2_oracle_large_residuals.txt).  The reference's own threshold tol = 32
# (test/testfuncs.jl:119) is calibrated on its n = 5 test matrices: its algorithm gives 33.8 at
# N = 512 and 47.5 at N = 1024 on these inputs, so the honest gate for the GPU at these orders is
# "no worse than the reference algorithm itself", and 32 wherever the reference meets it.
ORACLE_RESIDUAL_N1024_SEED2024_P4 = 47.46


def test_large_pschur_n512_vs_reference_algorithm(psd, oracle):
    n, p = 512, 4
    A = oracle.gen_real(2024, n, p, 1)
    To, Zo, lo, io, _ = oracle.rpschur_batched(A)   # ~7 s on one core
    ro = K.pschur_check(A[0], To[0], Zo[0], lo[0], tol=1e9, check_lambda=False)["residual_eps_a1"]
    T, Z, lam, info = psd.pschur_batched(A, "R")
    assert info[0] == 0 and io[0] == 0
    out = K.pschur_check(A[0], T[0], Z[0], lam[0], tol=max(32.0, 1.1 * ro), check_lambda=False)
    print("N=512 p=4: GPU residual %.1f, reference algorithm %.1f eps*|A|_1" % (out["residual_eps_a1"], ro))
    scale = np.max(np.abs(lo[0]))
    assert K.match_eigs(lo[0], lam[0]) <= 100 * n * EPS * scale
    assert np.count_nonzero(lo[0].imag > 0) == np.count_nonzero(lam[0].imag > 0)
    st = psd.default_handle().large_stats()
    assert st["status"] == 0 and st["sweeps"] >= 1, st  # the multishift iteration did the work (no fallback)


def test_large_pschur_n1024_reference_tolerance(psd, oracle):
    """N = 1024, p = 4 with Schur vectors: pschur_check of the reference (test/testfuncs.jl:56-145)
    with the residual gate set by the reference algorithm's own result on this input (see above),
    the BASELINE gates (10 N eps), the trace identity and the dominant eigenvalue against LAPACK
    on the explicitly formed product."""
    n, p = 1024, 4
    A = oracle.gen_real(2024, n, p, 1)
    T, Z, lam, info = psd.pschur_batched(A, "R")
    assert info[0] == 0
    out = K.pschur_check(A[0], T[0], Z[0], lam[0], tol=ORACLE_RESIDUAL_N1024_SEED2024_P4, check_lambda=False)
    print("N=1024 p=4: residual %.1f eps*|A|_1, orthogonality %.2f N eps" % (out["residual_eps_a1"], out["orth_epsn"]))
    P = np.linalg.multi_dot([K.M(A[0, j]) for j in range(p)])
    tr = np.trace(P)
    assert abs(lam[0].sum() - tr) <= 1e-9 * abs(tr)
    ref = np.linalg.eigvals(P)
    # the dominant (Perron) eigenvalue is well conditioned: relative 1e-10
    assert abs(np.max(np.abs(ref)) - np.max(np.abs(lam[0]))) <= 1e-10 * np.max(np.abs(ref))
    assert psd.default_handle().large_stats()["status"] == 0


@pytest.mark.parametrize("scale", [1e-160, 1e160])
def test_large_badly_scaled(psd, oracle, scale):
    """N = 256 with entries whose squares are not representable: the blocked reduction must keep
    the overflow / underflow safety of the reference's reflector (householder.jl:5-24, 80-100)."""
    n, p = 256, 2
    A = oracle.gen_real(555, n, p, 1)
    As = A.copy()
    As[:, 0] *= scale          # one factor tiny / huge, the other O(1): the product stays representable
    T, Z, lam, info = psd.pschur_batched(As, "R")
    assert info[0] == 0 and np.isfinite(T).all() and np.isfinite(Z).all()
    K.pschur_check(As[0], T[0], Z[0], lam[0], tol=32, check_lambda=False)
    T1, Z1, lam1, info1 = psd.pschur_batched(A, "R")
    s1 = np.max(np.abs(lam1[0]))
    assert K.match_eigs(lam1[0], lam[0] / scale) <= 1e-8 * s1
    H, Q = psd.phessenberg_batched(As)
    assert np.isfinite(H).all()
    _check_hess(As[0], H[0], Q[0])


@pytest.mark.parametrize("n,p,left", [(300, 1, False), (260, 2, True), (280, 6, False), (200, 9, True), (230, 12, False),
                                      (251, 4, False), (197, 3, True)])
def test_large_multishift_shapes(psd, oracle, n, p, left):
    """the multishift iteration at every window geometry (W = 64 for p <= 3, 56, 48, 40, 32 for
    larger periods), :R and :L, held to the reference's predicates and to the oracle's eigenvalues;
    the odd orders take the update kernel's unaligned staging path (8-byte cp.async instead of bulk
    copies)"""
    A = oracle.gen_real(4000 + 10 * p + n, n, p, 1)
    lr = "L" if left else "R"
    T, Z, lam, info = psd.pschur_batched(A, lr)
    assert info[0] == 0
    st = psd.default_handle().large_stats()
    assert st["status"] == 0 and st["rounds"] > 0 and st["final_blocks"] > 0, st
    To, Zo, lo, io, _ = oracle.rpschur_batched(A, left=left)
    ro = K.pschur_check(A[0], To[0], Zo[0], lo[0], left=left, tol=1e9, check_lambda=False)["residual_eps_a1"]
    K.pschur_check(A[0], T[0], Z[0], lam[0], left=left, tol=max(40.0, 1.25 * ro), check_lambda=False)
    scale = np.max(np.abs(lo[0]))
    assert K.match_eigs(lo[0], lam[0]) <= 100 * n * EPS * scale
    assert np.count_nonzero(lo[0].imag > 0) == np.count_nonzero(lam[0].imag > 0)
    # complex pairs adjacent, positive imaginary part first, standardised 2 x 2 blocks (a = d, bc < 0)
    js = p - 1 if left else 0
    T1 = K.M(T[0, js])
    k = 0
    while k < n:
        if lam[0][k].imag != 0:
            assert lam[0][k].imag > 0 and lam[0][k + 1] == np.conj(lam[0][k])
            assert T1[k + 1, k] != 0
            k += 2
        else:
            k += 1
    # eigenvalues only / T only give the same eigenvalues
    _, Z0, lam0, info0 = psd.pschur_batched(A, lr, wantT=False, wantZ=False)
    assert Z0 is None and info0[0] == 0
    assert K.match_eigs(lam[0], lam0[0]) <= 1000 * n * EPS * scale


def test_large_dev_calls_on_two_streams(psd, oracle):
    """two device-resident calls on different streams share the handle's scratch set: the library
    orders them with an event, so both give the result of a call made alone"""
    import ctypes as C
    import torch
    n, p, B = 32, 8, 3000
    L = psd.lib()
    h = psd.Handle([0])
    A = torch.from_numpy(oracle.gen_real(77, n, p, 2 * B)).cuda()
    E = torch.zeros((2 * B, n, 2), dtype=torch.float64, device="cuda")
    I = torch.full((2 * B,), -1, dtype=torch.int32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for k, st in enumerate((s1, s2)):
        psd.capi.check(L.psd_rpschur_batched_dev(
            h.ptr, 0, C.c_void_p(st.cuda_stream), n, p, B, 0, 0, 0, 30,
            C.c_void_p(A[k * B:].data_ptr()), None, C.c_void_p(E[k * B:].data_ptr()), C.c_void_p(I[k * B:].data_ptr())))
    torch.cuda.synchronize()
    assert (I == 0).all()
    _, _, lam, info = psd.pschur_batched(A.cpu().numpy(), "R", wantT=False, wantZ=False, handle=h)
    got = E.cpu().numpy()
    assert np.array_equal(got[..., 0] + 1j * got[..., 1], lam)


def test_large_back_to_back_calls(psd):
    """calls of different shapes on one handle: the scan-result slots of the workspace keep the
    values of earlier calls, so their sequence numbers must never repeat (a second call with 17 - 31
    scans used to leave numbers that the third call took for its own results: residual 1e14 eps)"""
    rng = np.random.default_rng(7)
    for n, p, graded in [(192, 4, False), (192, 4, True), (333, 3, False), (333, 3, True), (260, 2, False)]:
        A = rng.uniform(-1.0, 1.0, size=(1, p, n, n))
        if graded:
            A *= np.logspace(-3, 3, n)[None, None, :, None]
        T, Z, lam, info = psd.pschur_batched(A, "R")
        st = psd.default_handle().large_stats()
        assert info[0] == 0 and st["status"] == 0, (n, p, graded, st)
        K.pschur_check(A[0], T[0], Z[0], lam[0], tol=max(40.0, 0.3 * n), check_lambda=False)  # (BASELINE gate: 10 n)
