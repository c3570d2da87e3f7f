"""GPU edge cases of the batched entry points: empty and odd batch sizes, chunked / sharded
batches, non-convergence reported per problem, zero / triangular / badly scaled inputs
(failure detection, section 8(b): batch semantics)."""
import numpy as np
import pytest

import psd_checks as K

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def test_empty_batch(psd):
    A = np.zeros((0, 3, 5, 5))
    T, Z, lam, info = psd.pschur_batched(A, "R")
    assert T.shape == (0, 3, 5, 5) and lam.shape == (0, 5) and info.shape == (0,)
    out = psd.gpschur_batched(np.zeros((0, 2, 4, 4), dtype=np.complex128), [1, 0], "R")
    assert out[2].shape == (0, 4)


def test_nonconvergence_is_per_problem(psd, oracle):
    """maxitfac too small: info[b] = level at which convergence failed (:891-893), the call itself
    succeeds and returns every problem."""
    A = oracle.gen_real(3, 12, 3, 16)
    T, Z, lam, info = psd.pschur_batched(A, "R", maxitfac=1)
    assert info.shape == (16,) and (info >= 0).all() and (info <= 12).all()
    assert (info > 0).any()
    _, _, lam2, info2 = psd.pschur_batched(A, "R", wantT=False, wantZ=False, maxitfac=1)
    assert (info2 > 0).any() and (info2 <= 12).all()
    # the Python mirror of pschur! raises the reference's error for a single failing problem
    bad = int(np.argmax(info > 0))
    mats = [np.ascontiguousarray(A[bad, j].T) for j in range(3)]
    with pytest.raises(RuntimeError, match="convergence failed at level"):
        psd.pschur_(mats, "R", maxitfac=1)


def test_zero_and_structured_inputs(psd):
    n, p = 9, 4
    Z0 = np.zeros((2, p, n, n))
    T, Z, lam, info = psd.pschur_batched(Z0, "R")
    assert (info == 0).all() and not np.isnan(T).any() and (lam == 0).all()
    _, _, lam, info = psd.pschur_batched(Z0, "R", wantT=False, wantZ=False)
    assert (info == 0).all() and (lam == 0).all()
    # diagonal factors: eigenvalues are the products of the diagonals
    rng = np.random.default_rng(0)
    d = rng.uniform(0.5, 2.0, size=(p, n))
    A = np.zeros((1, p, n, n))
    for j in range(p):
        A[0, j][np.arange(n), np.arange(n)] = d[j]
    for kw in ({}, {"wantT": False, "wantZ": False}):
        _, _, lam, info = psd.pschur_batched(A, "R", **kw)
        assert info[0] == 0
        assert np.allclose(np.sort(lam[0].real), np.sort(np.prod(d, axis=0)), rtol=1e-13)
        assert (lam[0].imag == 0).all()


@pytest.mark.parametrize("scale", [1e-150, 1e120])
def test_badly_scaled_eigs_only(psd, oracle, scale):
    """magnitudes outside the range of the branch-free reflector: the warp kernel must redo the
    problem with the exactly rescaling variant (householder.jl:80-100)"""
    n, p = 16, 3
    A = oracle.gen_real(11, n, p, 8)
    As = A.copy()
    As[:, 0] *= scale
    _, _, lam, info = psd.pschur_batched(As, "R", wantT=False, wantZ=False)
    _, _, lam0, info0 = psd.pschur_batched(A, "R", wantT=False, wantZ=False)
    assert (info == 0).all() and (info0 == 0).all()
    for b in range(8):
        worst = K.match_eigs(lam0[b] * scale, lam[b])
        assert worst <= 100 * n * EPS * np.max(np.abs(lam0[b])) * scale


def test_chunked_and_sharded_batch(psd, oracle):
    """a batch larger than one staging chunk, on a handle that lists the same device twice (the
    multi-device sharding path of the host call, exercised on one GPU)"""
    n, p, B = 32, 8, 9001
    A = oracle.gen_real(1234, n, p, B)
    h2 = psd.Handle([0, 0])
    assert h2.ndev == 2
    _, _, lam2, info2 = psd.pschur_batched(A, "R", wantT=False, wantZ=False, handle=h2)
    _, _, lam1, info1 = psd.pschur_batched(A, "R", wantT=False, wantZ=False)
    assert (info1 == 0).all() and (info2 == 0).all()
    assert np.array_equal(lam1, lam2)  # same kernels, same inputs: bit-identical
    st = h2.stats()
    assert st["h2d_bytes"] == A.nbytes
    # spot-check against the oracle
    _, _, lo, io, _ = oracle.rpschur_batched(A[::997], wantT=False, wantZ=False)
    for k, b in enumerate(range(0, B, 997)):
        assert K.match_eigs(lo[k], lam1[b]) <= 100 * n * EPS * np.max(np.abs(lo[k]))
    h2.close()


def test_generalized_sharded(psd):
    import gpsd_cases as GCs
    A = GCs.rand_storage(9, 6, 3, 11, True)
    S = [1, 0, 1]
    h2 = psd.Handle([0, 0])
    a = psd.gpschur_batched(A, S, "R", handle=h2)
    b = psd.gpschur_batched(A, S, "R")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    h2.close()


def test_pageable_staging_with_T_and_Z(psd, oracle):
    """a pageable call large enough to be cut into several staged chunks (the copies out of the
    pinned staging buffers are deferred until a slot is reused): every problem's T, Z, eigenvalues
    must land in its own place - checked against separate small calls on slices of the batch"""
    n, p, B = 20, 3, 8000                      # 77 MB of factors: at least four chunks
    A = oracle.gen_real(4321, n, p, B)
    T, Z, lam, info, it = psd.pschur_batched(A, "R", return_iters=True)
    assert (info == 0).all() and (it > 0).all()
    for lo in (0, 1999, 3990, 7900):
        sl = slice(lo, lo + 100)
        Ts, Zs, ls, infos, its = psd.pschur_batched(A[sl], "R", return_iters=True)
        assert np.array_equal(T[sl], Ts) and np.array_equal(Z[sl], Zs) and np.array_equal(lam[sl], ls)
        assert np.array_equal(it[sl], its)
    K.pschur_check(A[B - 1], T[B - 1], Z[B - 1], lam[B - 1])
