"""GPU test of the FP64 tensor-core (DMMA) GEMM used by the large-N blocked updates."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(psd, h, ta, tb, M, N, K, alpha, beta, rng, reps=0):
    A = np.asfortranarray(rng.standard_normal((K, M) if ta else (M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if tb else (K, N)))
    Cm = np.asfortranarray(rng.standard_normal((M, N)))
    ref = alpha * (A.T if ta else A) @ (B.T if tb else B) + beta * Cm
    out = Cm.copy(order="F")
    ms = C.c_double(0.0)
    psd.capi.check(psd.lib().psd_dgemm_host(
        h.ptr, ta, tb, M, N, K, alpha, C.c_void_p(A.ctypes.data), A.shape[0],
        C.c_void_p(B.ctypes.data), B.shape[0], beta, C.c_void_p(out.ctypes.data), M, reps, C.byref(ms)))
    return out, ref, ms.value


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 16), (1, 1, 1), (37, 53, 29), (200, 130, 75), (64, 300, 1000),
                                   (513, 64, 2048)])
def test_dgemm_matches_numpy(psd, ta, tb, M, N, K):
    h = psd.default_handle()
    rng = np.random.default_rng(M * 7 + N * 3 + K + ta * 2 + tb)
    for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)):
        out, ref, _ = _gemm(psd, h, ta, tb, M, N, K, alpha, beta, rng)
        assert np.allclose(out, ref, rtol=0, atol=1e-11 * max(1.0, K) ** 0.5 * max(1.0, np.abs(ref).max()))
