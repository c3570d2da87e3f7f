"""GPU parity test of the row-wise periodic Hessenberg reduction (scope row a21,
rhessx.jl:53-109) against its numpy restatement and the structural predicates."""
import numpy as np
import pytest

import psd_checks as K
import rowhess_cases as RC
from oracle import rhessx as OR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,extra", [(5, 1, False), (5, 3, False), (6, 3, True), (7, 1, True),
                                       (20, 3, True), (30, 5, False)])
def test_rowhess_gpu(psd, n, p, extra):
    B = 4
    Ap0, A0, Q0 = RC.make(11, n, p, extra, B)
    Ap, A, Q = psd.rphessenberg_rowwise_batched(Ap0, A0, Q0)
    for b in range(B):
        RC.check(Ap0[b], None if A0 is None else A0[b], Ap[b], None if A is None else A[b], Q[b], n, p)
        # against the oracle: same Hessenberg / triangular factors up to column signs
        Apo = np.ascontiguousarray(K.M(Ap0[b]))
        Ao = [np.ascontiguousarray(K.M(A0[b, l])) for l in range(p - 1)]
        OR.rphessenberg(Apo, Ao, None)
        assert np.allclose(np.abs(K.M(Ap[b])), np.abs(Apo), atol=1e-10 * max(1.0, np.abs(Apo).max()))
        for l in range(p - 1):
            assert np.allclose(np.abs(K.M(A[b, l])), np.abs(Ao[l]), atol=1e-10 * max(1.0, np.abs(Ao[l]).max()))
