"""CPU test of the numpy restatement of _rphessenberg! (oracle/rhessx.py)."""
import numpy as np
import pytest

import psd_checks as K
import rowhess_cases as RC
from oracle import rhessx as OR


@pytest.mark.parametrize("n,p,extra", [(5, 1, False), (5, 3, False), (6, 3, True), (7, 1, True), (12, 4, True)])
def test_rowhess_oracle(n, p, extra):
    Ap0, A0, Q0 = RC.make(11, n, p, extra, 2)
    for b in range(2):
        Ap = np.ascontiguousarray(K.M(Ap0[b]))
        A = [np.ascontiguousarray(K.M(A0[b, l])) for l in range(p - 1)]
        Q = [np.eye(n) for _ in range(p)]
        OR.rphessenberg(Ap, A, Q)
        Aps = np.ascontiguousarray(Ap.T)
        As = np.array([a.T for a in A]) if p > 1 else None
        Qs = np.array([q.T for q in Q])
        RC.check(Ap0[b], None if A0 is None else A0[b], Aps, As, Qs, n, p)
