"""Numerical prototype (numpy) of the small-bulge multishift periodic QR iteration used by the
large-N path (csrc/psd_ms_*.cuh): convergence experiments only (how many sweeps / shifts per
eigenvalue for a given number of simultaneous shifts, with and without aggressive early
deflation).  Not part of the product and not an oracle.

Conventions: H[0] = H_1 upper Hessenberg, H[1..p-1] = H_2..H_p upper triangular, math
orientation, 0-based; transformations Z_j' H_j Z_{j+1}.
"""
import sys

import numpy as np

EPS = np.finfo(float).eps


def refl(x):
    """dlarfg: returns (v (v[0] = 1), tau, beta)"""
    x = np.array(x, dtype=float)
    alpha = x[0]
    xn = np.linalg.norm(x[1:])
    if xn == 0:
        return np.concatenate([[1.0], np.zeros(len(x) - 1)]), 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def lapply(A, r0, v, tau, c0, c1):
    nr = len(v)
    blk = A[r0:r0 + nr, c0:c1]
    A[r0:r0 + nr, c0:c1] = blk - tau * np.outer(v, v @ blk)


def rapply(A, c0, v, tau, r0, r1):
    nr = len(v)
    blk = A[r0:r1, c0:c0 + nr]
    A[r0:r1, c0:c0 + nr] = blk - tau * np.outer(blk @ v, v)


def bulge_step(H, Z, k, ihi, vstart=None):
    """one step of one bulge: the reflectors act on rows/columns k+1 .. k+nr"""
    p = len(H)
    n = H[0].shape[0]
    nr = min(3, ihi - k)
    if nr < 2:
        return
    r = k + 1
    if vstart is not None:
        v, tau, _ = refl(vstart[:nr])
    else:
        v, tau, beta = refl(H[0][r:r + nr, k])
        H[0][r:r + nr, k] = 0.0
        H[0][r, k] = beta
    rlast = min(r + nr, ihi)  # last row touched by the column operations
    lapply(H[0], r, v, tau, r, n)
    if p > 1:
        rapply(H[p - 1], r, v, tau, 0, rlast + 1)
    else:
        rapply(H[0], r, v, tau, 0, rlast + 1)
    rapply(Z[0], r, v, tau, 0, n)
    for j in range(p - 1, 0, -1):
        v, tau, beta = refl(H[j][r:r + nr, r])
        H[j][r:r + nr, r] = 0.0
        H[j][r, r] = beta
        lapply(H[j], r, v, tau, r + 1, n)
        rapply(H[j - 1], r, v, tau, 0, rlast + 1)
        rapply(Z[j], r, v, tau, 0, n)
        if nr == 3:
            v, tau, beta = refl(H[j][r + 1:r + 3, r + 1])
            H[j][r + 1:r + 3, r + 1] = 0.0
            H[j][r + 1, r + 1] = beta
            lapply(H[j], r + 1, v, tau, r + 2, n)
            rapply(H[j - 1], r + 1, v, tau, 0, rlast + 1)
            rapply(Z[j], r + 1, v, tau, 0, n)


def start_vector(H, ilo, s1, s2):
    """first column of (P - s1)(P - s2), P = H_1 ... H_p restricted to the active block"""
    p = len(H)
    T = np.eye(3)
    for j in range(1, p):
        T = T @ np.triu(H[j][ilo:ilo + 3, ilo:ilo + 3])
    P = H[0][ilo:ilo + 3, ilo:ilo + 3] @ T
    # P e1 = (P00, P10, 0);  P^2 e1 needs P[:, 0:2]
    h11, h21 = P[0, 0], P[1, 0]
    h12, h22, h32 = P[0, 1], P[1, 1], P[2, 1]
    tr = (s1 + s2).real
    det = (s1 * s2).real
    v = np.array([h11 * h11 + h12 * h21 - tr * h11 + det, h21 * (h11 + h22 - tr), h21 * h32])
    s = np.abs(v).sum()
    return v / s if s > 0 else v


def sweep(H, Z, ilo, ihi, shifts):
    """chase len(shifts)/2 tightly packed bulges from ilo to ihi"""
    pairs = [(shifts[i], shifts[i + 1]) for i in range(0, len(shifts) - 1, 2)]
    nb = len(pairs)
    pos = []  # positions k of the bulges in flight (first introduced = largest k)
    nsteps = 0
    nxt = 0
    while nxt < nb or pos:
        # advance every bulge in flight by one step (bottom first)
        newpos = []
        for k in pos:
            k2 = k + 1
            if k2 <= ihi - 2:
                bulge_step(H, Z, k2, ihi)
                nsteps += 1
                newpos.append(k2)
        pos = newpos
        # introduce the next bulge when the top of the block is free
        if nxt < nb and (not pos or pos[-1] >= ilo + 2) and ihi - ilo >= 2:
            v = start_vector(H, ilo, *pairs[nxt])
            bulge_step(H, Z, ilo - 1, ihi, vstart=v)
            nsteps += 1
            pos.append(ilo - 1)
            nxt += 1
    return nsteps


def trailing_shifts(H, ilo, ihi, ns):
    """eigenvalues of the trailing ns x ns block of the product (prototype: explicit product)"""
    lo = max(ilo, ihi - ns + 1)
    P = np.eye(ihi - lo + 1)
    for Hj in H:
        P = P @ Hj[lo:ihi + 1, lo:ihi + 1]
    ev = np.linalg.eigvals(P)
    # pair up: complex pairs adjacent, reals two by two
    cp = [e for e in ev if e.imag > 0]
    re = sorted([e.real for e in ev if e.imag == 0])
    out = []
    for e in cp:
        out += [e, np.conj(e)]
    for i in range(0, len(re) - 1, 2):
        out += [re[i], re[i + 1]]
    if len(re) % 2:
        out += [re[-1], re[-1]]
    return out


def deflate_scan(H, ilo, ihi):
    """zero negligible subdiagonals of H_1 in [ilo, ihi]; returns number found"""
    H1 = H[0]
    cnt = 0
    for k in range(ihi, ilo, -1):
        if H1[k, k - 1] != 0 and abs(H1[k, k - 1]) <= EPS * (abs(H1[k - 1, k - 1]) + abs(H1[k, k])):
            H1[k, k - 1] = 0.0
            cnt += 1
    return cnt


def aed(H, Z, ilo, ihi, nw):
    """aggressive early deflation without reordering, on the trailing nw x nw window.
    Prototype: uses scipy-free explicit periodic Schur via the single-bulge iteration on a copy.
    Returns (number deflated, shifts)."""
    raise NotImplementedError


def solve(H, Z, ns, nmin, maxsweeps=10000, rep=1, verbose=True, stale=0):
    n = H[0].shape[0]
    ihi = n - 1
    sweeps = 0
    steps = 0
    shifts_used = 0
    hist = []
    while ihi >= 0 and sweeps < maxsweeps:
        # find the active block
        ilo = ihi
        while ilo > 0 and H[0][ilo, ilo - 1] != 0:
            ilo -= 1
        m = ihi - ilo + 1
        if m <= nmin:
            ihi = ilo - 1  # deferred small block
            continue
        k = min(ns, (m // 3) * 2)
        k = max(2, k - k % 2)
        sh = trailing_shifts(H, ilo, ihi, k)
        if stale:
            # shifts computed from the state before the previous sweep (pipelined shift solve)
            key = (ilo,)
            prev = solve.pending.get(key)
            solve.pending = {key: sh}
            if prev is not None and len(prev) >= 2:
                sh = prev[:k] if len(prev) >= k else prev
        sh = sh * rep
        steps += sweep(H, Z, ilo, ihi, sh)
        shifts_used += len(sh)
        sweeps += 1
        d = deflate_scan(H, ilo, ihi)
        hist.append((ilo, ihi, len(sh), d))
        if verbose and sweeps % 10 == 0:
            print(f"  sweep {sweeps}: block [{ilo},{ihi}] ns {len(sh)} deflated {d}", flush=True)
    return sweeps, steps, shifts_used, hist


solve.pending = {}


def hess_tri(A):
    """periodic Hessenberg-triangular reduction (unblocked)"""
    p = len(A)
    n = A[0].shape[0]
    H = [a.copy() for a in A]
    Z = [np.eye(n) for _ in range(p)]
    for i in range(n - 1):
        for j in range(p - 1, 0, -1):
            v, tau, beta = refl(H[j][i:, i])
            H[j][i:, i] = 0
            H[j][i, i] = beta
            lapply(H[j], i, v, tau, i + 1, n)
            rapply(H[j - 1], i, v, tau, 0, n)
            rapply(Z[j], i, v, tau, 0, n)
        if i + 2 <= n - 1 or True:
            if n - (i + 1) >= 1:
                v, tau, beta = refl(H[0][i + 1:, i])
                H[0][i + 1:, i] = 0
                H[0][i + 1, i] = beta
                lapply(H[0], i + 1, v, tau, i + 1, n)
                rapply(H[p - 1] if p > 1 else H[0], i + 1, v, tau, 0, n)
                rapply(Z[0], i + 1, v, tau, 0, n)
    return H, Z


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ns = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    nmin = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    rep = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    stale = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    rng = np.random.default_rng(1)
    A = [rng.random((n, n)) for _ in range(p)]
    H, Z = hess_tri(A)
    for j in range(p):
        zn = Z[(j + 1) % p]
        assert np.linalg.norm(Z[j].T @ A[j] @ zn - H[j]) < 1e-10 * n
    sweeps, steps, su, hist = solve(H, Z, ns, nmin, rep=rep, stale=stale, verbose=False)
    res = max(np.linalg.norm(Z[j].T @ A[j] @ Z[(j + 1) % p] - H[j]) / np.linalg.norm(A[j]) for j in range(p))
    # block sizes
    sub = np.abs(np.diag(H[0], -1)) > 0
    blocks = []
    c = 1
    for s in sub:
        if s:
            c += 1
        else:
            blocks.append(c)
            c = 1
    blocks.append(c)
    print(f"n={n} p={p} ns={ns} nmin={nmin} rep={rep} stale={stale}: sweeps {sweeps}, bulge steps {steps} "
          f"(= {steps / n**2:.2f} n^2), shifts/eig {su / n:.2f}, residual {res / (n * EPS):.3f} n eps, "
          f"max block {max(blocks)}, blocks>2: {sum(b > 2 for b in blocks)}")
