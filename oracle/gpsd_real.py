"""ORACLE — TEST INFRASTRUCTURE ONLY (see gpsd_common.py header).

numpy restatement of the REAL generalized periodic Schur path of the reference:
  pschur!(A, S, lr; wantZ, wantT)            src/rgeneralized.jl:3-45       -> rgpschur
  pschur!(H1, Hs, S; ...) (MB03BD-style)     src/rgeneralized.jl:49-1083    -> rpqz
      deflation skeleton (:169-648) shared with the complex path: gpsd_complex.cpqz
      2x2 block handling                      :661-790                      -> _step_2x2
      double-shift sweep                      :890-1054                     -> _step_sweep
  _qzrots   (MB03AF 'Double')                 :1140-1359
  _qzrot2x2 (MB03AF 'Single', N = 2)          :1364-1396
  _rp2x2ssr! (MB03BF)                         src/rpschur2x2.jl:280-317
  _rpeigvals2x2 (MB03BB-like)                 src/rpschur2x2.jl:9-235
  _sanitize_reigpair!                         src/rpschur2x2.jl:238-275

Deviation (documented in DESIGN.md): the reference alternates 10 implicit-shift sweeps
(_qzrots) with 1 explicit-shift sweep (_rpeigvals2x2 + _shift2rot, :804-887); that branch
reads undefined names on its fallback (`hnorm`, :826), calls _qzrots with 4 of 5 arguments
(:840) and overrides its windowed view (`Ai = H1`, :1456-1457), so
every sweep here uses the implicit shifts of _qzrots.  Shifts only influence convergence speed.
"""
from __future__ import annotations

import math

import numpy as np

from .gpsd_common import EPS, givens, gphessenberg, lmul_g, phessenberg_householder, rmul_gadj
from .gpsd_complex import cpqz


def qzrots(H1, Hs, S, i1, n):
    """rgeneralized.jl:1140-1359.  i1 1-based start of the active block, n its order."""
    p = len(Hs) + 1
    o = i1 - 1  # 0-based offset of the window

    def W(Hm, r, c):  # Hlv[r, c] with Hlv = view(Hm, i1:nh, i1:nh), 1-based
        return Hm[o + r - 1, o + c - 1]

    c1, s1, r = givens(H1[i1 - 1, i1 - 1], H1[i1, i1 - 1])
    c2, s2, r = givens(r, 1.0)
    i2 = i1 + n - 1
    for l in range(p, 1, -1):
        Hl = Hs[l - 2]
        h11, h12, h22 = Hl[i1 - 1, i1 - 1], Hl[i1 - 1, i1], Hl[i1, i1]
        hnn = Hl[i2 - 1, i2 - 1]
        if S[l - 1]:
            a = c2 * (c1 * h11 + s1 * h12)
            b = s1 * c2 * h22
            g = s2 * hnn
            c1, s1, r = givens(a, b)
            c2, s2, _ = givens(r, g)
        else:
            a = c1 * s2 * h11
            g = s1 * h11
            b = s2 * (c1 * h12 + s1 * h22)
            d = c1 * h22 - s1 * h12
            c1, s1, r = givens(d, g)
            a = c1 * a + s1 * b
            b = c2 * hnn
            c2, s2, r = givens(b, a)
    a = s2 * H1[i2 - 1, i2 - 1] - c1 * c2
    b = -s1 * c2
    m = n - 1
    g = -s2 * W(H1, n, m)
    c2, s2, r = givens(a, g)
    c1, s1, r = givens(r, b)
    cx = c1 * c2
    sx = c1 * s2
    b = s1 * W(H1, n, m)
    a = cx * W(H1, n, m) + sx * W(H1, n, n)
    g = s1 * W(H1, m, m)
    d = cx * W(H1, m, m) + sx * W(H1, m, n)
    v1 = s1 * W(H1, 3, 2)
    v2 = cx * W(H1, 2, 1) + s1 * W(H1, 2, 2)
    v3 = cx * W(H1, 1, 1) + s1 * W(H1, 1, 2)
    c1, s1, r = givens(a, b)
    c2, s2, r = givens(g, r)
    c3, s3, r = givens(d, r)
    c4, s4, r = givens(v1, r)
    c5, s5, r = givens(v2, r)
    c6, s6, r = givens(v3, r)
    for i in range(p, 1, -1):
        Hi = Hs[i - 2]
        if S[i - 1]:
            ss = s3 * s4
            sss = s2 * ss
            ssss = s1 * sss
            v1 = c4 * W(Hi, 1, 3)
            v2 = c4 * W(Hi, 2, 3)
            v3 = c4 * W(Hi, 3, 3)
            a = s4 * c3 * W(Hi, m, m) + sss * c1 * W(Hi, m, n)
            b = ss * c2 * W(Hi, m, m) + ssss * W(Hi, m, n)
            g = sss * c1 * W(Hi, n, n)
            d = ssss * W(Hi, n, n)
            ss = s5 * s6
            cs = c5 * s6
            v1 = ss * v1 + cs * W(Hi, 1, 2) + c6 * W(Hi, 1, 1)
            v2 = ss * v2 + cs * W(Hi, 2, 2)
            v3 = ss * v3
            a, b, g, d = ss * a, ss * b, ss * g, ss * d
            c1, s1, r = givens(g, d)
            c2, s2, r = givens(b, r)
            c3, s3, r = givens(a, r)
            c4, s4, r = givens(v3, r)
            c5, s5, r = givens(v2, r)
            c6, s6, r = givens(v1, r)
        else:
            d = c1 * W(Hi, n, n)
            e = s1 * W(Hi, n, n)
            a = c2 * W(Hi, m, m)
            b = s2 * d
            g = -s2 * W(Hi, m, m)
            z = c2 * W(Hi, m, n) + s2 * e
            eta = -s2 * W(Hi, m, n) + c2 * e
            d = c1 * c2 * d + s1 * eta
            c2R, s2R, r = givens(d, -g)
            d = c3 * W(Hi, m, m)
            e = s3 * a
            eta = c3 * W(Hi, m, n) + s3 * b
            th = s3 * z
            g = -s3 * W(Hi, m, m)
            b = -s3 * W(Hi, m, n) + c3 * b
            a = c2R * c3 * a + s2R * (c1 * b + s1 * c3 * z)
            c3R, s3R, r = givens(a, -g)
            v1 = c4 * W(Hi, 3, 3)
            v2 = s4 * d
            v3 = s4 * e
            v4 = s4 * eta
            v5 = s4 * th
            b = -s4 * W(Hi, 3, 3)
            d = c4 * d
            e = c4 * e
            z = c4 * eta
            eta = c4 * th
            a = c3R * d + s3R * (c2R * e + s2R * (c1 * z + s1 * eta))
            c4R, s4R, r = givens(a, -b)
            b = c5 * W(Hi, 2, 2)
            d = c5 * W(Hi, 2, 3) + s5 * v1
            e = s5 * v2
            z = s5 * v3
            eta = s5 * v4
            th = s5 * v5
            g = -s5 * W(Hi, 2, 2)
            v1 = c5 * v1 - s5 * W(Hi, 2, 3)
            v2, v3, v4, v5 = c5 * v2, c5 * v3, c5 * v4, c5 * v5
            a = c4R * v1 + s4R * (c3R * v2 + s3R * (c2R * v3 + s2R * (c1 * v4 + s1 * v5)))
            c5R, s5R, r = givens(a, -g)
            g = -s6 * W(Hi, 1, 1)
            b = c6 * b - s6 * W(Hi, 1, 2)
            d = c6 * d - s6 * W(Hi, 1, 3)
            e, z, eta, th = c6 * e, c6 * z, c6 * eta, c6 * th
            a = c5R * b + s5R * (c4R * d + s4R * (c3R * e + s3R * (c2R * z + s2R * (c1 * eta + s1 * th))))
            c6R, s6R, r = givens(a, -g)
            c2, s2, c3, s3, c4, s4 = c2R, s2R, c3R, s3R, c4R, s4R
            c5, s5, c6, s6 = c5R, s5R, c6R, s6R
    v1 = s5 * s6
    v2 = s4 * v1
    v3 = s3 * v2
    a = c3 * v2 - c6
    b = c2 * v3 - c5 * s6
    g = -c4 * v1
    c2, s2, r = givens(b, g)
    c1, s1, r = givens(a, r)
    return c1, s1, c2, s2


def qzrot2x2(H2s, S):
    """rgeneralized.jl:1364-1396 (Hessenberg in last place)."""
    p = len(H2s)
    Hl = H2s[p - 1]
    c1, s1, r = givens(Hl[0, 0], Hl[1, 0])
    c2, s2, r = givens(r, 1.0)
    for l in range(p - 1, 0, -1):
        Hl = H2s[l - 1]
        if S[l - 1]:
            a = c2 * (c1 * Hl[0, 0] + s1 * Hl[0, 1])
            b = s1 * c2 * Hl[1, 1]
            g = s2 * Hl[1, 1]
            c1, s1, r = givens(a, b)
            c2, s2, _ = givens(r, g)
        else:
            a = c1 * s2 * Hl[0, 0]
            g = s1 * Hl[0, 0]
            b = s2 * (c1 * Hl[0, 1] + s1 * Hl[1, 1])
            d = c1 * Hl[1, 1] - s1 * Hl[0, 1]
            c1, s1, r = givens(d, g)
            a = c1 * a + s1 * b
            b = c2 * Hl[1, 1]
            c2, s2, r = givens(b, a)
    Hl = H2s[p - 1]
    a = s2 * Hl[1, 1] - c1 * c2
    b = -s1 * c2
    c1, s1, _ = givens(a, b)
    return c1, s1


def rp2x2ssr(H2s, S, maxit=20):
    """rpschur2x2.jl:280-317 (MB03BF): real single-shift 2x2 periodic QZ, Hessenberg last."""
    p = len(H2s)
    done = False
    for _ in range(maxit):
        c, s = qzrot2x2(H2s, S)
        rmul_gadj(H2s[p - 1], 1, 2, c, s, 1, 2)
        for l in range(1, p):
            Hl = H2s[l - 1]
            if S[l - 1]:
                lmul_g(Hl, 1, 2, c, s, 1, 2)
                c, s, r = givens(Hl[1, 1], -Hl[1, 0])
                Hl[1, 1] = r
                Hl[1, 0] = 0.0
                Hl[0, 0], Hl[0, 1] = (c * Hl[0, 0] + s * Hl[0, 1], c * Hl[0, 1] - s * Hl[0, 0])
            else:
                rmul_gadj(Hl, 1, 2, c, s, 1, 2)
                c, s, r = givens(Hl[0, 0], Hl[1, 0])
                Hl[0, 0] = r
                Hl[1, 0] = 0.0
                Hl[0, 1], Hl[1, 1] = (c * Hl[0, 1] + s * Hl[1, 1], c * Hl[1, 1] - s * Hl[0, 1])
        Hl = H2s[p - 1]
        lmul_g(Hl, 1, 2, c, s, 1, 2)
        done = abs(Hl[1, 0]) < EPS * max(abs(Hl[0, 0]), abs(Hl[0, 1]), abs(Hl[1, 1]))
        if done:
            break
    return done


def sanitize_reigpair(alpha, beta, scal):
    """rpschur2x2.jl:238-275."""
    good = True
    if any(a.imag != 0 for a in alpha):
        sl = scal[0] - scal[1]
        if sl >= 0:
            zt1 = alpha[1] * 2.0 ** (-sl)
            zt2 = alpha[0] - np.conj(zt1)
            cst = alpha[0].imag
        else:
            zt1 = alpha[0] * 2.0 ** sl
            zt2 = alpha[1] - np.conj(zt1)
            cst = alpha[1].imag
        misr = math.hypot(cst, zt1.imag)
        misc = abs(zt2) / 2
        cs = max(abs(alpha[0]), 1.0, abs(alpha[1]))
        good = min(misr, misc) <= cs * math.sqrt(EPS)
        if misr > misc:
            j = 0 if scal[0] >= scal[1] else 1
            at = (alpha[j] + np.conj(zt1)) / 2
            ai = abs(at.imag)
            alpha[0] = complex(at.real, ai)
            alpha[1] = np.conj(alpha[0])
        else:
            for j in range(2):
                alpha[j] = complex(alpha[j].real, 0.0)
    return good


def rpeigvals2x2(blocks, S, recip=False):
    """rpschur2x2.jl:9-235 with schurindex = 1 and natural order: blocks[l] is the 2x2 block of
    factor l+1 (blocks[0] the quasi-triangular one).  Returns alpha[2], beta[2], scal[2],
    converged, good."""
    k = len(blocks)
    Xs = [np.array(b, dtype=np.complex128) for b in blocks]
    X1 = Xs[0]
    converged = False
    for it in range(1, 81):
        lhs = abs(X1[1, 0])
        rhs = max(abs(X1[0, 0]), abs(X1[1, 1]))
        if rhs == 0:
            rhs = abs(X1[0, 1])
        if lhs <= EPS * rhs:
            converged = True
            break
        if it == 1:
            c, s, r = givens(1.0 - 2.0j, 2.0 + 2.0j)
        elif it % 40 == 0:
            c, s, r = givens(complex(k, 1.0), 1.0 - 2.0j)
        else:
            c, s = 1.0, 0.0 + 0j
            ct, st, r = givens(1.0 + 0j, 1.0 + 0j)
            for l in range(k, 1, -1):
                Xl = Xs[l - 1]
                z11, z21, z12, z22 = Xl[0, 0], Xl[1, 0], Xl[0, 1], Xl[1, 1]
                Zm = np.array([[z11, 0, 0], [0, z11, z12], [0, z21, z22]], dtype=np.complex128)
                if bool(S[l - 1]) != recip:
                    rmul_gadj(Zm, 1, 3, ct, st, 1, 3)
                    rmul_gadj(Zm, 1, 2, c, s, 1, 3)
                    ct, st, r = givens(Zm[0, 0], Zm[2, 0])
                    c, s, r = givens(z11, Zm[1, 0])
                else:
                    lmul_g(Zm, 1, 3, ct, st, 1, 3)
                    lmul_g(Zm, 1, 2, c, s, 1, 3)
                    ct, st, r = givens(Zm[2, 2], Zm[2, 0])
                    Zm[2, 2] = r
                    st = -st
                    rmul_gadj(Zm, 1, 3, ct, st, 1, 2)
                    c, s, r = givens(Zm[1, 1], Zm[1, 0])
                    Zm[1, 1] = r
                    s = -s
            Xl = Xs[0]
            z11, z21, z22 = Xl[0, 0], Xl[1, 0], Xl[1, 1]
            Zm = np.array([[z11, -z21, -z22], [z21, 0, 0]], dtype=np.complex128)
            rmul_gadj(Zm, 1, 3, ct, st, 1, 2)
            rmul_gadj(Zm, 1, 2, c, s, 1, 2)
            c, s, r = givens(Zm[0, 0], Zm[1, 0])
        ct, st = c, s
        for l in range(k, 1, -1):
            Y = Xs[l - 1]
            if bool(S[l - 1]) != recip:
                rmul_gadj(Y, 1, 2, c, s, 1, 2)
                c, s, r = givens(Y[0, 0], Y[1, 0])
                Y[0, 0] = r
                Y[1, 0] = 0
                lmul_g(Y, 1, 2, c, s, 2, 2)
            else:
                lmul_g(Y, 1, 2, c, s, 1, 2)
                c, s, r = givens(Y[1, 1], Y[1, 0])
                Y[1, 1] = r
                Y[1, 0] = 0
                s = -s
                rmul_gadj(Y, 1, 2, c, s, 1, 1)
        Y = Xs[0]
        lmul_g(Y, 1, 2, ct, st, 1, 2)
        rmul_gadj(Y, 1, 2, c, s, 1, 2)
    beta = [1.0, 1.0]
    scal = [0, 0]
    alpha = [1.0 + 0j, 1.0 + 0j]
    for j in range(2):
        aj = 1.0 + 0j
        for l in range(1, k + 1):
            z = Xs[l - 1][j, j]
            rhs = abs(z)
            if rhs != 0:
                sl = math.floor(math.log2(rhs))
                z = z * 2.0 ** (-sl)
            else:
                sl = 0
            if S[l - 1]:
                aj *= z
                scal[j] += sl
            elif rhs == 0:
                beta[j] = 0.0
            else:
                aj /= z
                scal[j] -= sl
            if l % 10 == 0 or l == k:
                rhs = abs(aj)
                if rhs == 0:
                    scal[j] = 0
                else:
                    sl = math.floor(math.log2(rhs))
                    aj *= 2.0 ** (-sl)
                    scal[j] += sl
        alpha[j] = aj
    if alpha[1].imag > 0:
        alpha[0], alpha[1] = alpha[1], alpha[0]
        beta[0], beta[1] = beta[1], beta[0]
        scal[0], scal[1] = scal[1], scal[0]
    good = sanitize_reigpair(alpha, beta, scal)
    return alpha, beta, scal, converged, good


def _chain(H1, Hs, S, Z, wantZ, j, c, s, ifirstm, ilastm, n, h1rows):
    """The rotation chain of :717-742 / :1022-1048: G on rows (j, j+1) of H1 from the left,
    propagated through factors p..2, back onto H1 from the right (rows ifirstm..h1rows)."""
    p = len(Hs) + 1
    lmul_g(H1, j, j + 1, c, s, j, ilastm)
    if wantZ:
        rmul_gadj(Z[0], j, j + 1, c, s, 1, n)
    for l in range(p, 1, -1):
        Hl = Hs[l - 2]
        if S[l - 1]:
            rmul_gadj(Hl, j, j + 1, c, s, ifirstm, j + 1)
            c, s, r = givens(Hl[j - 1, j - 1], Hl[j, j - 1])
            Hl[j - 1, j - 1] = r
            Hl[j, j - 1] = 0.0
            lmul_g(Hl, j, j + 1, c, s, j + 1, ilastm)
        else:
            lmul_g(Hl, j, j + 1, c, s, j, ilastm)
            c, s, r = givens(Hl[j, j], -Hl[j, j - 1])
            Hl[j, j] = r
            Hl[j, j - 1] = 0.0
            rmul_gadj(Hl, j, j + 1, c, s, ifirstm, j)
        if wantZ:
            rmul_gadj(Z[l - 1], j, j + 1, c, s, 1, n)
    rmul_gadj(H1, j, j + 1, c, s, ifirstm, h1rows)


def real_step(H1, Hs, S, Z, wantZ, ifirst, ilast, ifirstm, ilastm, alpha, beta, ascale):
    """rgeneralized.jl:661-1054.  Returns None, or True when a complex 2x2 block was split."""
    p = len(Hs) + 1
    n = H1.shape[0]

    def Hm(l):
        return H1 if l == 1 else Hs[l - 2]

    if ifirst + 1 == ilast:  # ---- 2x2 block (:661-790) ----
        j = ilast - 1
        S2 = [S[(l + 1) % p] for l in range(p)]  # circshift(S, -1)
        H2s = [np.array(Hm(l + 1 if l < p else 1)[j - 1:j + 1, j - 1:j + 1], dtype=np.float64)
               for l in range(1, p + 1)]
        done2 = False
        titer = 0
        while not done2 and titer < 2:
            titer += 1
            rp2x2ssr(H2s, S2)
            H2p = H2s[p - 1]
            if abs(H2p[1, 0]) < EPS * max(abs(H2p[0, 0]), abs(H2p[0, 1]), abs(H2p[1, 1])):
                done2 = True
                c1, s1 = 1.0, 1.0
                for l in range(p, 1, -1):
                    r = H2s[l - 2][1, 1]
                    Hl = Hs[l - 2]
                    if S[l - 1]:
                        c1, s1, r = givens(c1 * Hl[j - 1, j - 1], s1 * r)
                    else:
                        c1, s1, r = givens(c1 * r, s1 * Hl[j - 1, j - 1])
                r = H2s[p - 1][1, 1]
                c1, s1, r = givens(c1 * H1[j - 1, j - 1] - r * s1, c1 * H1[j, j - 1])
                _chain(H1, Hs, S, Z, wantZ, j, c1, s1, ifirstm, ilastm, n, ilastm)
        if not done2:
            blocks = [Hm(l)[j - 1:j + 1, j - 1:j + 1] for l in range(1, p + 1)]
            a2, b2, s2, conv, good = rpeigvals2x2(blocks, S, recip=False)
            alpha[j - 1:j + 1] = a2
            beta[j - 1:j + 1] = b2
            ascale[j - 1:j + 1] = s2
            return True
        return None
    # ---- double-shift sweep (:796-1054), implicit shifts ----
    c1, s1, c2, s2 = qzrots(H1, Hs, S, ifirst, ilast - ifirst + 1)
    if p > 1:
        # initial transformation enters between H1 and H2 (:890-943)
        j = ifirst
        g1 = (c2, s2)
        g2 = (c1, s1)
        rmul_gadj(H1, j + 1, j + 2, g1[0], g1[1], ifirstm, ilast)
        rmul_gadj(H1, j, j + 1, g2[0], g2[1], ifirstm, ilast)
        if wantZ:
            rmul_gadj(Z[1], j + 1, j + 2, g1[0], g1[1], 1, n)
            rmul_gadj(Z[1], j, j + 1, g2[0], g2[1], 1, n)
        for l in range(2, p + 1):
            Hl = Hs[l - 2]
            if S[l - 1]:
                lmul_g(Hl, j + 1, j + 2, g1[0], g1[1], j, ilastm)
                c, s, r = givens(Hl[j + 1, j + 1], -Hl[j + 1, j])
                Hl[j + 1, j + 1] = r
                Hl[j + 1, j] = 0.0
                g1 = (c, s)
                rmul_gadj(Hl, j + 1, j + 2, c, s, ifirstm, j + 1)
                lmul_g(Hl, j, j + 1, g2[0], g2[1], j, ilastm)
                c, s, r = givens(Hl[j, j], -Hl[j, j - 1])
                Hl[j, j] = r
                Hl[j, j - 1] = 0.0
                g2 = (c, s)
                rmul_gadj(Hl, j, j + 1, c, s, ifirstm, j)
            else:
                rmul_gadj(Hl, j + 1, j + 2, g1[0], g1[1], ifirstm, j + 2)
                c, s, r = givens(Hl[j, j], Hl[j + 1, j])
                Hl[j, j] = r
                Hl[j + 1, j] = 0.0
                g1 = (c, s)
                lmul_g(Hl, j + 1, j + 2, c, s, j + 2, ilastm)
                rmul_gadj(Hl, j, j + 1, g2[0], g2[1], ifirstm, j + 1)
                c, s, r = givens(Hl[j - 1, j - 1], Hl[j, j - 1])
                Hl[j - 1, j - 1] = r
                Hl[j, j - 1] = 0.0
                g2 = (c, s)
                lmul_g(Hl, j, j + 1, c, s, j + 1, ilastm)
            if wantZ:
                ln = (l % p) + 1
                rmul_gadj(Z[ln - 1], j + 1, j + 2, g1[0], g1[1], 1, n)
                rmul_gadj(Z[ln - 1], j, j + 1, g2[0], g2[1], 1, n)
        lmul_g(H1, j + 1, j + 2, g1[0], g1[1], ifirst, ilastm)
        lmul_g(H1, j, j + 1, g2[0], g2[1], ifirst, ilastm)
        i1, i2 = ifirst + 1, ilast - 2
    else:
        i1, i2 = ifirst - 1, ilast - 3
        g1 = (c2, s2)
        g2 = (c1, s1)
    j = ifirst  # only used by the p == 1 recurrence below
    for j1 in range(i1, i2 + 1):
        if j1 < ifirst:
            j = j1 + 1
            lmul_g(H1, j + 1, j + 2, g1[0], g1[1], j, ilastm)
            lmul_g(H1, j, j + 1, g2[0], g2[1], j, ilastm)
        else:
            j = (j + 1) if p == 1 else j1
            c2, s2, r2 = givens(H1[j, j - 2], H1[j + 1, j - 2])
            c1, s1, r1 = givens(H1[j - 1, j - 2], r2)
            H1[j - 1, j - 2] = r1
            H1[j, j - 2] = 0.0
            H1[j + 1, j - 2] = 0.0
            g1 = (c2, s2)
            g2 = (c1, s1)
            lmul_g(H1, j + 1, j + 2, c2, s2, j, ilastm)
            lmul_g(H1, j, j + 1, c1, s1, j, ilastm)
        if wantZ:
            rmul_gadj(Z[0], j + 1, j + 2, g1[0], g1[1], 1, n)
            rmul_gadj(Z[0], j, j + 1, g2[0], g2[1], 1, n)
        for l in range(p, 1, -1):
            Hl = Hs[l - 2]
            if S[l - 1]:
                rmul_gadj(Hl, j + 1, j + 2, g1[0], g1[1], ifirstm, j + 2)
                c, s, r = givens(Hl[j, j], Hl[j + 1, j])
                Hl[j, j] = r
                Hl[j + 1, j] = 0.0
                g1 = (c, s)
                lmul_g(Hl, j + 1, j + 2, c, s, j + 2, ilastm)
                rmul_gadj(Hl, j, j + 1, g2[0], g2[1], ifirstm, j + 1)
                c, s, r = givens(Hl[j - 1, j - 1], Hl[j, j - 1])
                Hl[j - 1, j - 1] = r
                Hl[j, j - 1] = 0.0
                g2 = (c, s)
                lmul_g(Hl, j, j + 1, c, s, j + 1, ilastm)
            else:
                lmul_g(Hl, j + 1, j + 2, g1[0], g1[1], j, ilastm)
                c, s, r = givens(Hl[j + 1, j + 1], -Hl[j + 1, j])
                Hl[j + 1, j + 1] = r
                Hl[j + 1, j] = 0.0
                g1 = (c, s)
                rmul_gadj(Hl, j + 1, j + 2, c, s, ifirstm, j + 1)
                lmul_g(Hl, j, j + 1, g2[0], g2[1], j, ilastm)
                c, s, r = givens(Hl[j, j], -Hl[j, j - 1])
                Hl[j, j] = r
                Hl[j, j - 1] = 0.0
                g2 = (c, s)
                rmul_gadj(Hl, j, j + 1, c, s, ifirstm, j)
            if wantZ:
                rmul_gadj(Z[l - 1], j + 1, j + 2, g1[0], g1[1], 1, n)
                rmul_gadj(Z[l - 1], j, j + 1, g2[0], g2[1], 1, n)
        lm = min(j + 3, ilastm)
        rmul_gadj(H1, j + 1, j + 2, g1[0], g1[1], ifirstm, lm)
        rmul_gadj(H1, j, j + 1, g2[0], g2[1], ifirstm, lm)
    # trailing single rotation (:1015-1048)
    j = ilast - 1
    c1, s1, r1 = givens(H1[j - 1, j - 2], H1[j, j - 2])
    H1[j - 1, j - 2] = r1
    H1[j, j - 2] = 0.0
    _chain(H1, Hs, S, Z, wantZ, j, c1, s1, ifirstm, ilastm, n, ilastm)
    return None


def rpqz(H1, Hs, S, wantZ=True, wantT=True, Q=None, maxitfac=120, rev=False):
    """rgeneralized.jl:49-1083 (complex-pair 2x2 blocks are left unstandardised in T1)."""
    for Hl in Hs:
        Hl[...] = np.triu(Hl)
    return cpqz(H1, Hs, S, wantZ=wantZ, wantT=wantT, Q=Q, maxitfac=maxitfac, rev=rev,
                real_step=real_step)


def rgpschur(A, S, lr="R", wantZ=True, wantT=True, maxitfac=120):
    """rgeneralized.jl:3-45 on a list of float64 math-orientation matrices (copied)."""
    p = len(A)
    A = [np.array(a, dtype=np.float64) for a in A]
    S = [bool(x) for x in S]
    left = lr == "L"
    if left:
        Aarg = [A[p - 1 - j] for j in range(p)]
        Sarg = list(reversed(S))
    else:
        Aarg, Sarg = A, S
    if all(S):
        Q = phessenberg_householder(Aarg)
        H1 = np.triu(Aarg[0], -1)
        Hs = [np.triu(Aarg[j]) for j in range(1, p)]
    else:
        if not Sarg[0]:
            raise ValueError("The leftmost entry in S must be true")
        Q = gphessenberg(Aarg, Sarg)
        H1 = Aarg[0]
        Hs = Aarg[1:]
    return rpqz(H1, Hs, Sarg, wantZ=wantZ, wantT=wantT, Q=(Q if wantZ else None),
                maxitfac=maxitfac, rev=left)
