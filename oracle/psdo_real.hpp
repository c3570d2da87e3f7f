// ORACLE — TEST INFRASTRUCTURE ONLY (see psdo_common.hpp header).
//
// Real standard periodic Schur path: restatement of
//   src/PeriodicSchurDecompositions.jl:120-152  (driver pschur!(A, lr))
//   src/PeriodicSchurDecompositions.jl:213-259  (phessenberg!)
//   src/PeriodicSchurDecompositions.jl:322-1096 (real periodic QR, MB03WD-derived,
//                                                reference-default ALGO_CONFIG :287-302)
// with LAPACK orghr/orgqr semantics for Matrix(H.Q) (:136-140,180).
#pragma once
#include "psdo_common.hpp"

namespace psdo {

// PeriodicSchurDecompositions.jl:229-247.  A = p matrices (col-major, ld=n each) in the
// internal rightwards order.  Reflectors are stored LAPACK-style below the (sub)diagonal,
// tau[j] has n entries (tau[0] uses n-1).
static inline void phessenberg(int n, int p, const std::vector<Mat>& A,
                               std::vector<std::vector<double>>& tau) {
  tau.assign(p, std::vector<double>(n, 0.0));
  for (int i = 1; i <= n - 1; i++) {
    int i1 = i + 1;
    for (int j = p; j >= 2; j--) {
      const Mat& Aj = A[j - 1];
      double* xi = &Aj(i, i);
      double t = reflector(xi, n - i + 1, 1);
      tau[j - 1][i - 1] = t;
      hh_lmul_adj(Aj, i, n - i + 1, i1, n, xi + 1, 1, t);
      hh_rmul(A[j - 2], 1, n, i, n - i + 1, xi + 1, 1, t);
    }
    const Mat& A1 = A[0];
    double* xi = &A1(i1, i);
    double t = reflector(xi, n - i, 1);
    tau[0][i - 1] = t;
    hh_lmul_adj(A1, i1, n - i, i1, n, xi + 1, 1, t);
    hh_rmul(A[p - 1], 1, n, i1, n - i, xi + 1, 1, t);
  }
}

// Matrix(pH[j].Q)  (LAPACK dorg2r): Q = H_1 H_2 ... H_{n-1} from reflectors stored in the
// columns of F below the diagonal (offset=0) or below the subdiagonal (offset=1, dorghr).
static inline void form_q(int n, const Mat& F, const std::vector<double>& tau, int offset,
                          const Mat& Q) {
  for (int j = 1; j <= n; j++)
    for (int i = 1; i <= n; i++) Q(i, j) = (i == j) ? 1.0 : 0.0;
  for (int i = n - 1; i >= 1; i--) {
    int r0 = i + offset;       // row carrying the implicit 1
    int m = n - r0 + 1;        // reflector order
    if (m < 1) continue;
    const double* v = (m > 1) ? &F(r0 + 1, i) : nullptr;
    hh_lmul_adj(Q, r0, m, r0, n, v, 1, tau[i - 1]);  // real: H' == H
  }
}

struct RealQRStats {
  int niter = 0;
  int maxits = 0;
};

// PeriodicSchurDecompositions.jl:322-1096.  H[0] upper Hessenberg, H[1..p-1] upper
// triangular; Z (p matrices) updated in place when wantZ.  wr/wi receive the eigenvalues.
// Returns 0 on success, or the level i at which convergence failed (:891-893).
static inline int real_periodic_qr(int n, int p, const std::vector<Mat>& H,
                                   const std::vector<Mat>& Z, bool wantT, bool wantZ,
                                   int maxitfac, double* lam_re, double* lam_im,
                                   RealQRStats* stats = nullptr) {
  const Mat& H1 = H[0];
  auto Hs = [&](int j) -> const Mat& { return H[j]; };  // Hs[j] == H_{j+1}, j = 1..p-1
  if (n == 1) {  // :333-352
    double l1 = H1(1, 1);
    for (int j = 2; j <= p; j++) l1 *= Hs(j - 1)(1, 1);
    lam_re[0] = l1;
    lam_im[0] = 0.0;
    return 0;
  }
  const double dat1 = 0.75, dat2 = -0.4375;
  std::vector<double> wr(n + 1, 0.0), wi(n + 1, 0.0), hsup(n + 1, 0.0);
  double* hdiag = wr.data();      // 1-based
  double* hsubdiag = wi.data();
  double* hsupdiag = hsup.data();
  std::vector<double> lre(n + 1, 0.0), lim(n + 1, 0.0);
  double v[3];

  const double unfl = DBL_MIN;
  const double ulp = DBL_EPSILON;
  // :366-375 with _AT_pwr16 = 4 : ulpx = ulp * ulp^(4/16) = ulp * sqrt(sqrt(ulp))
  double ulpx = ulp;
  {
    const int AT_hi = 0, AT_lo = 4;
    for (int k = 0; k < AT_hi; k++) ulpx *= ulp;
    double s = ulp;
    for (int iu : {8, 4, 2, 1}) {
      s = std::sqrt(s);
      if (AT_lo & iu) ulpx *= s;
    }
  }
  const double smlnum = unfl * (n / ulp);

  double s = ulp * n;
  if (n > 2)
    for (int r = 3; r <= n; r++) H1(r, 1) = 0.0;
  std::vector<double> hnorms(p + 1, 0.0);
  for (int j = 2; j <= p; j++) {
    for (int r = 2; r <= n; r++) Hs(j - 1)(r, 1) = 0.0;
    hnorms[j] = s * opnorm1(Hs(j - 1), 1, n, 1, n);
  }
  int i1 = 1, i2 = n;
  // _gethess!: triu!(H1, -1)  (:406)
  for (int c = 1; c <= n; c++)
    for (int r = c + 2; r <= n; r++) H1(r, c) = 0.0;
  const Mat& Hp = (p == 1) ? H1 : Hs(p - 1);

  int maxit = maxitfac * n;
  int maxitleft = maxit;
  int i = n;
  int maxits = 0, niter = 0;
  double tst1 = 0.0;

  double h33 = 0, h44 = 0, h43h34 = 0, h43 = 0, h34 = 0;
  double rt1r = 0, rt2r = 0, rt1i = 0, rt2i = 0;

  while (i >= 1) {
    int l = 1;
    int its = 1;
    bool splitting = false;
    double hh21 = 0, hh22 = 0, hh11 = 0, hh12 = 0, hh10 = 0;
    double hp11 = 0, hp12 = 0, hp22 = 0;
    while (its < maxitleft) {
      splitting = false;
      // :474-495
      hp22 = 1.0;
      if (i > l) {
        hp12 = 0.0;
        hp11 = 1.0;
        for (int j = 2; j <= p; j++) {
          const Mat& Hj = Hs(j - 1);
          hp22 *= Hj(i, i);
          hp12 = hp11 * Hj(i - 1, i) + hp12 * Hj(i, i);
          hp11 *= Hj(i - 1, i - 1);
        }
        hh21 = H1(i, i - 1) * hp11;
        hh22 = H1(i, i - 1) * hp12 + H1(i, i) * hp22;
        hdiag[i] = hh22;
        hsubdiag[i] = hh21;
      } else {
        hp22 *= H1(i, i);
        for (int j = 2; j <= p; j++) hp22 *= Hs(j - 1)(i, i);
        hdiag[i] = hp22;
      }
      // :497-576
      int klast = i;
      bool found = false;
      for (int k = i; k >= l + 1; k--) {
        klast = k;
        double hp00 = 1.0, hp01 = 0.0, hp02 = 0.0;
        if (k > l + 1) {
          for (int j = 2; j <= p; j++) {
            const Mat& Hj = Hs(j - 1);
            hp02 = hp00 * Hj(k - 2, k) + hp01 * Hj(k - 1, k) + hp02 * Hj(k, k);
            hp01 = hp00 * Hj(k - 2, k - 1) + hp01 * Hj(k - 1, k - 1);
            hp00 *= Hj(k - 2, k - 2);
          }
          hh10 = H1(k - 1, k - 2) * hp00;
          hh11 = H1(k - 1, k - 2) * hp01 + H1(k - 1, k - 1) * hp11;
          hh12 = H1(k - 1, k - 2) * hp02 + H1(k - 1, k - 1) * hp12 + H1(k - 1, k) * hp22;
          hsubdiag[k - 1] = hh10;
        } else {
          hh10 = 0.0;
          hh11 = H1(k - 1, k - 1) * hp11;
          hh12 = H1(k - 1, k - 1) * hp12 + H1(k - 1, k) * hp22;
        }
        hdiag[k - 1] = hh11;
        hsupdiag[n - i + k - 1] = hh12;

        tst1 = std::fabs(hh11) + std::fabs(hh22);
        if (tst1 == 0.0) tst1 = opnorm1(H1, l, i, l, i);
        // reference default: _slicot_convg = false (:540-565)
        if (std::fabs(hh21) <= smlnum) {
          found = true;
        } else if (std::fabs(hh21) <= ulp * tst1) {
          double ab = std::max(std::fabs(hh21), std::fabs(hh12));
          double ba = std::min(std::fabs(hh21), std::fabs(hh12));
          double aa = std::max(std::fabs(hh22), std::fabs(hh11 - hh22));
          double bb = std::min(std::fabs(hh22), std::fabs(hh11 - hh22));
          double stmp = aa + ab;
          found = ba * (ab / stmp) <= std::max(smlnum, ulpx * (bb * (aa / stmp)));
        }
        if (found) break;
        hp22 = hp11;
        hp11 = hp00;
        hp12 = hp01;
        hh22 = hh11;
        hh21 = hh10;
      }
      // :585
      l = (i > l) ? (found ? klast : l) : i;

      // :589-666
      if (l > 1 && wantT) {
        tst1 = std::fabs(H1(l - 1, l - 1)) + std::fabs(H1(l, l));
        if (tst1 == 0.0) tst1 = opnorm1(H1, l, i, l, i);
        if (std::fabs(H1(l, l - 1)) > std::max(ulp * tst1, smlnum)) {
          for (int k = i; k >= l; k--) {
            for (int j = 1; j <= p - 1; j++) {
              const Mat& Hj = (j == 1) ? H1 : Hs(j - 1);
              double xi[2] = {Hj(k, k), Hj(k, k - 1)};
              double t = reflector(xi, 2);
              Hj(k, k - 1) = 0.0;
              Hj(k, k) = xi[0];
              hh2_rmul(Hj, i1, k - 1, k - 1, xi[1], 1.0, t);
              hh2_lmul_adj(Hs(j), k - 1, k - 1, i2, xi[1], 1.0, t);
              if (wantZ) hh2_rmul(Z[j], 1, n, k - 1, xi[1], 1.0, t);
            }
            if (k < i) {
              double xi[2] = {Hp(k + 1, k + 1), Hp(k + 1, k)};
              double t = reflector(xi, 2);
              Hp(k + 1, k) = 0.0;
              Hp(k + 1, k + 1) = xi[0];
              hh2_rmul(Hp, i1, k, k, xi[1], 1.0, t);
              hh2_lmul_adj(H1, k, k, i2, xi[1], 1.0, t);
              if (wantZ) hh2_rmul(Z[0], 1, n, k, xi[1], 1.0, t);
            }
          }
          // _extra_rq = false (:653-659)
          Hp(l, l - 1) = 0.0;
        }
        H1(l, l - 1) = 0.0;
      }
      if (l >= i - 1) {
        splitting = true;
        break;
      }

      // :675-678
      if (!wantT) {
        i1 = l;
        i2 = i;
      }
      bool exc_shift = false;
      if (its == 10) {
        exc_shift = true;
        s = std::fabs(hsubdiag[l + 1]) + std::fabs(hsubdiag[l + 2]);
        h44 = dat1 * s + hdiag[l];
        h33 = h44;
        h43h34 = dat2 * s * s;
        h43 = s;
        h34 = dat2 * s;
      } else if (its % 10 == 0) {
        exc_shift = true;
        s = std::fabs(hsubdiag[i]) + std::fabs(hsubdiag[i - 1]);
        h44 = dat1 * s + hdiag[i];
        h33 = h44;
        h43h34 = dat2 * s * s;
        h43 = s;
        h34 = dat2 * s;
      } else {
        h44 = hdiag[i];
        h33 = hdiag[i - 1];
        h43h34 = hsubdiag[i] * hsupdiag[n - 1];
        h43 = hsubdiag[i];
        h34 = hsupdiag[n - 1];
        // _slicot_shifts = false: dlahqr-style (:730-762)
        s = std::fabs(h33) + std::fabs(h34) + std::fabs(h43) + std::fabs(h44);
        if (s == 0.0) {
          rt1r = rt2r = rt1i = rt2i = 0.0;
        } else {
          h33 /= s; h44 /= s; h34 /= s; h43 /= s;
          double trc = (h33 + h44) * 0.5;
          double disc = (h33 - trc) * (h44 - trc) - h34 * h43;
          double rtdisc = std::sqrt(std::fabs(disc));
          if (disc >= 0.0) {
            rt1r = trc * s;
            rt2r = rt1r;
            rt1i = rtdisc * s;
            rt2i = -rt1i;
          } else {
            rt1r = trc + rtdisc;
            rt2r = trc - rtdisc;
            rt1r = (std::fabs(rt1r - h44) <= std::fabs(rt2r - h44)) ? (rt1r * s) : (rt2r * s);
            rt2r = rt1r;
            rt1i = rt2i = 0.0;
          }
        }
      }

      // :766-803 with _allow_early_QR = false: mmax = l, loop runs once (m = l)
      int mlast = l;
      {
        int m = l;
        double h11 = hdiag[m];
        double h12 = hsupdiag[n - i + m];
        double h21 = hsubdiag[m + 1];
        double h22 = hdiag[m + 1];
        double v1, v2, v3;
        if (exc_shift) {
          double h44s = h44 - h11;
          double h33s = h33 - h11;
          v1 = (h33s * h44s - h43h34) / h21 + h12;
          v2 = h22 - h11 - h33s - h44s;
          v3 = hsubdiag[m + 2];
        } else {
          s = std::fabs(h11 - rt2r) + std::fabs(rt2i) + std::fabs(h21);
          double h21s = h21 / s;
          v1 = h21s * h12 + (h11 - rt1r) * ((h11 - rt2r) / s) - rt1i * (rt2i / s);
          v2 = h21s * (h11 + h22 - rt1r - rt2r);
          v3 = h21s * hsubdiag[m + 2];
        }
        s = std::fabs(v1) + std::fabs(v2) + std::fabs(v3);
        v[0] = v1 / s;
        v[1] = v2 / s;
        v[2] = v3 / s;
      }

      // :806-886 double-shift sweep
      for (int k = mlast; k <= i - 1; k++) {
        int nr = std::min(3, i - k + 1);
        int nrow = std::min(k + nr, i) - i1 + 1;
        if (k > mlast)
          for (int q = 0; q < nr; q++) v[q] = H1(k + q, k - 1);
        double tau1 = reflector(v, nr);
        if (k > mlast) {
          H1(k, k - 1) = v[0];
          H1(k + 1, k - 1) = 0.0;
          if (k < i - 1) H1(k + 2, k - 1) = 0.0;
        } else if (mlast > l) {
          H1(k, k - 1) = -H1(k, k - 1);
        }
        hh_lmul_adj(H1, k, nr, k, i2, v + 1, 1, tau1);
        hh_rmul(Hp, i1, i1 + nrow - 1, k, nr, v + 1, 1, tau1);
        if (wantZ) hh_rmul(Z[0], 1, n, k, nr, v + 1, 1, tau1);
        for (int j = p; j >= 2; j--) {
          const Mat& Hj = Hs(j - 1);
          for (int q = 0; q < nr; q++) v[q] = Hj(k + q, k);
          double t = reflector(v, nr);
          Hj(k, k) = v[0];
          Hj(k + 1, k) = 0.0;
          if (nr == 3) Hj(k + 2, k) = 0.0;
          hh_lmul_adj(Hj, k, nr, k + 1, i2, v + 1, 1, t);
          const Mat& Hjm1 = (j == 2) ? H1 : Hs(j - 2);
          hh_rmul(Hjm1, i1, i1 + nrow - 1, k, nr, v + 1, 1, t);
          if (wantZ) hh_rmul(Z[j - 1], 1, n, k, nr, v + 1, 1, t);
          if (nr == 3) {
            v[0] = Hj(k + 1, k + 1);
            v[1] = Hj(k + 2, k + 1);
            t = reflector(v, 2);
            Hj(k + 1, k + 1) = v[0];
            Hj(k + 2, k + 1) = 0.0;
            hh_lmul_adj(Hj, k + 1, 2, k + 2, i2, v + 1, 1, t);
            hh_rmul(Hjm1, i1, i1 + nrow - 1, k + 1, 2, v + 1, 1, t);
            if (wantZ) hh_rmul(Z[j - 1], 1, n, k + 1, 2, v + 1, 1, t);
          }
        }
      }
      its++;
    }  // QR iteration loop

    if (!splitting) return i;  // :891-893 "convergence failed at level i"

    // :895-1054 deflation
    if (l == i) {
      lre[i] = hdiag[i];
      lim[i] = 0.0;
    } else if (l == i - 1) {
      if (wantT) {
        hp22 = 1.0; hp12 = 0.0; hp11 = 1.0;
        for (int j = 2; j <= p; j++) {
          const Mat& Hj = Hs(j - 1);
          hp22 *= Hj(i, i);
          hp12 = hp11 * Hj(i - 1, i) + hp12 * Hj(i, i);
          hp11 *= Hj(i - 1, i - 1);
        }
        hh21 = H1(i, i - 1) * hp11;
        hh22 = H1(i, i - 1) * hp12 + H1(i, i) * hp22;
        hh11 = H1(i - 1, i - 1) * hp11;
        hh12 = H1(i - 1, i - 1) * hp12 + H1(i - 1, i) * hp22;
      } else {
        hh11 = hdiag[i - 1];
        hh12 = hsupdiag[n - 1];
        hh21 = hsubdiag[i];
        hh22 = hdiag[i];
      }
      double a = hh11, b = hh12, c = hh21, d = hh22, cs, sn;
      gs2x2(a, b, c, d, cs, sn, lre[i - 1], lim[i - 1], lre[i], lim[i]);
      hdiag[i - 1] = lre[i - 1]; hdiag[i] = lre[i];
      hsubdiag[i - 1] = lim[i - 1]; hsubdiag[i] = lim[i];
      if (wantT) {
        int jmin = 0, jmax = 0;
        for (int j = 2; j <= p; j++) {
          const Mat& Hj = Hs(j - 1);
          if (jmin == 0 && std::fabs(Hj(i - 1, i - 1)) <= hnorms[j]) jmin = j;
          if (std::fabs(Hj(i, i)) <= hnorms[j]) jmax = j;
        }
        if (jmin != 0 && jmax != 0) {
          if (jmin - 1 <= p - jmax + 1) jmax = 0; else jmin = 0;
        }
        if (jmin != 0) {
          // :959-977 ("rarely encountered").  The reference stores xi[2] (the reflector
          // essential part) into Hj[i,i] at :970; the mathematically consistent value is
          // beta = xi[1], which is what we store (do not replicate defects).
          for (int j = 1; j <= jmin - 1; j++) {
            const Mat& Hj = (j == 1) ? H1 : Hs(j - 1);
            double xi[2] = {Hj(i, i), Hj(i, i - 1)};
            double t = reflector(xi, 2);
            Hj(i, i - 1) = 0.0;
            Hj(i, i) = xi[0];
            hh2_rmul(Hj, i1, i - 1, i - 1, xi[1], 1.0, t);
            hh2_lmul_adj(Hs(j), i - 1, i - 1, i2, xi[1], 1.0, t);
            if (wantZ) hh2_rmul(Z[j], 1, n, i - 1, xi[1], 1.0, t);
          }
        } else {
          bool replaceG = (jmax > 0) && (hsubdiag[i - 1] == 0.0);
          double a1 = std::hypot(lre[i - 1], lim[i - 1]);
          double a2 = std::hypot(lre[i], lim[i]);
          bool prodzero = (lre[i] == 0.0 && lim[i] == 0.0) || (lre[i - 1] == 0.0 && lim[i - 1] == 0.0);
          if (prodzero) {
            replaceG = true;
          } else if (hsubdiag[i - 1] == 0.0) {
            if (std::min(a1, a2) / std::max(a1, a2) < DBL_EPSILON) replaceG = true;
          }
          for (int its2 = 1; its2 <= 20; its2++) {
            if (replaceG) {
              double r;
              givens_real(H1(i - 1, i - 1), H1(i, i - 1), cs, sn, r);
            }
            rot_rows(H1, i - 1, i, i - 1, i2, cs, sn);
            rot_cols_adj(Hp, i - 1, i, i1, i, cs, sn);
            if (wantZ) rot_cols_adj(Z[0], i - 1, i, 1, n, cs, sn);
            for (int j = p; j >= std::max(2, jmax + 1); j--) {
              const Mat& Hj = Hs(j - 1);
              v[0] = Hj(i - 1, i - 1);
              v[1] = Hj(i, i - 1);
              double t = reflector(v, 2);
              Hj(i - 1, i - 1) = v[0];
              Hj(i, i - 1) = 0.0;
              hh_lmul_adj(Hj, i - 1, 2, i, i2, v + 1, 1, t);
              const Mat& Hjm1 = (j == 2) ? H1 : Hs(j - 2);
              hh_rmul(Hjm1, i1, i, i - 1, 2, v + 1, 1, t);
              if (wantZ) hh_rmul(Z[j - 1], 1, n, i - 1, 2, v + 1, 1, t);
            }
            if (!replaceG ||
                (std::fabs(H1(i, i - 1)) < std::max(smlnum, ulp * std::max(a1, a2))))
              break;
            replaceG = true;
          }
          if (jmax > 0) {
            H1(i, i - 1) = 0.0;
            if (jmax > 1) Hs(jmax - 1)(i, i - 1) = 0.0;
          } else if (hh21 == 0.0) {
            H1(i, i - 1) = 0.0;
          }
          if (replaceG) {
            // :1039-1051; the reference compares against λ[1],λ[2] (:1048), a slip for
            // λ[i-1],λ[i]; the intended comparison is restated here.
            double l1 = H1(i - 1, i - 1);
            for (int j = 1; j <= p - 1; j++) l1 *= Hs(j)(i - 1, i - 1);
            double d1 = std::hypot(l1 - lre[i - 1], lim[i - 1]);
            double d2 = std::hypot(l1 - lre[i], lim[i]);
            if (d1 > d2) {
              std::swap(lre[i - 1], lre[i]);
              std::swap(lim[i - 1], lim[i]);
            }
          }
        }
      }
    }
    maxitleft -= its;
    i = l - 1;
    maxits = std::max(maxits, its);
    niter += its;
  }
  // :1066-1073
  for (int k = 1; k <= n - 1; k++)
    if (lim[k] == 0.0) H1(k + 1, k) = 0.0;
  for (int k = 1; k <= n; k++) {
    lam_re[k - 1] = lre[k];
    lam_im[k - 1] = lim[k];
  }
  if (stats) {
    stats->niter = niter;
    stats->maxits = maxits;
  }
  return 0;
}

// PeriodicSchurDecompositions.jl:120-152 + :1078-1093: full driver on one problem.
// A: [p][n][n] col-major in the USER's factor order; on return holds the T factors in the
// user's order (T1 at position 1 for :R, p for :L).  Zout: [p][n][n] or nullptr, in the
// reference's result order.  eig: n complex (re,im interleaved).  Returns info.
static inline int rpschur(int n, int p, double* A, double* Zout, double* eig, bool left,
                          bool wantT, bool wantZ, int maxitfac, RealQRStats* stats = nullptr) {
  size_t nn = (size_t)n * n;
  std::vector<Mat> H(p), Z(p);
  for (int j = 1; j <= p; j++) {
    int ju = left ? (p + 1 - j) : j;  // Aarg[j] = A[p+1-j]  (:127-131)
    H[j - 1] = Mat{A + (size_t)(ju - 1) * nn, n};
  }
  std::vector<std::vector<double>> tau;
  phessenberg(n, p, H, tau);
  std::vector<double> Zbuf;
  if (wantZ) {
    Zbuf.assign((size_t)p * nn, 0.0);
    for (int j = 1; j <= p; j++) {
      Z[j - 1] = Mat{Zbuf.data() + (size_t)(j - 1) * nn, n};
      form_q(n, H[j - 1], tau[j - 1], j == 1 ? 1 : 0, Z[j - 1]);
    }
  }
  // Hs = R (triu copies), H1 = triu(H1,-1)  (:147-149)
  for (int j = 1; j <= p; j++) {
    int keep = (j == 1) ? 1 : 0;
    for (int c = 1; c <= n; c++)
      for (int r = c + keep + 1; r <= n; r++) H[j - 1](r, c) = 0.0;
  }
  std::vector<double> lre(n), lim(n);
  int info = real_periodic_qr(n, p, H, Z, wantT, wantZ, maxitfac, lre.data(), lim.data(), stats);
  for (int k = 0; k < n; k++) {
    eig[2 * k] = lre[k];
    eig[2 * k + 1] = lim[k];
  }
  if (wantZ && Zout) {
    for (int l = 1; l <= p; l++) {
      int src = (!left || l == 1) ? l : (p + 2 - l);  // Zr[l] = Z[p+2-l]  (:1081-1084)
      std::memcpy(Zout + (size_t)(l - 1) * nn, Zbuf.data() + (size_t)(src - 1) * nn,
                  nn * sizeof(double));
    }
  }
  return info;
}

}  // namespace psdo
