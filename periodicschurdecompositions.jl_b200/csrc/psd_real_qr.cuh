// Real periodic QR iteration of one problem by one CTA (device functions only): the context
// struct and periodic_qr_cta, shared by the batched small-N kernel (psd_real_kernel.cuh) and by
// the shift / final-block kernels of the large-N multishift path (psd_ms_kernels.cuh).
// Reference: pschur!(H1, Hs; ...) PeriodicSchurDecompositions.jl:322-1096, _gs2x2! rschur2x2.jl:9-96.
#pragma once
#include "psd_device.cuh"

namespace psd {

// number of doubles of "small" per-problem state
__host__ __device__ inline long long rp_small_doubles(int n, int p) {
  return 8LL * (n + 2) + (p + 2);
}

struct RCtx {
  int n, p, tid, nt;
  double* H;     // H(j) = H + (j-1)*hs, leading dimension ldh
  long long hs;
  int ldh;
  double* Z;     // Z(j) = Z + zoff(j), leading dimension ldz
  long long zs;
  int ldz;
  bool zmap_left;  // global-mode: internal Z index -> reference result index for :L
  double *hdiag, *hsub, *hsup, *t0, *t1, *t2, *lre, *lim, *hnorms;
  PSD_DEV double* Hp(int j) const { return H + (long long)(j - 1) * hs; }
  PSD_DEV double* Zp(int j) const {
    int s = j;
    if (zmap_left && j > 1) s = p + 2 - j;  // Zr[l] = Z[p+2-l]  (:1081-1084)
    return Z + (long long)(s - 1) * zs;
  }
};

#define PSD_EL(ptr, ld, r, c) (ptr)[((r)-1) + (size_t)((c)-1) * (ld)]

PSD_DEV void hh_apply_n(int nr, int tid, int nt, double* L, int ldl, int r, int cl0, int cl1,
                        double* R, int ldr, int rr0, int rr1, int rc, double* Zm, int ldz, int nz,
                        int zc, double v1, double v2, double tau) {
  if (nr == 3)
    hh_apply<3>(tid, nt, L, ldl, r, cl0, cl1, R, ldr, rr0, rr1, rc, Zm, ldz, nz, zc, v1, v2, tau);
  else
    hh_apply<2>(tid, nt, L, ldl, r, cl0, cl1, R, ldr, rr0, rr1, rc, Zm, ldz, nz, zc, v1, v2, tau);
}

// opnorm(view(H, r0:r1, c0:c1), 1): rare fallback, computed serially by every thread.
PSD_DEV double opnorm1_serial(const double* Hm, int ld, int r0, int r1, int c0, int c1) {
  double m = 0.0;
  for (int c = c0; c <= c1; c++) {
    double s = 0.0;
    for (int r = r0; r <= r1; r++) s += fabs(PSD_EL(Hm, ld, r, c));
    m = fmax(m, s);
  }
  return m;
}

// ---------------------------------------------------------------------------------------
// Real periodic QR iteration, PeriodicSchurDecompositions.jl:322-1096.
// Returns info (0, or the level i at which convergence failed, :891-893).
// ---------------------------------------------------------------------------------------
PSD_DEV int periodic_qr_cta(const RCtx& c, bool wantT, bool wantZ, int maxitfac, int* niter_out) {
  const int n = c.n, p = c.p, ld = c.ldh, tid = c.tid, nt = c.nt;
  double* H1 = c.Hp(1);
  double* Hpp = c.Hp(p);
  double *hdiag = c.hdiag, *hsub = c.hsub, *hsup = c.hsup;
  double *lre = c.lre, *lim = c.lim;
#define H1_(r, cc) PSD_EL(H1, ld, r, cc)
#define HJ_(j, r, cc) PSD_EL(c.Hp(j), ld, r, cc)

  if (n == 1) {  // :333-352
    if (tid == 0) {
      double l1 = H1_(1, 1);
      for (int j = 2; j <= p; j++) l1 *= HJ_(j, 1, 1);
      lre[1] = l1;
      lim[1] = 0.0;
    }
    __syncthreads();
    *niter_out = 0;
    return 0;
  }

  const double dat1 = 0.75, dat2 = -0.4375;
  const double ulp = DBL_EPSILON;
  const double ulpx = ulp * sqrt(sqrt(ulp));  // :366-375, _AT_pwr16 = 4
  const double smlnum = DBL_MIN * ((double)n / ulp);

  // hnorms[j] = eps*n*opnorm(Hs[j-1],1)  (:384-388): one thread per column + max-reduce
  for (int j = 2 + tid; j <= p; j += nt) c.hnorms[j] = 0.0;
  for (int e = tid; e <= n; e += nt) {
    lre[e] = 0.0;
    lim[e] = 0.0;
  }
  __syncthreads();
  if (wantT) {
    // only used by the wantT deflation branch (:937-949)
    for (int w = tid; w < (p - 1) * n; w += nt) {
      int j = 2 + w / n, col = 1 + w % n;
      const double* Hj = c.Hp(j);
      double s = 0.0;
      for (int r = 1; r <= col; r++) s += fabs(PSD_EL(Hj, ld, r, col));
      // non-negative doubles order like their bit patterns
      atomicMax((unsigned long long*)&c.hnorms[j], (unsigned long long)__double_as_longlong(s));
    }
    __syncthreads();
    for (int j = 2 + tid; j <= p; j += nt) c.hnorms[j] *= ulp * n;
    __syncthreads();
  }

  int i1 = 1, i2 = n;
  int maxitleft = maxitfac * n;
  int i = n;
  int niter = 0;
  double v0 = 0, v1 = 0, v2 = 0;

  while (i >= 1) {
    int l = 1;
    int its = 1;
    bool splitting = false;
    while (its < maxitleft) {
      // ---- product band (:474-529), one thread per row r in l..i ----
      for (int r = l + tid; r <= i; r += nt) {
        double q0 = 1.0, q1 = 0.0, q2 = 0.0;
        const bool h1b = (r + 1 <= i), h2b = (r + 2 <= i);
        for (int j = 2; j <= p; j++) {
          const double* Hj = c.Hp(j);
          if (h2b)
            q2 = q0 * PSD_EL(Hj, ld, r, r + 2) + q1 * PSD_EL(Hj, ld, r + 1, r + 2) +
                 q2 * PSD_EL(Hj, ld, r + 2, r + 2);
          if (h1b) q1 = q0 * PSD_EL(Hj, ld, r, r + 1) + q1 * PSD_EL(Hj, ld, r + 1, r + 1);
          q0 *= PSD_EL(Hj, ld, r, r);
        }
        c.t0[r] = q0;
        c.t1[r] = q1;
        c.t2[r] = q2;
      }
      __syncthreads();
      for (int r = l + tid; r <= i; r += nt) {
        if (r > l) {
          const double hs = H1_(r, r - 1);
          hsub[r] = hs * c.t0[r - 1];
          hdiag[r] = hs * c.t1[r - 1] + H1_(r, r) * c.t0[r];
          if (r < i) hsup[r] = hs * c.t2[r - 1] + H1_(r, r) * c.t1[r] + H1_(r, r + 1) * c.t0[r + 1];
        } else {
          hdiag[r] = H1_(r, r) * c.t0[r];
          if (r < i) hsup[r] = H1_(r, r) * c.t1[r] + H1_(r, r + 1) * c.t0[r + 1];
        }
      }
      __syncthreads();
      // ---- search for a negligible subdiagonal of the product (:497-585) ----
      // every thread scans redundantly from the bottom (values are in shared/L1 memory).
      int lnew = l;
      if (i > l) {
        for (int k = i; k >= l + 1; k--) {
          const double hh21 = hsub[k], hh22 = hdiag[k], hh11 = hdiag[k - 1], hh12 = hsup[k - 1];
          bool found = false;
          if (fabs(hh21) <= smlnum) {
            found = true;
          } else {
            double tst1 = fabs(hh11) + fabs(hh22);
            if (tst1 == 0.0) tst1 = opnorm1_serial(H1, ld, l, i, l, i);
            if (fabs(hh21) <= ulp * tst1) {
              double ab = fmax(fabs(hh21), fabs(hh12));
              double ba = fmin(fabs(hh21), fabs(hh12));
              double aa = fmax(fabs(hh22), fabs(hh11 - hh22));
              double bb = fmin(fabs(hh22), fabs(hh11 - hh22));
              double st = aa + ab;
              found = ba * (ab / st) <= fmax(smlnum, ulpx * (bb * (aa / st)));
            }
          }
          if (found) {
            lnew = k;
            break;
          }
        }
        l = lnew;
      } else {
        l = i;
      }

      // ---- RQ step when the product subdiagonal is small but H1[l,l-1] is not (:589-666)
      if (l > 1 && wantT) {
        double tst1 = fabs(H1_(l - 1, l - 1)) + fabs(H1_(l, l));
        if (tst1 == 0.0) tst1 = opnorm1_serial(H1, ld, l, i, l, i);
        const bool dorq = (p > 1) && fabs(H1_(l, l - 1)) > fmax(ulp * tst1, smlnum);
        __syncthreads();  // all reads above done before anyone writes
        if (dorq) {
          for (int k = i; k >= l; k--) {
            for (int j = 1; j <= p - 1; j++) {
              double* Hj = c.Hp(j);
              double x0 = PSD_EL(Hj, ld, k, k), w1 = PSD_EL(Hj, ld, k, k - 1), dum = 0.0;
              __syncthreads();
              double t = refl_small<2>(x0, w1, dum);
              if (tid == 0) {
                PSD_EL(Hj, ld, k, k - 1) = 0.0;
                PSD_EL(Hj, ld, k, k) = x0;
              }
              hh2_apply(tid, nt, c.Hp(j + 1), ld, k - 1, k - 1, i2, Hj, ld, i1, k - 1, k - 1,
                        wantZ ? c.Zp(j + 1) : nullptr, c.ldz, n, k - 1, w1, 1.0, t);
              __syncthreads();
            }
            if (k < i) {
              double x0 = PSD_EL(Hpp, ld, k + 1, k + 1), w1 = PSD_EL(Hpp, ld, k + 1, k), dum = 0.0;
              __syncthreads();
              double t = refl_small<2>(x0, w1, dum);
              if (tid == 0) {
                PSD_EL(Hpp, ld, k + 1, k) = 0.0;
                PSD_EL(Hpp, ld, k + 1, k + 1) = x0;
              }
              hh2_apply(tid, nt, H1, ld, k, k, i2, Hpp, ld, i1, k, k,
                        wantZ ? c.Zp(1) : nullptr, c.ldz, n, k, w1, 1.0, t);
              __syncthreads();
            }
          }
          if (tid == 0) PSD_EL(Hpp, ld, l, l - 1) = 0.0;  // _extra_rq = false (:653-659)
        }
        if (tid == 0) H1_(l, l - 1) = 0.0;
        __syncthreads();
      }
      if (l >= i - 1) {
        splitting = true;
        break;
      }

      if (!wantT) {
        i1 = l;
        i2 = i;
      }
      // ---- shifts (:679-764) and first column of the shift polynomial (:766-803) ----
      {
        const int m = l;
        const double h11 = hdiag[m], h12 = hsup[m], h21 = hsub[m + 1], h22 = hdiag[m + 1];
        const double hs3 = hsub[m + 2];
        double s;
        if (its == 10 || its % 10 == 0) {
          if (its == 10)
            s = fabs(hsub[l + 1]) + fabs(hsub[l + 2]);
          else
            s = fabs(hsub[i]) + fabs(hsub[i - 1]);
          const double h44 = dat1 * s + ((its == 10) ? hdiag[l] : hdiag[i]);
          const double h33 = h44;
          const double h43h34 = dat2 * s * s;
          const double h44s = h44 - h11, h33s = h33 - h11;
          v0 = (h33s * h44s - h43h34) / h21 + h12;
          v1 = h22 - h11 - h33s - h44s;
          v2 = hs3;
        } else {
          double h44 = hdiag[i], h33 = hdiag[i - 1], h43 = hsub[i], h34 = hsup[i - 1];
          double rt1r, rt2r, rt1i, rt2i;
          s = fabs(h33) + fabs(h34) + fabs(h43) + fabs(h44);
          if (s == 0.0) {
            rt1r = rt2r = rt1i = rt2i = 0.0;
          } else {
            h33 /= s; h44 /= s; h34 /= s; h43 /= s;
            const double trc = (h33 + h44) * 0.5;
            const double disc = (h33 - trc) * (h44 - trc) - h34 * h43;
            const double rtdisc = sqrt(fabs(disc));
            if (disc >= 0.0) {
              rt1r = trc * s; rt2r = rt1r; rt1i = rtdisc * s; rt2i = -rt1i;
            } else {
              rt1r = trc + rtdisc;
              rt2r = trc - rtdisc;
              rt1r = (fabs(rt1r - h44) <= fabs(rt2r - h44)) ? (rt1r * s) : (rt2r * s);
              rt2r = rt1r;
              rt1i = rt2i = 0.0;
            }
          }
          s = fabs(h11 - rt2r) + fabs(rt2i) + fabs(h21);
          const double h21s = h21 / s;
          v0 = h21s * h12 + (h11 - rt1r) * ((h11 - rt2r) / s) - rt1i * (rt2i / s);
          v1 = h21s * (h11 + h22 - rt1r - rt2r);
          v2 = h21s * hs3;
        }
        s = fabs(v0) + fabs(v1) + fabs(v2);
        v0 /= s; v1 /= s; v2 /= s;
      }

      // ---- double-shift sweep (:806-886) ----
      for (int k = l; k <= i - 1; k++) {
        const int nr = min(3, i - k + 1);
        const int rlast = min(k + nr, i);  // last row touched by column operations
        double x0, w1, w2;
        if (k > l) {
          x0 = H1_(k, k - 1);
          w1 = H1_(k + 1, k - 1);
          w2 = (nr == 3) ? H1_(k + 2, k - 1) : 0.0;
        } else {
          x0 = v0; w1 = v1; w2 = (nr == 3) ? v2 : 0.0;
        }
        __syncthreads();  // reads of the reflector source complete before it is overwritten
        double tau = (nr == 3) ? refl_small<3>(x0, w1, w2) : refl_small<2>(x0, w1, w2);
        if (k > l && tid < nr) H1_(k + tid, k - 1) = (tid == 0) ? x0 : 0.0;
        if (p > 1) {
          hh_apply_n(nr, tid, nt, H1, ld, k, k, i2, Hpp, ld, i1, rlast, k,
                     wantZ ? c.Zp(1) : nullptr, c.ldz, n, k, w1, w2, tau);
        } else {  // left and right targets coincide: serialise (reference order: left first)
          hh_apply_n(nr, tid, nt, H1, ld, k, k, i2, nullptr, ld, 1, 0, k, nullptr, c.ldz, n, k, w1,
                     w2, tau);
          __syncthreads();
          hh_apply_n(nr, tid, nt, nullptr, ld, k, 1, 0, H1, ld, i1, rlast, k,
                     wantZ ? c.Zp(1) : nullptr, c.ldz, n, k, w1, w2, tau);
        }
        __syncthreads();
        for (int j = p; j >= 2; j--) {
          double* Hj = c.Hp(j);
          double* Hjm1 = c.Hp(j - 1);
          double* Zj = wantZ ? c.Zp(j) : nullptr;
          x0 = PSD_EL(Hj, ld, k, k);
          w1 = PSD_EL(Hj, ld, k + 1, k);
          w2 = (nr == 3) ? PSD_EL(Hj, ld, k + 2, k) : 0.0;
          __syncthreads();
          tau = (nr == 3) ? refl_small<3>(x0, w1, w2) : refl_small<2>(x0, w1, w2);
          if (tid < nr) PSD_EL(Hj, ld, k + tid, k) = (tid == 0) ? x0 : 0.0;
          hh_apply_n(nr, tid, nt, Hj, ld, k, k + 1, i2, Hjm1, ld, i1, rlast, k, Zj, c.ldz, n, k, w1,
                     w2, tau);
          __syncthreads();
          if (nr == 3) {
            x0 = PSD_EL(Hj, ld, k + 1, k + 1);
            w1 = PSD_EL(Hj, ld, k + 2, k + 1);
            w2 = 0.0;
            __syncthreads();
            tau = refl_small<2>(x0, w1, w2);
            if (tid < 2) PSD_EL(Hj, ld, k + 1 + tid, k + 1) = (tid == 0) ? x0 : 0.0;
            hh_apply<2>(tid, nt, Hj, ld, k + 1, k + 2, i2, Hjm1, ld, i1, rlast, k + 1, Zj, c.ldz, n,
                        k + 1, w1, w2, tau);
            __syncthreads();
          }
        }
      }
      its++;
    }  // QR iterations

    if (!splitting) {
      *niter_out = niter + its;
      return i;  // :891-893
    }

    // ---- deflation (:895-1054) ----
    if (l == i) {
      if (tid == 0) {
        lre[i] = hdiag[i];
        lim[i] = 0.0;
      }
    } else {  // l == i-1
      double hh11, hh12, hh21, hh22;
      if (wantT) {
        double hp22 = 1.0, hp12 = 0.0, hp11 = 1.0;
        for (int j = 2; j <= p; j++) {
          const double* Hj = c.Hp(j);
          hp22 *= PSD_EL(Hj, ld, i, i);
          hp12 = hp11 * PSD_EL(Hj, ld, i - 1, i) + hp12 * PSD_EL(Hj, ld, i, i);
          hp11 *= PSD_EL(Hj, ld, i - 1, i - 1);
        }
        hh21 = H1_(i, i - 1) * hp11;
        hh22 = H1_(i, i - 1) * hp12 + H1_(i, i) * hp22;
        hh11 = H1_(i - 1, i - 1) * hp11;
        hh12 = H1_(i - 1, i - 1) * hp12 + H1_(i - 1, i) * hp22;
      } else {
        hh11 = hdiag[i - 1]; hh12 = hsup[i - 1]; hh21 = hsub[i]; hh22 = hdiag[i];
      }
      double a = hh11, b = hh12, cc = hh21, d = hh22, cs, sn, l1r, l1i, l2r, l2i;
      gs2x2(a, b, cc, d, cs, sn, l1r, l1i, l2r, l2i);
      if (wantT) {
        int jmin = 0, jmax = 0;
        for (int j = 2; j <= p; j++) {
          const double* Hj = c.Hp(j);
          if (jmin == 0 && fabs(PSD_EL(Hj, ld, i - 1, i - 1)) <= c.hnorms[j]) jmin = j;
          if (fabs(PSD_EL(Hj, ld, i, i)) <= c.hnorms[j]) jmax = j;
        }
        if (jmin != 0 && jmax != 0) {
          if (jmin - 1 <= p - jmax + 1) jmax = 0; else jmin = 0;
        }
        __syncthreads();
        if (jmin != 0) {
          // :959-977 (beta stored at Hj[i,i]; see oracle/psdo_real.hpp for the note on :970)
          for (int j = 1; j <= jmin - 1; j++) {
            double* Hj = c.Hp(j);
            double x0 = PSD_EL(Hj, ld, i, i), w1 = PSD_EL(Hj, ld, i, i - 1), dum = 0.0;
            __syncthreads();
            double t = refl_small<2>(x0, w1, dum);
            if (tid == 0) {
              PSD_EL(Hj, ld, i, i - 1) = 0.0;
              PSD_EL(Hj, ld, i, i) = x0;
            }
            hh2_apply(tid, nt, c.Hp(j + 1), ld, i - 1, i - 1, i2, Hj, ld, i1, i - 1, i - 1,
                      wantZ ? c.Zp(j + 1) : nullptr, c.ldz, n, i - 1, w1, 1.0, t);
            __syncthreads();
          }
        } else {
          bool replaceG = (jmax > 0) && (l1i == 0.0);
          const double a1 = hypot(l1r, l1i), a2 = hypot(l2r, l2i);
          if (a1 == 0.0 || a2 == 0.0) {
            replaceG = true;
          } else if (l1i == 0.0) {
            if (fmin(a1, a2) / fmax(a1, a2) < DBL_EPSILON) replaceG = true;
          }
          for (int its2 = 1; its2 <= 20; its2++) {
            if (replaceG) {
              double rr;
              givens_real(H1_(i - 1, i - 1), H1_(i, i - 1), cs, sn, rr);
            }
            __syncthreads();
            if (p > 1) {
              rot_apply(tid, nt, H1, ld, i - 1, i - 1, i2, Hpp, ld, i1, i, i - 1,
                        wantZ ? c.Zp(1) : nullptr, c.ldz, n, i - 1, cs, sn);
            } else {
              rot_apply(tid, nt, H1, ld, i - 1, i - 1, i2, nullptr, ld, 1, 0, i - 1, nullptr, c.ldz,
                        n, i - 1, cs, sn);
              __syncthreads();
              rot_apply(tid, nt, nullptr, ld, i - 1, 1, 0, H1, ld, i1, i, i - 1,
                        wantZ ? c.Zp(1) : nullptr, c.ldz, n, i - 1, cs, sn);
            }
            __syncthreads();
            for (int j = p; j >= max(2, jmax + 1); j--) {
              double* Hj = c.Hp(j);
              double x0 = PSD_EL(Hj, ld, i - 1, i - 1), w1 = PSD_EL(Hj, ld, i, i - 1), w2 = 0.0;
              __syncthreads();
              double t = refl_small<2>(x0, w1, w2);
              if (tid == 0) {
                PSD_EL(Hj, ld, i - 1, i - 1) = x0;
                PSD_EL(Hj, ld, i, i - 1) = 0.0;
              }
              hh_apply<2>(tid, nt, Hj, ld, i - 1, i, i2, c.Hp(j - 1), ld, i1, i, i - 1,
                          wantZ ? c.Zp(j) : nullptr, c.ldz, n, i - 1, w1, w2, t);
              __syncthreads();
            }
            if (!replaceG || (fabs(H1_(i, i - 1)) < fmax(smlnum, ulp * fmax(a1, a2)))) break;
            replaceG = true;
          }
          __syncthreads();
          if (tid == 0) {
            if (jmax > 0) {
              H1_(i, i - 1) = 0.0;
              if (jmax > 1) HJ_(jmax, i, i - 1) = 0.0;
            } else if (hh21 == 0.0) {
              H1_(i, i - 1) = 0.0;
            }
          }
          if (replaceG) {
            // eigenvalue order may have been swapped by the rotation (:1039-1051)
            double q1 = H1_(i - 1, i - 1);
            for (int j = 2; j <= p; j++) q1 *= HJ_(j, i - 1, i - 1);
            if (hypot(q1 - l1r, l1i) > hypot(q1 - l2r, l2i)) {
              double t;
              t = l1r; l1r = l2r; l2r = t;
              t = l1i; l1i = l2i; l2i = t;
            }
          }
        }
      }
      if (tid == 0) {
        lre[i - 1] = l1r; lim[i - 1] = l1i;
        lre[i] = l2r; lim[i] = l2i;
      }
    }
    __syncthreads();
    maxitleft -= its;
    i = l - 1;
    niter += its;
  }
  // :1066-1073
  for (int k = 1 + tid; k <= n - 1; k += nt)
    if (lim[k] == 0.0) H1_(k + 1, k) = 0.0;
  __syncthreads();
  *niter_out = niter;
  return 0;
#undef H1_
#undef HJ_
}

}  // namespace psd
