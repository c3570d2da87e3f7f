// Real standard periodic Schur kernel: one CTA per periodic problem.
//
// Replaces, on device, the reference call chain
//   pschur!(A, lr)                    PeriodicSchurDecompositions.jl:120-152
//   phessenberg!(A) + Matrix(H.Q)     PeriodicSchurDecompositions.jl:213-259, 136-140
//   pschur!(H1, Hs; ...)              PeriodicSchurDecompositions.jl:322-1096
//   _gs2x2!                           rschur2x2.jl:9-96
// with the reference-default ALGO_CONFIG switches (:287-302): LAPACK-style shifts,
// LAPACK + Ahues-Tisseur convergence test with eps^(1+4/16), no early QR start, no extra RQ.
//
// Data layout: the p factors (and, when requested, the p Schur-vector matrices) of the
// problem live either in shared memory (leading dimension ldh, odd, so that both row-wise
// and column-wise warp accesses are bank-conflict free) or, when they do not fit, in place
// in global memory (L2 resident); the code below is written on generic pointers and is
// identical for both.  Factor index j is always the INTERNAL rightwards index: for the :L
// orientation internal j maps to user factor p+1-j (:127-131).
#pragma once
#include "psd_device.cuh"
#include "psd_real_eig32.cuh"

namespace psd {

struct RpschurParams {
  int n, p;
  long long batch;
  int left, wantT, wantZ, maxitfac;
  double* A;        // [batch][p][n*n]  in/out (user order)
  double* Z;        // [batch][p][n*n]  out (reference result order) or nullptr
  double* eig;      // [batch][n][2]
  int* info;        // [batch]
  int* iters;       // [batch] or nullptr: total QR iterations (reference `niter`)
  int use_smem;     // factors staged in shared memory?
  int ldh;          // leading dimension of staged matrices
  int reduce_only;  // 1: stop after the Hessenberg-triangular reduction (psd_rphess_batched)
  int skip_reduce;  // 1: input is already Hessenberg/triangular (pschur!(H1,Hs) entry, :322)
  int z_preset;     // with skip_reduce: Z already holds the Q_j of an earlier reduction (Q kwarg, :326)
  unsigned long long* counter;  // dynamic work queue over problems
  double* scratch;  // per-CTA small arrays when they do not fit in smem (or nullptr)
  long long scratch_stride;
  double* packed_out;  // reduce_only: write packed Hessenberg-triangular factors here instead
                       // of A ([batch][pk_problem_size(n,p)], layout of psd_real_eig32.cuh)
};

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
// Periodic Hessenberg-triangular reduction, PeriodicSchurDecompositions.jl:229-247, with
// the Schur-vector accumulation fused in: instead of storing LAPACK-packed reflectors and
// materialising Q afterwards (orghr/orgqr, :136-140) each reflector is applied to Z[j]
// from the right as it is generated (Q_j = H_1 H_2 ... H_{n-1}).
// Every warp recomputes the column norm redundantly so no broadcast step is needed; one CTA
// barrier per reflector.
// ---------------------------------------------------------------------------------------
// mode 0: left + right + Z, then finalise the column; mode 1: left only (no finalise);
// mode 2: right + Z only, then finalise.  Modes 1,2 serialise the two sides when the left
// and right targets are the same matrix (p == 1).
PSD_DEV void reduce_step(const RCtx& c, double* Aj, double* Ajm1, double* Zj, int r0, int col,
                         int mode = 0) {
  const int n = c.n, ld = c.ldh;
  const int m = n - r0 + 1;  // reflector order
  if (m <= 1) return;
  const int lane = c.tid & 31;
  const double* x = &PSD_EL(Aj, ld, r0, col);
  const double alpha = x[0];
  double amax = 0.0;
  for (int k = 1 + lane; k < m; k += 32) amax = fmax(amax, fabs(x[k]));
  amax = warp_max(amax);
  if (amax == 0.0) {  // tau = 0, H = I  (householder.jl:76-78)
    __syncthreads();
    return;
  }
  double mm = fmax(amax, fabs(alpha));
  double s = 1.0;
  if (mm < 1e-140 || mm > 1e140) s = pow2_rescale(mm);
  double ssq = 0.0;
  for (int k = 1 + lane; k < m; k += 32) {
    double y = x[k] * s;
    ssq = fma(y, y, ssq);
  }
  ssq = warp_sum(ssq);
  const double al = alpha * s;
  const double beta = -copysign(sqrt(fma(al, al, ssq)), al);
  const double tau = (beta - al) / beta;
  const double tv = s / (al - beta);  // v_k = x_k * tv
  // H = I - tau w w^T, w = (1, tv*x[1:]).
  const int nL = (mode == 2) ? 0 : (n - col);  // columns col+1..n of Aj
  const int nR = (mode == 1) ? 0 : n;          // rows of A_{j-1}
  const int nZ = (Zj && mode != 1) ? n : 0;
  const int tot = nL + nR + nZ;
  for (int w = c.tid; w < tot; w += c.nt) {
    if (w < nL) {
      double* a = &PSD_EL(Aj, ld, r0, col + 1 + w);
      double d = 0.0;
      for (int k = 1; k < m; k++) d = fma(x[k], a[k], d);
      d = tau * fma(d, tv, a[0]);
      a[0] -= d;
      const double g = d * tv;
      for (int k = 1; k < m; k++) a[k] = fma(-g, x[k], a[k]);
    } else {
      double* a;
      int lda;
      if (w < nL + nR) {
        a = &PSD_EL(Ajm1, ld, 1 + (w - nL), r0);
        lda = ld;
      } else {
        a = &PSD_EL(Zj, c.ldz, 1 + (w - nL - nR), r0);
        lda = c.ldz;
      }
      double d = 0.0;
      for (int k = 1; k < m; k++) d = fma(a[(size_t)k * lda], x[k], d);
      d = tau * fma(d, tv, a[0]);
      a[0] -= d;
      const double g = d * tv;
      for (int k = 1; k < m; k++) a[(size_t)k * lda] = fma(-g, x[k], a[(size_t)k * lda]);
    }
  }
  __syncthreads();
  if (mode == 1) return;
  // column `col` of Aj below r0 is dead from here on: store beta and exact zeros.
  for (int k = c.tid; k < m; k += c.nt) PSD_EL(Aj, ld, r0 + k, col) = (k == 0) ? beta / s : 0.0;
}

PSD_DEV void phessenberg_cta(const RCtx& c, bool wantZ) {
  const int n = c.n, p = c.p;
  if (wantZ) {
    for (int j = 1; j <= p; j++) {
      double* Zj = c.Zp(j);
      for (int e = c.tid; e < n * n; e += c.nt) {
        int r = e % n, cc = e / n;
        Zj[r + (size_t)cc * c.ldz] = (r == cc) ? 1.0 : 0.0;
      }
    }
  }
  __syncthreads();
  for (int i = 1; i <= n - 1; i++) {
    for (int j = p; j >= 2; j--)
      reduce_step(c, c.Hp(j), c.Hp(j - 1), wantZ ? c.Zp(j) : nullptr, i, i);
    if (p > 1) {
      reduce_step(c, c.Hp(1), c.Hp(p), wantZ ? c.Zp(1) : nullptr, i + 1, i);
    } else {
      reduce_step(c, c.Hp(1), c.Hp(1), wantZ ? c.Zp(1) : nullptr, i + 1, i, 1);
      reduce_step(c, c.Hp(1), c.Hp(1), wantZ ? c.Zp(1) : nullptr, i + 1, i, 2);
    }
  }
  __syncthreads();
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

// ---------------------------------------------------------------------------------------
// Kernel: persistent CTAs pull problems from a global counter (convergence-dependent run
// time per problem), stage the factors, reduce, iterate, and write results back.
// ---------------------------------------------------------------------------------------
extern __shared__ __align__(16) double psd_smem[];

__global__ void rpschur_kernel(RpschurParams P) {
  const int n = P.n, p = P.p, tid = threadIdx.x, nt = blockDim.x;
  const size_t nn = (size_t)n * n;
  __shared__ long long s_b;

  double* small = P.scratch ? (P.scratch + (long long)blockIdx.x * P.scratch_stride) : psd_smem;
  double* mats = P.scratch ? psd_smem : (psd_smem + rp_small_doubles(n, p));

  RCtx c;
  c.n = n; c.p = p; c.tid = tid; c.nt = nt;
  c.hdiag = small;
  c.hsub = c.hdiag + (n + 2);
  c.hsup = c.hsub + (n + 2);
  c.t0 = c.hsup + (n + 2);
  c.t1 = c.t0 + (n + 2);
  c.t2 = c.t1 + (n + 2);
  c.lre = c.t2 + (n + 2);
  c.lim = c.lre + (n + 2);
  c.hnorms = c.lim + (n + 2);

  const bool wantZ = P.wantZ && P.Z;
  const bool left = P.left != 0;

  for (;;) {
    if (tid == 0) s_b = (long long)atomicAdd(P.counter, 1ULL);
    __syncthreads();
    const long long b = s_b;
    __syncthreads();
    if (b >= P.batch) break;

    double* Ab = P.A + (size_t)b * p * nn;
    double* Zb = wantZ ? (P.Z + (size_t)b * p * nn) : nullptr;

    if (P.use_smem) {
      c.ldh = P.ldh; c.ldz = P.ldh;
      c.H = mats; c.hs = (long long)P.ldh * n;
      c.Z = mats + (size_t)p * P.ldh * n; c.zs = (long long)P.ldh * n;
      c.zmap_left = false;
      // stage: internal factor j <- user factor (left ? p+1-j : j)
      for (int j = 1; j <= p; j++) {
        const double* src = Ab + (size_t)((left ? (p + 1 - j) : j) - 1) * nn;
        double* dst = c.Hp(j);
        for (int e = tid; e < (int)nn; e += nt) {
          int r = e % n, cc = e / n;
          dst[r + (size_t)cc * c.ldh] = src[e];
        }
      }
    } else {
      c.ldh = n; c.ldz = n;
      if (left) {
        c.H = Ab + (size_t)(p - 1) * nn; c.hs = -(long long)nn;
      } else {
        c.H = Ab; c.hs = (long long)nn;
      }
      c.Z = Zb; c.zs = (long long)nn;
      c.zmap_left = left;
    }
    __syncthreads();

    if (!P.skip_reduce) {
      phessenberg_cta(c, wantZ);
    } else {
      if (wantZ && !P.z_preset) {
        for (int j = 1; j <= p; j++) {
          double* Zj = c.Zp(j);
          for (int e = tid; e < (int)nn; e += nt) {
            int r = e % n, cc = e / n;
            Zj[r + (size_t)cc * c.ldz] = (r == cc) ? 1.0 : 0.0;
          }
        }
      }
      // enforce structure: :380-387, :406
      for (int j = 1; j <= p; j++) {
        double* Hj = c.Hp(j);
        const int keep = (j == 1) ? 1 : 0;
        for (int e = tid; e < (int)nn; e += nt) {
          int r = e % n, cc = e / n;
          if (r > cc + keep) Hj[r + (size_t)cc * c.ldh] = 0.0;
        }
      }
      __syncthreads();
    }

    int info = 0, niter = 0;
    if (!P.reduce_only) info = periodic_qr_cta(c, P.wantT != 0, wantZ, P.maxitfac, &niter);

    // write back
    if (!P.reduce_only) {
      double* eg = P.eig + (size_t)b * 2 * n;
      for (int k = tid; k < n; k += nt) {
        eg[2 * k] = c.lre[k + 1];
        eg[2 * k + 1] = c.lim[k + 1];
      }
      if (tid == 0) {
        P.info[b] = info;
        if (P.iters) P.iters[b] = niter;
      }
    }
    if (P.packed_out) {
      // packed factors, each normalised by an exact power of two (see pk_problem_stride)
      double* dstb = P.packed_out + (size_t)b * pk_problem_stride(n, p);
      for (int j = 1 + tid; j <= p; j += nt) c.hnorms[j] = 0.0;
      __syncthreads();
      for (int j = 1; j <= p; j++) {
        const double* src = c.Hp(j);
        double m = 0.0;
        for (int e = tid; e < (int)nn; e += nt) m = fmax(m, fabs(src[(e % n) + (size_t)(e / n) * c.ldh]));
        m = warp_max(m);
        if ((tid & 31) == 0)
          atomicMax((unsigned long long*)&c.hnorms[j], (unsigned long long)__double_as_longlong(m));
      }
      __syncthreads();
      int escale = 0;
      for (int j = 1; j <= p; j++) {
        const int kl = (j == 1) ? 3 : 1;
        double* dst = dstb + ((j == 1) ? 0 : pk_size(3, n) + (j - 2) * pk_size(1, n));
        const double* src = c.Hp(j);
        const double m = c.hnorms[j];
        double sc = 1.0;
        if (m > 0.0 && m < 1.7e308) {
          int e;
          (void)frexp(m, &e);
          sc = scalbn(1.0, -e);
          escale += e;
        }
        for (int e = tid; e < (int)nn; e += nt) {
          int r = e % n, cc = e / n;
          if (r <= cc + kl) dst[pk_off(kl, cc) + r] = src[r + (size_t)cc * c.ldh] * sc;
        }
      }
      if (tid == 0) dstb[pk_problem_size(n, p)] = (double)escale;
    } else if (P.use_smem) {
      if (P.wantT || P.reduce_only) {
        for (int j = 1; j <= p; j++) {
          double* dst = Ab + (size_t)((left ? (p + 1 - j) : j) - 1) * nn;
          const double* src = c.Hp(j);
          for (int e = tid; e < (int)nn; e += nt) {
            int r = e % n, cc = e / n;
            dst[e] = src[r + (size_t)cc * c.ldh];
          }
        }
      }
      if (wantZ) {
        for (int j = 1; j <= p; j++) {
          int s = (left && j > 1) ? (p + 2 - j) : j;
          double* dst = Zb + (size_t)(s - 1) * nn;
          const double* src = c.Zp(j);
          for (int e = tid; e < (int)nn; e += nt) {
            int r = e % n, cc = e / n;
            dst[e] = src[r + (size_t)cc * c.ldz];
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace psd
