// Eigenvalue-only real periodic QR for small problems (n <= 32): ONE WARP PER PROBLEM.
//
// This is the B200 fast path of BASELINE config 2 (p=8, N=32, wantT=wantZ=false).  It
// implements the same iteration as periodic_qr_cta (PeriodicSchurDecompositions.jl:322-1096
// with wantT=false: updates restricted to the active window l..i, :675-678) but is organised
// around the hardware instead of around the reference's loop nest:
//
//  * lane L owns row L and column L of every factor; a 2- or 3-element reflector is applied
//    from the left by "column lanes" and from the right by "row lanes", so every update is
//    lane-local and no reduction is ever needed inside a sweep;
//  * the entries a reflector is generated from are forwarded with warp shuffles from the
//    lanes that have just produced them, so the serial chain reflector -> update -> next
//    reflector never waits on a shared-memory round trip;
//  * reflectors are kept un-normalised, H = I + g u u^T with g = -2/(u^T u): one rsqrt and
//    one reciprocal (MUFU seed + Newton steps) instead of dlarfg's sqrt + three divisions
//    (householder.jl:66-108); orthogonality of H depends only on g, not on the accuracy of
//    the norm;
//  * the factors live in shared memory in packed form (upper triangle + kl subdiagonals,
//    column c at offset c(c+1)/2 + kl*c): 36.4 KB per p=8,N=32 problem instead of 64 KB,
//    i.e. 6 resident problems per SM instead of 3, and both the column-lane and the row-lane
//    access patterns are bank-conflict free (triangular numbers mod 16 are a permutation);
//  * band quantities of the product (hdiag/hsub/hsup, :474-529) live in registers, one row
//    per lane, and the deflation scan (:531-585) is a single ballot.
//
// Indices in this file are 0-based.
#pragma once
#include "psd_device.cuh"

namespace psd {

__host__ __device__ inline int pk_off(int kl, int c) { return c * (c + 1) / 2 + kl * c; }
__host__ __device__ inline int pk_size(int kl, int n) { return n * (n + 1) / 2 + kl * n; }
// doubles per packed problem: H1 with 3 subdiagonals (Hessenberg + bulge), others with 1
__host__ __device__ inline int pk_problem_size(int n, int p) {
  return pk_size(3, n) + (p - 1) * pk_size(1, n);
}

struct EigParams {
  int n, p;
  long long batch;
  int maxitfac;
  const double* packed;  // [batch][pk_problem_size]
  double* eig;           // [batch][n][2]
  int* info;
  int* iters;
  unsigned long long* counter;
};

PSD_DEV double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// MUFU-seeded reciprocal square root / reciprocal with two Newton steps (inputs are kept in
// a safe range by the caller's power-of-two prescale).
PSD_DEV double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
PSD_DEV double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// Un-normalised reflector from (x0, x1[, x2]):  H = I + g u u^T, u = (x0 - beta, x1, x2),
// H x = beta e1.  Returns beta; g = 0 means H = I (householder.jl:76-78).
template <int M>
PSD_DEV double refl_u(double x0, double x1, double x2, double& u0, double& g) {
  const double amax = fmax(fabs(x1), (M == 3) ? fabs(x2) : 0.0);
  if (amax == 0.0) {
    u0 = 0.0;
    g = 0.0;
    return x0;
  }
  const double m = fmax(amax, fabs(x0));
  if (m < 1e-140 || m > 1e140) {
    // rare: fall back to the exactly-scaled dlarfg-style computation
    const double s = pow2_rescale(m);
    const double al = x0 * s, y1 = x1 * s, y2 = (M == 3) ? x2 * s : 0.0;
    const double nrm = sqrt(fma(al, al, fma(y1, y1, y2 * y2)));
    const double beta = -copysign(nrm, al);
    const double w0 = al - beta;
    // u = (w0, y1, y2)/s  ->  g' = -2/(u^T u) = -2 s^2 / (w0^2 + y1^2 + y2^2)
    u0 = w0 / s;
    g = (-2.0 * s) * (s / fma(w0, w0, fma(y1, y1, y2 * y2)));
    return beta / s;
  }
  const double ssq = (M == 3) ? fma(x1, x1, x2 * x2) : x1 * x1;
  const double nn = fma(x0, x0, ssq);
  const double nrm = nn * fast_rsqrt(nn);
  const double beta = -copysign(nrm, x0);
  u0 = x0 - beta;
  g = -2.0 * fast_rcp(fma(u0, u0, ssq));
  return beta;
}

extern __shared__ __align__(16) double psd_smem_eig[];

__global__ void __launch_bounds__(256) rpqr_eig32_kernel(EigParams P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.n, p = P.p;
  const int s1 = pk_size(3, n), sj = pk_size(1, n);
  const int psize = s1 + (p - 1) * sj;
  double* H1 = psd_smem_eig + (size_t)warp * psize;
  double* HT = H1 + s1;  // factor j (2..p) at HT + (j-2)*sj
  const int r = lane;
  // this lane's column offsets in the two packed layouts, and those of columns r+1, r+2
  const int oc1 = pk_off(3, r), ocj = pk_off(1, r);
  const int ocj1 = pk_off(1, r + 1), ocj2 = pk_off(1, r + 2);
  const int oc1p = pk_off(3, r + 1), oc1m = (r > 0) ? pk_off(3, r - 1) : 0;

  const double dat1 = 0.75, dat2 = -0.4375;
  const double ulp = DBL_EPSILON;
  const double ulpx = ulp * sqrt(sqrt(ulp));
  const double smlnum = DBL_MIN * ((double)n / ulp);

  for (;;) {
    long long b = 0;
    if (lane == 0) b = (long long)atomicAdd(P.counter, 1ULL);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= P.batch) break;
    {
      const double* src = P.packed + (size_t)b * psize;
      for (int e = lane; e < psize; e += 32) H1[e] = src[e];
    }
    __syncwarp();

    double lre = 0.0, lim = 0.0;  // eigenvalue r lives in lane r
    int info = 0, niter = 0;

    if (n == 1) {
      double q = H1[0];
      for (int j = 2; j <= p; j++) q *= HT[(j - 2) * sj];
      lre = q;
    } else {
      int i = n - 1;
      int maxitleft = P.maxitfac * n;
      while (i >= 0) {
        int l = 0;
        int its = 1;
        bool splitting = false;
        double hdiag = 0.0, hsub = 0.0, hsup = 0.0;
        while (its < maxitleft) {
          // ---- band of the product, one row per lane (:474-529) ----
          const bool act = (r >= l && r <= i);
          const bool h1b = act && (r + 1 <= i), h2b = act && (r + 2 <= i);
          double q0 = 1.0, q1 = 0.0, q2 = 0.0;
          if (act) {
            const double* Hj = HT;
            for (int j = 2; j <= p; j++, Hj += sj) {
              if (h2b) q2 = q0 * Hj[ocj2 + r] + q1 * Hj[ocj2 + r + 1] + q2 * Hj[ocj2 + r + 2];
              if (h1b) q1 = q0 * Hj[ocj1 + r] + q1 * Hj[ocj1 + r + 1];
              q0 *= Hj[ocj + r];
            }
          }
          const double t0m = __shfl_up_sync(0xffffffffu, q0, 1);
          const double t1m = __shfl_up_sync(0xffffffffu, q1, 1);
          const double t2m = __shfl_up_sync(0xffffffffu, q2, 1);
          const double t0p = __shfl_down_sync(0xffffffffu, q0, 1);
          if (act) {
            const double hd = H1[oc1 + r];
            const double hu = (r < i) ? H1[oc1p + r] : 0.0;
            if (r > l) {
              const double hs = H1[oc1m + r];
              hsub = hs * t0m;
              hdiag = hs * t1m + hd * q0;
              if (r < i) hsup = hs * t2m + hd * q1 + hu * t0p;
            } else {
              hsub = 0.0;
              hdiag = hd * q0;
              if (r < i) hsup = hd * q1 + hu * t0p;
            }
          }
          // ---- negligible-subdiagonal search (:497-585): lane k tests H[k,k-1] ----
          {
            const double hh11 = __shfl_up_sync(0xffffffffu, hdiag, 1);
            const double hh12 = __shfl_up_sync(0xffffffffu, hsup, 1);
            const bool tst = (r >= l + 1 && r <= i);
            bool found = false;
            double tst1 = fabs(hh11) + fabs(hdiag);
            const unsigned zmask = __ballot_sync(0xffffffffu, tst && tst1 == 0.0 && fabs(hsub) > smlnum);
            if (zmask) {
              // opnorm(H1[l:i,l:i], 1) fallback (:536-538): column sums, then warp max
              double cs = 0.0;
              if (act)
                for (int rr = l; rr <= min(r + 1, i); rr++) cs += fabs(H1[oc1 + rr]);
              cs = warp_max(cs);
              if (tst1 == 0.0) tst1 = cs;
            }
            if (tst) {
              const double a21 = fabs(hsub);
              if (a21 <= smlnum) {
                found = true;
              } else if (a21 <= ulp * tst1) {
                const double a12 = fabs(hh12);
                const double ab = fmax(a21, a12), ba = fmin(a21, a12);
                const double d12 = fabs(hh11 - hdiag);
                const double aa = fmax(fabs(hdiag), d12), bb = fmin(fabs(hdiag), d12);
                const double st = aa + ab;
                found = ba * (ab / st) <= fmax(smlnum, ulpx * (bb * (aa / st)));
              }
            }
            const unsigned fm = __ballot_sync(0xffffffffu, found);
            if (i > l) {
              if (fm) l = 31 - __clz(fm);
            } else {
              l = i;
            }
          }
          if (l >= i - 1) {
            splitting = true;
            break;
          }
          // ---- shifts (:679-764) and first column of the shift polynomial (:766-803) ----
          double v0, v1, v2;
          {
            const double h11 = shfl_d(hdiag, l), h12 = shfl_d(hsup, l);
            const double h21 = shfl_d(hsub, l + 1), h22 = shfl_d(hdiag, l + 1);
            const double hs3 = shfl_d(hsub, l + 2);
            const double hdi = shfl_d(hdiag, i), hdi1 = shfl_d(hdiag, i - 1);
            const double hsi = shfl_d(hsub, i), hsi1 = shfl_d(hsub, i - 1);
            const double hpi1 = shfl_d(hsup, i - 1);
            double s;
            if (its == 10 || its % 10 == 0) {
              if (its == 10)
                s = fabs(h21) + fabs(hs3);
              else
                s = fabs(hsi) + fabs(hsi1);
              const double h44 = dat1 * s + ((its == 10) ? h11 : hdi);
              const double h33 = h44;
              const double h43h34 = dat2 * s * s;
              const double h44s = h44 - h11, h33s = h33 - h11;
              v0 = (h33s * h44s - h43h34) / h21 + h12;
              v1 = h22 - h11 - h33s - h44s;
              v2 = hs3;
            } else {
              double h44 = hdi, h33 = hdi1, h43 = hsi, h34 = hpi1;
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

          // ---- double-shift sweep restricted to the window l..i (:806-886) ----
          // (f0,f1,f2): forwarded source of the next 3-reflector
          double f0 = v0, f1 = v1, f2 = v2;
          for (int k = l; k <= i - 1; k++) {
            const bool three = (k + 2 <= i);
            const int offk = pk_off(1, k), offk1 = pk_off(1, k + 1), offk2 = pk_off(1, k + 2);
            const int o3k = pk_off(3, k), o3k1 = pk_off(3, k + 1), o3k2 = pk_off(3, k + 2);
            double u0, g, beta;
            // ================= reflector on H1 (left) / H_p (right) =================
            beta = three ? refl_u<3>(f0, f1, f2, u0, g) : refl_u<2>(f0, f1, 0.0, u0, g);
            const double u1 = f1, u2 = three ? f2 : 0.0;
            __syncwarp();
            if (k > l && r == k - 1) {  // H1[k,k-1] = beta, bulge entries of column k-1 -> 0
              double* q = H1 + oc1 + k;
              q[0] = beta;
              q[1] = 0.0;
              if (three) q[2] = 0.0;
            }
            // left: H1 rows k..k+2, columns k..i (column lanes)
            if (r >= k && r <= i) {
              double* q = H1 + oc1 + k;
              const double a0 = q[0], a1 = q[1], a2 = three ? q[2] : 0.0;
              const double s = g * fma(u2, a2, fma(u1, a1, u0 * a0));
              q[0] = fma(s, u0, a0);
              q[1] = fma(s, u1, a1);
              if (three) q[2] = fma(s, u2, a2);
            }
            if (p == 1) {
              // right: H1 itself, rows l..min(k+3,i) (row lanes); forward H1[k+1..k+3, k]
              __syncwarp();
              const int rmax = min(k + 3, i);
              double a0n = 0.0;
              if (r >= l && r <= rmax) {
                double *q0p = H1 + o3k + r, *q1p = H1 + o3k1 + r, *q2p = H1 + o3k2 + r;
                const double a0 = *q0p, a1 = *q1p, a2 = three ? *q2p : 0.0;
                const double s = g * fma(u2, a2, fma(u1, a1, u0 * a0));
                a0n = fma(s, u0, a0);
                *q0p = a0n;
                *q1p = fma(s, u1, a1);
                if (three) *q2p = fma(s, u2, a2);
              }
              f0 = shfl_d(a0n, k + 1);
              f1 = shfl_d(a0n, min(k + 2, 31));
              f2 = shfl_d(a0n, min(k + 3, 31));
              continue;
            }
            // right: H_p rows l..k+2, columns k..k+2 (row lanes); forward H_p[k..k+2, k]
            {
              double* Hp_ = HT + (p - 2) * sj;
              const int rmax = three ? k + 2 : k + 1;
              double a0n = 0.0;
              if (r >= l && r <= rmax) {
                double *q0p = Hp_ + offk + r, *q1p = Hp_ + offk1 + r, *q2p = Hp_ + offk2 + r;
                const double a0 = (r <= k + 1) ? *q0p : 0.0;
                const double a1 = *q1p, a2 = three ? *q2p : 0.0;
                const double s = g * fma(u2, a2, fma(u1, a1, u0 * a0));
                a0n = fma(s, u0, a0);
                if (r < k) *q0p = a0n;
                *q1p = fma(s, u1, a1);
                if (three) *q2p = fma(s, u2, a2);
              }
              f0 = shfl_d(a0n, k);
              f1 = shfl_d(a0n, k + 1);
              f2 = shfl_d(a0n, min(k + 2, 31));
            }
            // ================= factors p..2 =================
            for (int j = p; j >= 2; j--) {
              double* Hj = HT + (j - 2) * sj;
              double* Hm = (j == 2) ? H1 : (Hj - sj);  // H_{j-1}
              beta = three ? refl_u<3>(f0, f1, f2, u0, g) : refl_u<2>(f0, f1, 0.0, u0, g);
              const double w1 = f1, w2 = three ? f2 : 0.0;
              // right: H_{j-1} columns k..k+2 (row lanes) and forward the next source
              if (j > 2) {
                const int rmax = three ? k + 2 : k + 1;
                double a0n = 0.0;
                if (r >= l && r <= rmax) {
                  double *q0p = Hm + offk + r, *q1p = Hm + offk1 + r, *q2p = Hm + offk2 + r;
                  const double a0 = (r <= k + 1) ? *q0p : 0.0;
                  const double a1 = *q1p, a2 = three ? *q2p : 0.0;
                  const double s = g * fma(w2, a2, fma(w1, a1, u0 * a0));
                  a0n = fma(s, u0, a0);
                  if (r < k) *q0p = a0n;
                  *q1p = fma(s, w1, a1);
                  if (three) *q2p = fma(s, w2, a2);
                }
                f0 = shfl_d(a0n, k);
                f1 = shfl_d(a0n, k + 1);
                f2 = shfl_d(a0n, min(k + 2, 31));
              } else {
                // H1 is Hessenberg: rows l..min(k+3,i); next source is H1[k+1..k+3, k]
                const int rmax = min(k + 3, i);
                double a0n = 0.0;
                if (r >= l && r <= rmax) {
                  double *q0p = H1 + o3k + r, *q1p = H1 + o3k1 + r, *q2p = H1 + o3k2 + r;
                  const double a0 = *q0p, a1 = *q1p, a2 = three ? *q2p : 0.0;
                  const double s = g * fma(w2, a2, fma(w1, a1, u0 * a0));
                  a0n = fma(s, u0, a0);
                  *q0p = a0n;
                  *q1p = fma(s, w1, a1);
                  if (three) *q2p = fma(s, w2, a2);
                }
                f0 = shfl_d(a0n, k + 1);
                f1 = shfl_d(a0n, min(k + 2, 31));
                f2 = shfl_d(a0n, min(k + 3, 31));
              }
              __syncwarp();  // row-lane writes to H_j (previous stage) visible to column lanes
              // left: H_j rows k..k+2, columns k+1..i (column lanes); column k gets (beta, 0)
              double y0 = 0.0, y1 = 0.0;
              if (r == k) {
                double* q = Hj + ocj + k;
                q[0] = beta;
                q[1] = 0.0;
              } else if (r > k && r <= i) {
                double* q = Hj + ocj + k;
                const double a0 = q[0], a1 = q[1], a2 = three ? q[2] : 0.0;
                const double s = g * fma(w2, a2, fma(w1, a1, u0 * a0));
                q[0] = fma(s, u0, a0);
                y0 = fma(s, w1, a1);
                y1 = three ? fma(s, w2, a2) : 0.0;
                q[1] = y0;
                if (three) q[2] = y1;
              }
              if (three) {
                // second reflector, order 2, from H_j[k+1..k+2, k+1] held by column lane k+1
                y0 = shfl_d(y0, k + 1);
                y1 = shfl_d(y1, k + 1);
                double c0, h;
                const double beta2 = refl_u<2>(y0, y1, 0.0, c0, h);
                // left: H_j rows k+1,k+2, columns k+2..i
                if (r == k + 1) {
                  double* q = Hj + ocj + k + 1;
                  q[0] = beta2;
                  q[1] = 0.0;
                } else if (r > k + 1 && r <= i) {
                  double* q = Hj + ocj + k + 1;
                  const double a0 = q[0], a1 = q[1];
                  const double s = h * fma(y1, a1, c0 * a0);
                  q[0] = fma(s, c0, a0);
                  q[1] = fma(s, y1, a1);
                }
                // right: H_{j-1} columns k+1,k+2, rows l..k+2 (k+3 for H1)
                if (j > 2) {
                  if (r >= l && r <= k + 2) {
                    double *q1p = Hm + offk1 + r, *q2p = Hm + offk2 + r;
                    const double a1 = *q1p, a2 = *q2p;
                    const double s = h * fma(y1, a2, c0 * a1);
                    *q1p = fma(s, c0, a1);
                    *q2p = fma(s, y1, a2);
                  }
                } else {
                  if (r >= l && r <= min(k + 3, i)) {
                    double *q1p = H1 + o3k1 + r, *q2p = H1 + o3k2 + r;
                    const double a1 = *q1p, a2 = *q2p;
                    const double s = h * fma(y1, a2, c0 * a1);
                    *q1p = fma(s, c0, a1);
                    *q2p = fma(s, y1, a2);
                  }
                }
              }
            }  // factors
          }    // k
          __syncwarp();
          its++;
        }  // QR iterations

        if (!splitting) {
          info = i + 1;  // "convergence failed at level i" (:891-893), 1-based level
          niter += its;
          break;
        }
        // ---- deflation (:895-934, wantT = false) ----
        if (l == i) {
          if (r == i) {
            lre = hdiag;
            lim = 0.0;
          }
        } else {
          double a = shfl_d(hdiag, i - 1), bq = shfl_d(hsup, i - 1);
          double cq = shfl_d(hsub, i), d = shfl_d(hdiag, i);
          double cs, sn, l1r, l1i, l2r, l2i;
          gs2x2(a, bq, cq, d, cs, sn, l1r, l1i, l2r, l2i);
          if (r == i - 1) {
            lre = l1r;
            lim = l1i;
          } else if (r == i) {
            lre = l2r;
            lim = l2i;
          }
        }
        maxitleft -= its;
        niter += its;
        i = l - 1;
      }
    }
    if (r < n) {
      double* eg = P.eig + ((size_t)b * n + r) * 2;
      eg[0] = lre;
      eg[1] = lim;
    }
    if (lane == 0) {
      P.info[b] = info;
      if (P.iters) P.iters[b] = niter;
    }
    __syncwarp();
  }
}

}  // namespace psd
