// Small-bulge multishift periodic QR sweep for the large-N real standard path: the in-window
// bulge chase (host + device code, shared by the CUDA kernel in psd_ms_kernels.cuh and by the CPU
// emulation harness under tests/ms_emul/, which runs exactly these functions).
//
// What it replaces: the reference's single double-shift bulge sweep
// (PeriodicSchurDecompositions.jl:806-886) applied ~N^2 times one after the other.  Here a sweep
// carries many shift pairs at once: tightly packed chains ("packets") of NB 3x3 bulges, one packet
// per diagonal window of order W held in shared memory together with the window's accumulated
// orthogonal factors U_j (one per factor of the product).  A round moves every packet in flight
// down by D = W/2 rows inside its window; afterwards the off-window parts of H_j and the Schur
// vectors Z_j are updated with U_j by FP64 tensor-core GEMMs (psd_ms_kernels.cuh):
//     H_j[win, right of win]   <- U_j' * H_j[win, right of win]
//     H_j[above win, win]      <- H_j[above win, win] * U_{j+1}
//     Z_j[:, win]              <- Z_j[:, win] * U_j
// One bulge step is the reference's step k (:806-886): reflector from H_1[k+1:k+3, k] (or from the
// first column of the shift polynomial of the product, :768-803, when the bulge is introduced),
// applied to the rows of H_1 and the columns of H_p; then for j = p..2 a 3- and a 2-reflector that
// restore the triangular form of H_j, pushed into the columns of H_{j-1}.
//
// Parallel organisation inside a window: every bulge of the packet owns two warps, and all
// bulges advance in lockstep.  A step has three barrier-separated phases instead of the 2p + 1 a
// literal transcription needs:
//   C  chain: the 2p - 1 reflectors of the step depend only on the hanging column of H_1 and on
//      the 3 x 3 diagonal blocks of H_p .. H_2, so one warp per bulge evaluates the whole chain
//      H_1 -> H_p -> ... -> H_2 in registers and posts the reflectors in a shared-memory mailbox;
//   R  every column update of the step (H_j by the reflectors of its right neighbour, U_j);
//   L  every row update of the step, and the exact structural entries (beta, 0) of the columns
//      the reflectors were generated from.
// Within R different bulges touch disjoint columns, within L disjoint rows, and a row update
// commutes with the column updates of the other bulges, so there are no races and the result does
// not depend on the thread schedule.  (The one place where the order matters - the hanging column
// of H_1, whose entries below the subdiagonal are annihilated, not rotated - is written in C.)
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PSD_HD __host__ __device__ __forceinline__
#else
#define PSD_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define PSD_MS_WARPSYNC() __syncwarp()
#else
#define PSD_MS_WARPSYNC() ((void)0)
#endif

namespace psd {
namespace ms {

constexpr int MS_MAXP = 12;   // largest period the windowed path handles
constexpr int MS_MAXNB = 8;   // most bulges per packet (two warps each: 512 threads per window)
constexpr int MS_MAXBLK = 8;   // active diagonal blocks reported by a scan

// Window geometry chosen from the period (shared memory holds 2 p windows of order W).
struct Geom {
  int W;   // window order (multiple of 8, <= 64)
  int D;   // rows a packet advances per round (W / 2)
  int NB;  // bulges per packet
  int LD;  // leading dimension of the staged windows (odd: W + 1)
};

PSD_HD Geom geom_for(int p) {
  Geom g;
  int W = 64;
  while (W > 24 && (long long)p * W * (W + 1) * 16 > 215000) W -= 8;
  g.W = W;
  g.D = W / 2;
  g.NB = (W - 5 - g.D) / 3 + 1;
  if (g.NB > MS_MAXNB) g.NB = MS_MAXNB;
  g.LD = W + 1;
  return g;
}

// One window of one round.  All indices are global 0-based row/column numbers of the N x N
// factors; [ilo, ihi] is the active (unreduced) diagonal block the sweep works on.
struct WinDesc {
  int s;       // first row/column of the window
  int wl;      // window order (<= W)
  int kbase;   // position of the leading bulge at t = 0; bulge i sits at kbase - 3 i
  int nbul;    // bulges in this packet
  int T;       // lockstep steps of this round
  int ilo, ihi;
  int pair0;   // index of the shift pair of bulge 0 (bulge i uses (pair0 + i) mod npairs)
  int npairs;  // distinct shift pairs of the set
  int pair_off;  // first pair of the set in the pair buffer
  int intro;   // 1: bulges are introduced in this window (positions start at ilo - 1)
  int idle;    // set by the chase kernel in the device copy: every U_j of the window is the identity
               // (all bulges of the packet had been chased off), its updates are skipped
};

struct Ctx {
  int p, W, LD;
  double* Hw;  // [p][W * LD]  staged windows, (r, c) at r + c * LD
  double* Uw;  // [p][W * LD]  accumulated U_j
  const double* shifts;  // [npairs][4] = (re1, im1, re2, im2)
  int* bihi;   // [nbul] end of the block each bulge really works in (clamp_block_end)
  double* mbox;  // [MS_MAXNB][MS_MAXP][8] reflectors of the current step (see mb())
  WinDesc d;
  PSD_HD double* H(int j) const { return Hw + (size_t)(j - 1) * W * LD; }
  PSD_HD double* U(int j) const { return Uw + (size_t)(j - 1) * W * LD; }
};

// Per (bulge, role) state of the current step.
struct BState {
  int b;       // bulge index inside the packet
  int active;  // bulge takes a step at this time
  int intro;   // ... and it is the introduction step (reflector from the shift polynomial)
  int k;       // window-relative position (column the bulge hangs from; -1 at introduction)
  int r;       // first row/column the reflectors act on (k + 1)
  int nr;      // 3, or 2 at the bottom of the active block
  int ihi;     // end of the block this bulge works in
};

// Mailbox entry of (bulge b, factor j): the reflectors generated on H_j in this step,
//   [0] v1 [1] v2 [2] tau1   first reflector (order nr) on rows/columns r ..
//   [3] u1 [4] tau2          second reflector (order 2) on r+1, r+2 (tau2 = 0: none)
//   [5] beta1 [6] a0n [7] beta2   structural entries of the generating columns
constexpr int MB_STRIDE = 8;
PSD_HD double* mb(const Ctx& c, int b, int j) { return c.mbox + ((size_t)b * MS_MAXP + (j - 1)) * MB_STRIDE; }

// dlarfg for 2 or 3 entries held in registers (householder.jl:66-108); exact power-of-two
// prescale instead of the reference's sfmin loop.  x0 <- beta, (v1, v2) <- essential part.
// Division-free formulation (the reflector generation of all bulges runs redundantly in every
// warp, so it sits on the FP64 pipe): with r = 1/sqrt(a^2 + |y|^2), norm = 1/r,
//   beta = -sign(a) norm,  tau = (beta - a)/beta = 1 + |a| r,  1/(a - beta) = sign(a)/(|a| + norm).
// On the device: MUFU seeds (about 20 bits) and two Newton steps, no special-case code (the
// arguments are sums of squares / sums of magnitudes prescaled into the safe range).
PSD_HD double ms_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}
PSD_HD double ms_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / x;
#endif
}

PSD_HD double refl3(int nr, double& x0, double& v1, double& v2) {
  if (nr < 3) v2 = 0.0;
  const double amax = fmax(fabs(v1), fabs(v2));
  if (amax == 0.0) return 0.0;
  const double m = fmax(amax, fabs(x0));
  double s = 1.0, is = 1.0;
  if (m < 1e-140 || m > 1e140) {
    int e;
    (void)frexp(m, &e);
    const int ee = (-e > 1000) ? 1000 : -e;  // 2^-e overflows for subnormal m
    s = ldexp(1.0, ee);
    is = ldexp(1.0, -ee);
  }
  const double al = x0 * s, y1 = v1 * s, y2 = v2 * s;
  const double ss = fma(al, al, fma(y1, y1, y2 * y2));
  const double r = ms_rsqrt(ss);
  const double nrm = ss * r;
  const double aa = fabs(al);
  const double tau = fma(aa, r, 1.0);
  const double t = copysign(ms_rcp(aa + nrm), al);
  v1 = y1 * t;
  v2 = y2 * t;
  x0 = -copysign(nrm, al) * is;
  return tau;
}

// First column of (P - s1 I)(P - s2 I), P = H_1 H_2 ... H_p, from the leading 3 x 3 blocks at
// window-relative index o (PeriodicSchurDecompositions.jl:768-803 for a general shift pair).
PSD_HD void start_vector(const Ctx& c, int o, int pair, double& x0, double& w1, double& w2) {
  const int LD = c.LD;
  // T = H_2 ... H_p restricted to the leading 3 x 3 (upper triangular); only T[0:2, 0:2] is needed
  double t00 = 1.0, t01 = 0.0, t11 = 1.0;
  for (int j = 2; j <= c.p; j++) {
    const double* Hj = c.H(j) + o + (size_t)o * LD;
    const double a00 = Hj[0], a01 = Hj[LD], a11 = Hj[1 + LD];
    t01 = t00 * a01 + t01 * a11;
    t00 *= a00;
    t11 *= a11;
  }
  const double* H1 = c.H(1) + o + (size_t)o * LD;
  const double h00 = H1[0], h10 = H1[1];
  const double h01 = H1[LD], h11 = H1[1 + LD], h21 = H1[2 + LD];
  // P[:, 0] = H1[:, 0] t00 ; P[:, 1] = H1[:, 0] t01 + H1[:, 1] t11
  const double p00 = h00 * t00, p10 = h10 * t00;
  const double p01 = h00 * t01 + h01 * t11, p11 = h10 * t01 + h11 * t11, p21 = h21 * t11;
  double tr = 0.0, det = 0.0;  // no shift set yet: a plain (unshifted) double step
  if (pair >= 0) {
    const double* sh = c.shifts + 4 * ((size_t)c.d.pair_off + pair);
    tr = sh[0] + sh[2];                    // s1 + s2 (real for a conjugate or real pair)
    det = sh[0] * sh[2] - sh[1] * sh[3];  // s1 s2
  }
  double s = fabs(p00) + fabs(p10) + fabs(p01) + fabs(p11) + fabs(p21) + fabs(tr);
  if (s == 0.0) s = 1.0;
  const double is = 1.0 / s;
  x0 = (p00 * is) * p00 + (p01 * is) * p10 - (tr * is) * p00 + (det * is);
  w1 = p10 * ((p00 + p11 - tr) * is);
  w2 = p10 * (p21 * is);
  const double nv = fabs(x0) + fabs(w1) + fabs(w2);
  if (nv > 0.0) {
    x0 /= nv; w1 /= nv; w2 /= nv;
  }
}

// A packet is planned from block bounds that are a few rounds old.  Deflations made since then
// show up as exact zeros on the subdiagonal of H_1 (only the scan writes them, and it leaves the
// rows of every chain in flight alone): bulge b stops at the first such zero below its own reach
// (rows up to k_b + 3), wherever the bulges ahead of it are - the ones that have already left at
// that zero are carried along as no-ops (their generating vectors are exactly (x, 0, 0)).
// Returns the end of the block bulge b really works in.
PSD_HD int clamp_block_end(const Ctx& c, int b) {
  const WinDesc& d = c.d;
  int z = d.intro ? d.ilo : d.kbase - 3 * b + 3;
  if (z < d.s) z = d.s;
  const double* H1 = c.H(1);
  for (; z + 1 < d.s + d.wl && z < d.ihi; z++)
    if (H1[(z + 1 - d.s) + (size_t)(z - d.s) * c.LD] == 0.0) return z;
  return d.ihi;
}

// Rows whose subdiagonal entries the scan after a round must not touch: the chain of the packet
// as it sits after the round (first, last), or false when the packet has left (last window).
PSD_HD bool chain_after_round(const WinDesc& d, int W, int D, int& first, int& last) {
  if (d.s + W >= d.ihi + 1) return false;
  const int top = d.intro ? d.ilo + D : d.s + D;
  first = top - 1;
  last = top + 3 * d.nbul + 1;
  return true;
}

// Time-dependent part of the state of bulge b at lockstep time t.
PSD_HD void bulge_setup(const Ctx& c, BState& st, int b, int t) {
  const WinDesc& d = c.d;
  const int posg = d.kbase - 3 * b + t;
  st.b = b;
  const int ihi = (b < d.nbul) ? c.bihi[b] : d.ihi;
  st.ihi = ihi;
  st.active = (b < d.nbul) && (posg >= d.ilo - 1) && (posg <= ihi - 2) && (d.intro || posg >= d.s);
  st.intro = st.active && (posg == d.ilo - 1);
  st.k = posg - d.s;
  st.r = st.k + 1;
  const int rem = ihi - posg;  // rows below the hanging column inside the active block
  st.nr = rem >= 3 ? 3 : rem;
}

// One row (a0, a1, a2) times the reflector pair from the right / one column from the left (the
// arithmetic is the same): first reflector (v1, v2, t1) on all three, second (u1, t2) on the last two.
PSD_HD void refl_pair_apply(double& a0, double& a1, double& a2, double v1, double v2, double t1, double u1,
                            double t2) {
  const double s1 = t1 * (a0 + v1 * a1 + v2 * a2);
  a0 -= s1; a1 -= s1 * v1; a2 -= s1 * v2;
  const double s2 = t2 * (a1 + u1 * a2);
  a1 -= s2; a2 -= s2 * u1;
}

// Phase C: the reflector chain of the step, in registers (role 0; every lane computes the same
// values, lane 0 posts them).  Writes the structural entries of the hanging column of H_1.
template <int LANES>
PSD_HD void phase_chain(const Ctx& c, BState& st, int role, int lane) {
  if (!st.active || role != 0) return;
  const int LD = c.LD, r = st.r, nr = st.nr, p = c.p;
  double x0, w1, w2;
  double* H1 = c.H(1);
  if (st.intro) {
    start_vector(c, r, c.d.npairs > 0 ? (c.d.pair0 + st.b) % c.d.npairs : -1, x0, w1, w2);
  } else {
    const double* x = H1 + r + (size_t)st.k * LD;
    x0 = x[0]; w1 = x[1]; w2 = (nr == 3) ? x[2] : 0.0;
  }
  double pt1 = refl3(nr, x0, w1, w2);
  double pv1 = w1, pv2 = w2, pu1 = 0.0, pt2 = 0.0;
  PSD_MS_WARPSYNC();  // every lane has read the hanging column before lane 0 overwrites it
  if (lane == 0) {
    double* m = mb(c, st.b, 1);
    m[0] = pv1; m[1] = pv2; m[2] = pt1; m[3] = 0.0; m[4] = 0.0; m[5] = x0;
    if (!st.intro) {
      double* x = H1 + r + (size_t)st.k * LD;
      x[0] = x0; x[1] = 0.0;
      if (nr == 3) x[2] = 0.0;
    }
  }
  for (int j = p; j >= 2; j--) {
    // C = B Q_prev for the upper triangular diagonal block B of H_j (columns 0, 1 are enough)
    const double* B = c.H(j) + r + (size_t)r * LD;
    double c00 = B[0], c01 = B[LD], c02 = 0.0;
    double c10 = 0.0, c11 = B[1 + LD], c12 = 0.0;
    double c20 = 0.0, c21 = 0.0, c22 = 0.0;
    if (nr == 3) {
      c02 = B[2 * LD]; c12 = B[1 + 2 * LD]; c22 = B[2 + 2 * LD];
    }
    refl_pair_apply(c00, c01, c02, pv1, pv2, pt1, pu1, pt2);
    refl_pair_apply(c10, c11, c12, pv1, pv2, pt1, pu1, pt2);
    refl_pair_apply(c20, c21, c22, pv1, pv2, pt1, pu1, pt2);
    // QR of C: 3-reflector from column 0, then the 2-reflector from rows 1, 2 of column 1
    double y0 = c00, y1 = c10, y2 = c20;
    const double t1 = refl3(nr, y0, y1, y2);
    double a0 = c01, a1 = c11, a2 = c21;
    const double s1 = t1 * (a0 + y1 * a1 + y2 * a2);
    a0 -= s1; a1 -= s1 * y1; a2 -= s1 * y2;
    double t2 = 0.0, u1 = 0.0;
    if (nr == 3) {
      double dum = 0.0;
      t2 = refl3(2, a1, a2, dum);
      u1 = a2;
    }
    if (lane == 0) {
      double* m = mb(c, st.b, j);
      m[0] = y1; m[1] = y2; m[2] = t1; m[3] = u1; m[4] = t2; m[5] = y0; m[6] = a0; m[7] = a1;
    }
    pv1 = y1; pv2 = y2; pt1 = t1; pu1 = u1; pt2 = t2;
  }
}

// Phase R: every column update of the step.  H_j gets the reflectors generated on its right
// neighbour jn (role 0, rows 0 .. last row of the block / of the new bulge), U_jn the same
// reflectors (role 1, all rows).  The phase is latency bound (a few hundred elements per warp), so
// the strips of two factors and both 32-row passes of a strip are loaded before anything is
// computed or stored: 12 independent loads in flight per lane.
template <int LANES>
PSD_HD void phase_rights(const Ctx& c, BState& st, int role, int lane) {
  if (!st.active) return;
  const int LD = c.LD, r = st.r, nr = st.nr, p = c.p;
  const int ihl = st.ihi - c.d.s;
  constexpr int NP = (LANES == 32) ? 2 : 1;  // passes handled per trip (W <= 64 = 2 x 32 lanes)
  for (int j0 = 1; j0 <= p; j0 += 2) {
    double* col[2];
    int nrow[2];
    double v1[2], v2[2], t1[2], u1[2], t2[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int j = j0 + q;
      if (j > p) { nrow[q] = 0; col[q] = c.Hw; v1[q] = v2[q] = t1[q] = u1[q] = t2[q] = 0.0; continue; }
      const int jn = (j == p) ? 1 : j + 1;
      const double* m = mb(c, st.b, jn);
      v1[q] = m[0]; v2[q] = m[1]; t1[q] = m[2]; u1[q] = m[3]; t2[q] = m[4];
      if (role == 0) {
        col[q] = c.H(j) + (size_t)r * LD;
        nrow[q] = (j == 1) ? ((r + nr < ihl ? r + nr : ihl) + 1) : (r + nr);
      } else {
        col[q] = c.U(jn) + (size_t)r * LD;
        nrow[q] = c.d.wl;
      }
    }
    const int nmax = nrow[0] > nrow[1] ? nrow[0] : nrow[1];
    for (int rb = 0; rb < nmax; rb += NP * LANES) {
      double a[2][NP][3];
      bool on[2][NP];
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int h = 0; h < NP; h++) {
          const int rr = rb + h * LANES + lane;
          on[q][h] = rr < nrow[q];
          const int ri = on[q][h] ? rr : 0;
          a[q][h][0] = col[q][ri];
          a[q][h][1] = col[q][ri + LD];
          a[q][h][2] = (nr == 3) ? col[q][ri + 2 * LD] : 0.0;
        }
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int h = 0; h < NP; h++) refl_pair_apply(a[q][h][0], a[q][h][1], a[q][h][2], v1[q], v2[q], t1[q], u1[q], t2[q]);
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int h = 0; h < NP; h++) {
          if (!on[q][h]) continue;
          const int rr = rb + h * LANES + lane;
          col[q][rr] = a[q][h][0];
          col[q][rr + LD] = a[q][h][1];
          if (nr == 3) col[q][rr + 2 * LD] = a[q][h][2];
        }
    }
  }
}

// Phase L: every row update of the step (the two roles split the columns), and the exact
// structural entries of the generating columns of H_2 .. H_p.  Two factors per trip (independent
// loads first), as in phase R.
template <int LANES>
PSD_HD void phase_lefts(const Ctx& c, BState& st, int role, int lane) {
  if (!st.active) return;
  const int LD = c.LD, wl = c.d.wl, r = st.r, nr = st.nr, p = c.p;
  for (int j0 = 1; j0 <= p; j0 += 2) {
    double* row[2];
    int c0[2], c1[2];
    double v1[2], v2[2], t1[2], u1[2], t2[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int j = j0 + q;
      if (j > p) { c0[q] = c1[q] = 0; row[q] = c.Hw; v1[q] = v2[q] = t1[q] = u1[q] = t2[q] = 0.0; continue; }
      const double* m = mb(c, st.b, j);
      v1[q] = m[0]; v2[q] = m[1]; t1[q] = m[2]; u1[q] = m[3]; t2[q] = m[4];
      double* Hj = c.H(j);
      int cfirst = r;
      if (j > 1) {
        cfirst = (nr == 3) ? r + 2 : r + 1;
        if (role == 0 && lane == 0) {
          double* x = Hj + r + (size_t)r * LD;
          x[0] = m[5]; x[1] = 0.0;
          if (nr == 3) {
            x[2] = 0.0;
            double* y = x + LD;
            y[0] = m[6]; y[1] = m[7]; y[2] = 0.0;
          }
        }
      }
      const int mid = cfirst + (wl - cfirst + 1) / 2;
      c0[q] = (role == 0) ? cfirst : mid;
      c1[q] = (role == 0) ? mid : wl;
      row[q] = Hj + r;
    }
    const int w0 = c1[0] - c0[0], w1 = c1[1] - c0[1];
    const int wmax = w0 > w1 ? w0 : w1;
    for (int cb = 0; cb < wmax; cb += LANES) {
      double a[2][3];
      bool on[2];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int cc = c0[q] + cb + lane;
        on[q] = cc < c1[q];
        const double* e = row[q] + (size_t)(on[q] ? cc : c0[q]) * LD;
        a[q][0] = e[0];
        a[q][1] = e[1];
        a[q][2] = (nr == 3) ? e[2] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 2; q++) refl_pair_apply(a[q][0], a[q][1], a[q][2], v1[q], v2[q], t1[q], u1[q], t2[q]);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (!on[q]) continue;
        double* e = row[q] + (size_t)(c0[q] + cb + lane) * LD;
        e[0] = a[q][0];
        e[1] = a[q][1];
        if (nr == 3) e[2] = a[q][2];
      }
    }
  }
}

// The whole in-window chase.  `ex.each(f)` runs f(b, role, lane, st) for the caller's own
// (bulge, role, lane) on the device and for every (bulge, role) in turn in the host emulation;
// `ex.barrier()` separates the phases.
template <class Exec>
PSD_HD void chase_window(const Ctx& c, Exec& ex) {
  const WinDesc& d = c.d;
  constexpr int L = Exec::LANES;
  for (int t = 0; t < d.T; t++) {
    ex.each([&](int b, int role, int lane, BState& st) {
      bulge_setup(c, st, b, t);
      phase_chain<L>(c, st, role, lane);
    });
    ex.barrier();
    ex.tick(1);
    ex.each([&](int b, int role, int lane, BState& st) { phase_rights<L>(c, st, role, lane); });
    ex.barrier();
    ex.tick(2);
    ex.each([&](int b, int role, int lane, BState& st) { phase_lefts<L>(c, st, role, lane); });
    ex.barrier();
    ex.tick(3);
  }
}

// Deflation criterion of the scan (the "Test 1" of the periodic QZ drivers,
// rgeneralized.jl:1086-1112): |h(k,k-1)| <= max(smlnum, ulp (|h(k-1,k-1)| + |h(k,k)|)) on H_1.
PSD_HD bool ms_negligible(double sub, double d0, double d1, double smlnum) {
  const double tst = fabs(d0) + fabs(d1);
  return fabs(sub) <= fmax(smlnum, DBL_EPSILON * tst);
}

// Eigenvalues lre/lim[info .. m-1] (0-based; the first `info` did not converge) -> shift pairs
// (re1, im1, re2, im2): conjugate pairs as they come (positive imaginary part first,
// rschur2x2.jl:89-91), real eigenvalues two by two, a leftover real one doubled.
// perturb != 0 spreads the values deterministically (exceptional shifts).  Returns the count.
PSD_HD int pair_shifts(const double* lre, const double* lim, int info, int m, double perturb, double* pairs) {
  int np = 0;
  double pend = 0.0;
  bool have = false;
  unsigned h = 12345u;
  for (int k = info; k < m; k++) {
    double re = lre[k], im = lim[k];
    if (perturb != 0.0) {
      h = h * 1664525u + 1013904223u;
      const double f = 1.0 + perturb * ((double)(h >> 8) / 16777216.0 - 0.5);
      re *= f; im *= f;
    }
    if (im > 0.0 && k + 1 < m) {
      double* q = pairs + 4 * (size_t)np++;
      q[0] = re; q[1] = im; q[2] = re; q[3] = -im;
      k++;
    } else if (im == 0.0) {
      if (have) {
        double* q = pairs + 4 * (size_t)np++;
        q[0] = pend; q[1] = 0.0; q[2] = re; q[3] = 0.0;
        have = false;
      } else {
        pend = re;
        have = true;
      }
    }
  }
  if (have) {
    double* q = pairs + 4 * (size_t)np++;
    q[0] = pend; q[1] = 0.0; q[2] = pend; q[3] = 0.0;
  }
  return np;
}

}  // namespace ms
}  // namespace psd
