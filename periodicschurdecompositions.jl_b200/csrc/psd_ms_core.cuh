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
// Parallel organisation inside a window: every bulge of the packet owns two warps.  All bulges
// advance in lockstep; a step is 2p barrier-separated phases (generate + row update of H_j /
// column update of H_{j-1} and U_j), in each of which different bulges touch disjoint rows
// (row phases) or disjoint columns (column phases), so there are no races and the result does not
// depend on the thread schedule.  Reflectors are computed redundantly by every lane from
// broadcast shared-memory reads (no shuffles, no mailbox).
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
};

struct Ctx {
  int p, W, LD;
  double* Hw;  // [p][W * LD]  staged windows, (r, c) at r + c * LD
  double* Uw;  // [p][W * LD]  accumulated U_j
  const double* shifts;  // [npairs][4] = (re1, im1, re2, im2)
  int* bihi;   // [nbul] end of the block each bulge really works in (clamp_block_end)
  WinDesc d;
  PSD_HD double* H(int j) const { return Hw + (size_t)(j - 1) * W * LD; }
  PSD_HD double* U(int j) const { return Uw + (size_t)(j - 1) * W * LD; }
};

// Per (bulge, role) state carried through the phases of one step.
struct BState {
  int b;       // bulge index inside the packet
  int active;  // bulge takes a step at this time
  int intro;   // ... and it is the introduction step (reflector from the shift polynomial)
  int k;       // window-relative position (column the bulge hangs from; -1 at introduction)
  int r;       // first row/column the reflectors act on (k + 1)
  int nr;      // 3, or 2 at the bottom of the active block
  int ihi;     // end of the block this bulge works in
  // reflectors generated in the last gen phase: first (order nr) at r, second (order 2) at r + 1
  double v1, v2, tau1, u1, tau2;
  int have2;
  // structural entries of the generating columns, written at the start of the next phase
  double beta1, a0n, beta2;
  int defer_j;  // factor whose columns get them (0 = none)
};

// dlarfg for 2 or 3 entries held in registers (householder.jl:66-108); exact power-of-two
// prescale instead of the reference's sfmin loop.  x0 <- beta, (v1, v2) <- essential part.
// Division-free formulation (the reflector generation of all bulges runs redundantly in every
// warp, so it sits on the FP64 pipe): with r = 1/sqrt(a^2 + |y|^2), norm = 1/r,
//   beta = -sign(a) norm,  tau = (beta - a)/beta = 1 + |a| r,  1/(a - beta) = sign(a)/(|a| + norm).
PSD_HD double ms_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
PSD_HD double ms_rcp(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
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
  const double* sh = c.shifts + 4 * ((size_t)c.d.pair_off + pair);
  const double tr = sh[0] + sh[2];                    // s1 + s2 (real for a conjugate or real pair)
  const double det = sh[0] * sh[2] - sh[1] * sh[3];  // s1 s2
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
  st.defer_j = 0;
  st.have2 = 0;
}

// Row phase of factor j: generate the reflector(s) of every active bulge and apply them to the
// rows r .. r+nr-1 of H_j inside the window.  role 0 / 1 split the columns.  LANES = 32 on the
// device (lane = threadIdx & 31), 1 in the host emulation.
// part: 1 = generate only, 2 = row update only (after a barrier), 3 = both.
template <int LANES>
PSD_HD void phase_gen_left(const Ctx& c, BState& st, int j, int role, int lane, int part = 3) {
  if (!st.active) return;
  const int LD = c.LD, wl = c.d.wl, r = st.r, nr = st.nr;
  double* Hj = c.H(j);
  int cfirst;  // first column updated by the row operation
  if (part == 2) {
    cfirst = (j == 1) ? r : (st.have2 ? r + 2 : r + 1);
  } else
  if (j == 1) {
    double x0, w1, w2;
    if (st.intro) {
      start_vector(c, r, (c.d.pair0 + st.b) % c.d.npairs, x0, w1, w2);
    } else {
      const double* x = Hj + r + (size_t)st.k * LD;
      x0 = x[0]; w1 = x[1]; w2 = (nr == 3) ? x[2] : 0.0;
    }
    st.tau1 = refl3(nr, x0, w1, w2);
    st.v1 = w1; st.v2 = w2;
    st.beta1 = x0;
    st.have2 = 0;
    st.defer_j = st.intro ? 0 : 1;
    cfirst = r;
  } else {
    const double* x = Hj + r + (size_t)r * LD;
    double x0 = x[0], w1 = x[1], w2 = (nr == 3) ? x[2] : 0.0;
    st.tau1 = refl3(nr, x0, w1, w2);
    st.v1 = w1; st.v2 = w2;
    st.beta1 = x0;
    st.defer_j = j;
    if (nr == 3) {
      // column r+1 after the first reflector, then the 2-reflector that clears H_j[r+2, r+1]
      const double* y = Hj + r + (size_t)(r + 1) * LD;
      double a0 = y[0], a1 = y[1], a2 = y[2];
      const double s1 = st.tau1 * (a0 + st.v1 * a1 + st.v2 * a2);
      a0 -= s1; a1 -= s1 * st.v1; a2 -= s1 * st.v2;
      double dum = 0.0;
      st.tau2 = refl3(2, a1, a2, dum);
      st.u1 = a2;
      st.a0n = a0;
      st.beta2 = a1;
      st.have2 = 1;
      cfirst = r + 2;
    } else {
      st.have2 = 0;
      cfirst = r + 1;
    }
  }
  if (part == 1) return;
  PSD_MS_WARPSYNC();  // every lane has read the generating entries before any lane writes
  // the two roles split the columns
  int c0 = cfirst, c1 = wl;
  {
    const int mid = cfirst + (wl - cfirst + 1) / 2;
    if (role == 0) c1 = mid; else c0 = mid;
  }
  const double v1 = st.v1, v2 = st.v2, t1 = st.tau1, u1 = st.u1, t2 = st.tau2;
  const bool two = st.have2 != 0;
  for (int cc = c0 + lane; cc < c1; cc += LANES) {
    double* a = Hj + r + (size_t)cc * LD;
    double a0 = a[0], a1 = a[1], a2 = (nr == 3) ? a[2] : 0.0;
    const double s1 = t1 * (a0 + v1 * a1 + v2 * a2);
    a0 -= s1; a1 -= s1 * v1; a2 -= s1 * v2;
    if (two) {
      const double s2 = t2 * (a1 + u1 * a2);
      a1 -= s2; a2 -= s2 * u1;
    }
    a[0] = a0; a[1] = a1;
    if (nr == 3) a[2] = a2;
  }
}

// Structural entries of the generating columns (beta, exact zeros) of the factor the reflectors
// were generated on; stored by one lane in the phase after the generation.
PSD_HD void flush_deferred(const Ctx& c, BState& st, int role, int lane) {
  if (!st.active || role != 0 || lane != 0 || st.defer_j == 0) return;
  const int LD = c.LD, r = st.r, nr = st.nr;
  double* G = c.H(st.defer_j);
  if (st.defer_j == 1) {
    double* x = G + r + (size_t)st.k * LD;
    x[0] = st.beta1; x[1] = 0.0;
    if (nr == 3) x[2] = 0.0;
  } else {
    double* x = G + r + (size_t)r * LD;
    x[0] = st.beta1; x[1] = 0.0;
    if (nr == 3) {
      x[2] = 0.0;
      double* y = G + r + (size_t)(r + 1) * LD;
      y[0] = st.a0n; y[1] = st.beta2; y[2] = 0.0;
    }
  }
  st.defer_j = 0;
}

// Column phase: the reflectors generated on factor jn (the phase before) are applied to the
// columns r .. r+nr-1 of H_j (role 0, rows 0 .. rlast) and of U_jn (role 1, all rows); role 0
// first stores the structural entries of the generating columns of H_jn.
template <int LANES>
PSD_HD void phase_right(const Ctx& c, BState& st, int j, int jn, int role, int lane) {
  if (!st.active) return;
  const int LD = c.LD, r = st.r, nr = st.nr;
  flush_deferred(c, st, role, lane);
  double* M;
  int nrow;
  if (role == 0) {
    M = c.H(j);
    const int ihl = st.ihi - c.d.s;
    nrow = (j == 1) ? ((r + nr < ihl ? r + nr : ihl) + 1) : (r + nr);
  } else {
    M = c.U(jn);
    nrow = c.d.wl;
  }
  const double v1 = st.v1, v2 = st.v2, t1 = st.tau1, u1 = st.u1, t2 = st.tau2;
  const bool two = st.have2 != 0;
  double* col = M + (size_t)r * LD;
  for (int rr = lane; rr < nrow; rr += LANES) {
    double a0 = col[rr], a1 = col[rr + LD], a2 = (nr == 3) ? col[rr + 2 * LD] : 0.0;
    const double s1 = t1 * (a0 + v1 * a1 + v2 * a2);
    a0 -= s1; a1 -= s1 * v1; a2 -= s1 * v2;
    if (two) {
      const double s2 = t2 * (a1 + u1 * a2);
      a1 -= s2; a2 -= s2 * u1;
    }
    col[rr] = a0; col[rr + LD] = a1;
    if (nr == 3) col[rr + 2 * LD] = a2;
  }
  st.defer_j = 0;
}

// The whole in-window chase.  `ex.each(f)` runs f(b, role, lane, st) for the caller's own
// (bulge, role, lane) on the device and for every (bulge, role) in turn in the host emulation;
// `ex.barrier()` separates the phases.
template <class Exec>
PSD_HD void chase_window(const Ctx& c, Exec& ex) {
  const int p = c.p;
  const WinDesc& d = c.d;
  constexpr int L = Exec::LANES;
  for (int t = 0; t < d.T; t++) {
    if (d.intro) {
      // a bulge that is being introduced reads the leading 3 x 3 blocks, which its own row update
      // changes: generate first, update after a barrier
      ex.each([&](int b, int role, int lane, BState& st) {
        bulge_setup(c, st, b, t);
        phase_gen_left<L>(c, st, 1, role, lane, 1);
      });
      ex.barrier();
      ex.each([&](int b, int role, int lane, BState& st) { phase_gen_left<L>(c, st, 1, role, lane, 2); });
    } else {
      ex.each([&](int b, int role, int lane, BState& st) {
        bulge_setup(c, st, b, t);
        phase_gen_left<L>(c, st, 1, role, lane);
      });
    }
    ex.barrier();
    for (int j = p; j >= 2; j--) {
      const int jn = (j == p) ? 1 : j + 1;
      ex.each([&](int b, int role, int lane, BState& st) { phase_right<L>(c, st, j, jn, role, lane); });
      ex.barrier();
      ex.each([&](int b, int role, int lane, BState& st) { phase_gen_left<L>(c, st, j, role, lane); });
      ex.barrier();
    }
    if (p == 1) {
      // H_1 is both the generating and the target factor: its structural entries must be in
      // place before the column phase of the neighbouring bulges reads them
      ex.each([&](int b, int role, int lane, BState& st) { flush_deferred(c, st, role, lane); });
      ex.barrier();
    }
    {
      const int jn = (p == 1) ? 1 : 2;
      ex.each([&](int b, int role, int lane, BState& st) { phase_right<L>(c, st, 1, jn, role, lane); });
      ex.barrier();
    }
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
