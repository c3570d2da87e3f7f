// Periodic Hessenberg-triangular reduction for small problems (n <= 32), no Schur vectors:
// ONE WARP PER PROBLEM, output in the packed layout consumed by rpqr_eig32_kernel.
//
// Same mathematics as phessenberg! (PeriodicSchurDecompositions.jl:229-247): for each column
// i, factors p..2 get a QR-type reflector (rows i..n-1) that is pushed into the right
// neighbour from the right, then H_1 gets the Hessenberg reflector (rows i+1..n-1) which is
// pushed into H_p.  Lane L owns row L / column L; the only cross-lane operation per
// reflector is one warp-wide sum of squares.  Reflectors are used un-normalised,
// H = I + g u u^T with g = -2/(u^T u) (see psd_real_eig32.cuh).
#pragma once
#include "psd_real_eig32.cuh"

namespace psd {

struct Hess32Params {
  int n, p;
  long long batch;
  int left;             // :L orientation: internal factor j <- user factor p+1-j (:127-131)
  int ld;               // smem leading dimension (odd)
  const double* A;      // [batch][p][n*n]
  double* packed_out;   // [batch][pk_problem_stride(n,p)]
  unsigned long long* counter;
};

// One reflector step: generate from Aj[r0.., col]; apply from the left to Aj[r0.., col+1..],
// from the right to Am[:, r0..].  0-based.  All 32 lanes must call.
PSD_DEV void hess32_step(double* Aj, double* Am, int n, int ld, int r0, int col, int lane) {
  double* xc = Aj + col * ld;  // source column
  const double alpha = xc[r0];
  const double xr = (lane > r0 && lane < n) ? xc[lane] : 0.0;
  double ssq = warp_sum(xr * xr);
  if (ssq == 0.0) {
    // either an exactly zero tail (H = I, householder.jl:76-78) or underflow of the squares
    const double amax = warp_max(fabs(xr));
    if (amax == 0.0) return;
  }
  double nn = fma(alpha, alpha, ssq);
  double sc = 1.0;
  if (!(ssq > 1e-280 && nn < 1e280)) {
    // rare: rescale by an exact power of two so that the squares are representable
    const double m = fmax(warp_max(fabs(xr)), fabs(alpha));
    sc = pow2_rescale(m);
    const double y = xr * sc;
    ssq = warp_sum(y * y);
    nn = fma(alpha * sc, alpha * sc, ssq);
  }
  const double al = alpha * sc;
  const double nrm = nn * fast_rsqrt(nn);
  const double betas = -copysign(nrm, al);
  const double u0s = al - betas;
  // H = I + g u u^T with u = (u0s, sc*x[r0+1..]) ; fold sc into the coefficients
  const double gs = -2.0 * fast_rcp(fma(u0s, u0s, ssq));
  double g = gs, u0 = u0s, beta = betas;
  if (sc != 1.0) {  // for u = (u0, x) with u0 = u0s/sc
    g = gs * sc * sc;
    u0 = u0s / sc;
    beta = betas / sc;
  }
  if (Am != Aj) {
    // Left update of Aj (lane = column, columns col+1..n-1) and right update of Am (lane = row)
    // touch different matrices: both dot products and both updates run in the same loops, so the
    // two dependent chains overlap (instruction-level parallelism within the warp).
    const bool la = (lane > col && lane < n), rw = (lane < n);
    double* cl = Aj + (la ? lane : col + 1 < n ? col + 1 : col) * ld;  // inactive lanes read a valid column
    double* ar = Am + (rw ? lane : 0);
    double dl0 = u0 * cl[r0], dl1 = 0.0, dr0 = ar[r0 * ld] * u0, dr1 = 0.0;
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 2) {
      const bool p1 = k + 1 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0;
      const double l0 = cl[k], l1 = p1 ? cl[k + 1] : 0.0;
      const double q0 = ar[k * ld], q1 = p1 ? ar[(k + 1) * ld] : 0.0;
      dl0 = fma(x0, l0, dl0); dl1 = fma(x1, l1, dl1);
      dr0 = fma(q0, x0, dr0); dr1 = fma(q1, x1, dr1);
    }
    const double sl = g * (dl0 + dl1), sr = g * (dr0 + dr1);
    if (la) cl[r0] = fma(sl, u0, cl[r0]);
    if (rw) ar[r0 * ld] = fma(sr, u0, ar[r0 * ld]);
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 2) {
      const bool p1 = k + 1 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0;
      const double l0 = cl[k], l1 = p1 ? cl[k + 1] : 0.0;
      const double q0 = ar[k * ld], q1 = p1 ? ar[(k + 1) * ld] : 0.0;
      if (la) {
        cl[k] = fma(sl, x0, l0);
        if (p1) cl[k + 1] = fma(sl, x1, l1);
      }
      if (rw) {
        ar[k * ld] = fma(sr, x0, q0);
        if (p1) ar[(k + 1) * ld] = fma(sr, x1, q1);
      }
    }
    __syncwarp();
    if (lane == r0) xc[lane] = beta;
    else if (lane > r0 && lane < n) xc[lane] = 0.0;
    __syncwarp();
    return;
  }
  // left: columns col+1..n-1 of Aj (lane = column).  Loops are unrolled by 4 with independent
  // accumulators and predicated loads so that the shared-memory loads pipeline.
  if (lane > col && lane < n) {
    double* a = Aj + lane * ld;
    double d0 = u0 * a[r0], d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 4) {
      const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
      const double a0 = a[k], a1 = p1 ? a[k + 1] : 0.0, a2 = p2 ? a[k + 2] : 0.0, a3 = p3 ? a[k + 3] : 0.0;
      d0 = fma(x0, a0, d0); d1 = fma(x1, a1, d1); d2 = fma(x2, a2, d2); d3 = fma(x3, a3, d3);
    }
    const double s = g * ((d0 + d1) + (d2 + d3));
    a[r0] = fma(s, u0, a[r0]);
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 4) {
      const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
      const double a0 = a[k], a1 = p1 ? a[k + 1] : 0.0, a2 = p2 ? a[k + 2] : 0.0, a3 = p3 ? a[k + 3] : 0.0;
      a[k] = fma(s, x0, a0);
      if (p1) a[k + 1] = fma(s, x1, a1);
      if (p2) a[k + 2] = fma(s, x2, a2);
      if (p3) a[k + 3] = fma(s, x3, a3);
    }
  }
  if (Am == Aj) __syncwarp();
  // right: rows 0..n-1 of Am (lane = row), columns r0..n-1
  if (lane < n) {
    double* a = Am + lane;
    double d0 = a[r0 * ld] * u0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 4) {
      const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
      const double a0 = a[k * ld], a1 = p1 ? a[(k + 1) * ld] : 0.0, a2 = p2 ? a[(k + 2) * ld] : 0.0,
                   a3 = p3 ? a[(k + 3) * ld] : 0.0;
      d0 = fma(a0, x0, d0); d1 = fma(a1, x1, d1); d2 = fma(a2, x2, d2); d3 = fma(a3, x3, d3);
    }
    const double s = g * ((d0 + d1) + (d2 + d3));
    a[r0 * ld] = fma(s, u0, a[r0 * ld]);
#pragma unroll 2
    for (int k = r0 + 1; k < n; k += 4) {
      const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
      const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
      const double a0 = a[k * ld], a1 = p1 ? a[(k + 1) * ld] : 0.0, a2 = p2 ? a[(k + 2) * ld] : 0.0,
                   a3 = p3 ? a[(k + 3) * ld] : 0.0;
      a[k * ld] = fma(s, x0, a0);
      if (p1) a[(k + 1) * ld] = fma(s, x1, a1);
      if (p2) a[(k + 2) * ld] = fma(s, x2, a2);
      if (p3) a[(k + 3) * ld] = fma(s, x3, a3);
    }
  }
  __syncwarp();
  if (lane == r0) xc[lane] = beta;
  else if (lane > r0 && lane < n) xc[lane] = 0.0;
  __syncwarp();
}

extern __shared__ __align__(16) double psd_smem_hess[];

template <int NN, int PP>
__global__ void __launch_bounds__(256) rphess_warp32_kernel_t(Hess32Params P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = NN ? NN : P.n, p = PP ? PP : P.p, ld = NN ? (NN % 2 == 0 ? NN + 1 : NN) : P.ld;
  const size_t nn = (size_t)n * n;
  const int fs = ld * n;  // doubles per staged factor
  double* S = psd_smem_hess + (size_t)warp * p * fs;
  const int psize = pk_problem_size(n, p);
  for (;;) {
    long long b = 0;
    if (lane == 0) b = (long long)atomicAdd(P.counter, 1ULL);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= P.batch) break;
    const double* Ab = P.A + (size_t)b * p * nn;
    // stage the whole problem with asynchronous 8-byte copies (LDGSTS): every load of the
    // problem is in flight before the first one is waited for
    for (int j = 0; j < p; j++) {
      const double* src = Ab + (size_t)(P.left ? (p - 1 - j) : j) * nn;
      double* dst = S + j * fs;
      if (lane < n)
        for (int c = 0; c < n; c++) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + c * ld + lane);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + c * n + lane));
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    // normalise every factor by an exact power of two (see pk_problem_stride)
    int escale = 0;
    for (int j = 0; j < p; j++) {
      double* dst = S + j * fs;
      double m = 0.0;
      if (lane < n)
        for (int c = 0; c < n; c++) m = fmax(m, fabs(dst[c * ld + lane]));
      m = warp_max(m);
      if (m > 0.0 && m < 1.7e308) {
        int e;
        (void)frexp(m, &e);
        if (e != 0) {
          const double sc = scalbn(1.0, (-e > 1000) ? 1000 : -e);
          if (lane < n)
            for (int c = 0; c < n; c++) dst[c * ld + lane] *= sc;
          escale += e;
        }
      }
    }
    __syncwarp();
    for (int i = 0; i < n - 1; i++) {
      for (int j = p - 1; j >= 1; j--) hess32_step(S + j * fs, S + (j - 1) * fs, n, ld, i, i, lane);
      if (n - (i + 1) > 1) hess32_step(S, S + (p - 1) * fs, n, ld, i + 1, i, lane);
    }
    // packed output: H1 with 3 subdiagonals of storage, H2..Hp with 1 (zeros below structure)
    double* dstb = P.packed_out + (size_t)b * (psize + PK_STATE);
    if (lane == 0) dstb[psize] = (double)escale;
    for (int j = 0; j < p; j++) {
      const int kl = (j == 0) ? 3 : 1;
      const int keep = (j == 0) ? 1 : 0;
      double* dst = dstb + ((j == 0) ? 0 : pk_size(3, n) + (j - 1) * pk_size(1, n));
      const double* src = S + j * fs;
      for (int c = 0; c < n; c++) {
        const int off = pk_off(kl, c);
        if (lane <= c + kl && lane < n) dst[off + lane] = (lane <= c + keep) ? src[c * ld + lane] : 0.0;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Two warps per problem (p >= 2).  Shared memory holds only three problems per SM, so the plain
// kernel runs three warps per SM and every problem pays the full latency of its serial chain of
// p (n-1) reflector steps.  Here the two independent halves of a step run on two warps (two
// schedulers): warp 0 of the pair applies the reflector to the factor itself from the left, warp 1
// pushes it into the neighbour factor from the right; both generate the reflector redundantly
// from the same shared-memory column (no exchange), and one 64-thread named barrier per step
// orders the neighbour's update before the next reflector is generated from it.
// ---------------------------------------------------------------------------------------------
PSD_DEV void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

PSD_DEV void hess32_step_pair(double* Aj, double* Am, int n, int ld, int r0, int col, int lane, int role, int barid) {
  double* xc = Aj + col * ld;  // source column
  const double alpha = xc[r0];
  const double xr = (lane > r0 && lane < n) ? xc[lane] : 0.0;
  double ssq = warp_sum(xr * xr);
  if (ssq == 0.0) {
    const double amax = warp_max(fabs(xr));
    if (amax == 0.0) return;  // H = I for both warps (same data, same decision)
  }
  double nn = fma(alpha, alpha, ssq);
  double sc = 1.0;
  if (!(ssq > 1e-280 && nn < 1e280)) {
    const double m = fmax(warp_max(fabs(xr)), fabs(alpha));
    sc = pow2_rescale(m);
    const double y = xr * sc;
    ssq = warp_sum(y * y);
    nn = fma(alpha * sc, alpha * sc, ssq);
  }
  const double al = alpha * sc;
  const double nrm = nn * fast_rsqrt(nn);
  const double betas = -copysign(nrm, al);
  const double u0s = al - betas;
  const double gs = -2.0 * fast_rcp(fma(u0s, u0s, ssq));
  double g = gs, u0 = u0s, beta = betas;
  if (sc != 1.0) {
    g = gs * sc * sc;
    u0 = u0s / sc;
    beta = betas / sc;
  }
  if (role == 0) {
    // left: columns col+1..n-1 of Aj (lane = column)
    if (lane > col && lane < n) {
      double* a = Aj + lane * ld;
      double d0 = u0 * a[r0], d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 2
      for (int k = r0 + 1; k < n; k += 4) {
        const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
        const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
        const double a0 = a[k], a1 = p1 ? a[k + 1] : 0.0, a2 = p2 ? a[k + 2] : 0.0, a3 = p3 ? a[k + 3] : 0.0;
        d0 = fma(x0, a0, d0); d1 = fma(x1, a1, d1); d2 = fma(x2, a2, d2); d3 = fma(x3, a3, d3);
      }
      const double s = g * ((d0 + d1) + (d2 + d3));
      a[r0] = fma(s, u0, a[r0]);
#pragma unroll 2
      for (int k = r0 + 1; k < n; k += 4) {
        const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
        const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
        const double a0 = a[k], a1 = p1 ? a[k + 1] : 0.0, a2 = p2 ? a[k + 2] : 0.0, a3 = p3 ? a[k + 3] : 0.0;
        a[k] = fma(s, x0, a0);
        if (p1) a[k + 1] = fma(s, x1, a1);
        if (p2) a[k + 2] = fma(s, x2, a2);
        if (p3) a[k + 3] = fma(s, x3, a3);
      }
    }
  } else {
    // right: rows 0..n-1 of Am (lane = row), columns r0..n-1
    if (lane < n) {
      double* a = Am + lane;
      double d0 = a[r0 * ld] * u0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll 2
      for (int k = r0 + 1; k < n; k += 4) {
        const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
        const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
        const double a0 = a[k * ld], a1 = p1 ? a[(k + 1) * ld] : 0.0, a2 = p2 ? a[(k + 2) * ld] : 0.0,
                     a3 = p3 ? a[(k + 3) * ld] : 0.0;
        d0 = fma(a0, x0, d0); d1 = fma(a1, x1, d1); d2 = fma(a2, x2, d2); d3 = fma(a3, x3, d3);
      }
      const double s = g * ((d0 + d1) + (d2 + d3));
      a[r0 * ld] = fma(s, u0, a[r0 * ld]);
#pragma unroll 2
      for (int k = r0 + 1; k < n; k += 4) {
        const bool p1 = k + 1 < n, p2 = k + 2 < n, p3 = k + 3 < n;
        const double x0 = xc[k], x1 = p1 ? xc[k + 1] : 0.0, x2 = p2 ? xc[k + 2] : 0.0, x3 = p3 ? xc[k + 3] : 0.0;
        const double a0 = a[k * ld], a1 = p1 ? a[(k + 1) * ld] : 0.0, a2 = p2 ? a[(k + 2) * ld] : 0.0,
                     a3 = p3 ? a[(k + 3) * ld] : 0.0;
        a[k * ld] = fma(s, x0, a0);
        if (p1) a[(k + 1) * ld] = fma(s, x1, a1);
        if (p2) a[(k + 2) * ld] = fma(s, x2, a2);
        if (p3) a[(k + 3) * ld] = fma(s, x3, a3);
      }
    }
  }
  pair_barrier(barid);  // the neighbour's update is complete; nobody reads the source column any more
  if (role == 0) {
    if (lane == r0) xc[lane] = beta;
    else if (lane > r0 && lane < n) xc[lane] = 0.0;
  }
}

template <int NN, int PP>
__global__ void __launch_bounds__(512) rphess_pair32_kernel_t(Hess32Params P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp >> 1, role = warp & 1, barid = 1 + grp;
  const int n = NN ? NN : P.n, p = PP ? PP : P.p, ld = NN ? (NN % 2 == 0 ? NN + 1 : NN) : P.ld;
  const size_t nn = (size_t)n * n;
  const int fs = ld * n;  // doubles per staged factor
  double* S = psd_smem_hess + (size_t)grp * p * fs;
  __shared__ long long s_b[8];
  __shared__ int s_e[8];
  const int psize = pk_problem_size(n, p);
  for (;;) {
    if (role == 0 && lane == 0) s_b[grp] = (long long)atomicAdd(P.counter, 1ULL);
    pair_barrier(barid);
    const long long b = s_b[grp];
    if (b >= P.batch) break;
    const double* Ab = P.A + (size_t)b * p * nn;
    // each warp of the pair stages, normalises and later packs every other factor
    int escale = 0;
    for (int j = role; j < p; j += 2) {
      const double* src = Ab + (size_t)(P.left ? (p - 1 - j) : j) * nn;
      double* dst = S + j * fs;
      if (lane < n)
        for (int c = 0; c < n; c++) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + c * ld + lane);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + c * n + lane));
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    for (int j = role; j < p; j += 2) {
      double* dst = S + j * fs;
      double m = 0.0;
      if (lane < n)
        for (int c = 0; c < n; c++) m = fmax(m, fabs(dst[c * ld + lane]));
      m = warp_max(m);
      if (m > 0.0 && m < 1.7e308) {
        int e;
        (void)frexp(m, &e);
        if (e != 0) {
          const double sc = scalbn(1.0, (-e > 1000) ? 1000 : -e);
          if (lane < n)
            for (int c = 0; c < n; c++) dst[c * ld + lane] *= sc;
          escale += e;
        }
      }
    }
    if (role == 1 && lane == 0) s_e[grp] = escale;
    pair_barrier(barid);
    if (role == 0) escale += s_e[grp];
    for (int i = 0; i < n - 1; i++) {
      for (int j = p - 1; j >= 1; j--) hess32_step_pair(S + j * fs, S + (j - 1) * fs, n, ld, i, i, lane, role, barid);
      if (n - (i + 1) > 1) hess32_step_pair(S, S + (p - 1) * fs, n, ld, i + 1, i, lane, role, barid);
    }
    pair_barrier(barid);  // includes the last deferred column write of warp 0
    double* dstb = P.packed_out + (size_t)b * (psize + PK_STATE);
    if (role == 0 && lane == 0) dstb[psize] = (double)escale;
    for (int j = role; j < p; j += 2) {
      const int kl = (j == 0) ? 3 : 1;
      const int keep = (j == 0) ? 1 : 0;
      double* dst = dstb + ((j == 0) ? 0 : pk_size(3, n) + (j - 1) * pk_size(1, n));
      const double* src = S + j * fs;
      for (int c = 0; c < n; c++) {
        const int off = pk_off(kl, c);
        if (lane <= c + kl && lane < n) dst[off + lane] = (lane <= c + keep) ? src[c * ld + lane] : 0.0;
      }
    }
    // the barrier at the top of the loop keeps the next problem's staging behind this packing
  }
}

}  // namespace psd
