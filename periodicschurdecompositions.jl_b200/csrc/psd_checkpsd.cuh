// Device version of the reference's integrity check of a real periodic Schur decomposition
// (checkpsd, diagnostics.jl:190-263, S all true): for every problem b and factor l
//   tri   = || part of T_l below its (quasi-)triangle ||_F      (tril(T, -2) for l = schurindex, else -1)
//   orth  = || Z_l Z_l' - I ||_F
//   err   = || Z_l T_l Z_l1' - A_l ||_F / (eps ||A_l||_1)   (:R;  Z_l1 T_l Z_l' for :L), l1 = l mod p + 1
// One CTA works on a chunk of columns of one (problem, factor): column c of the product is
// Za (T (row c of Zb)'), two matrix-vector products with the matrices streamed column by column
// (coalesced along the rows, one thread per row strip), CB columns at a time so that every matrix
// column that is read serves CB right-hand sides.  Partial sums of squares are added atomically.
// The norms land in out[b][l][4] = { err2 (sum of squares), ||A_l||_1, tri2, orth2 }; the host
// finishes (square roots, eps, thresholds).
#pragma once
#include <cuda_runtime.h>

namespace psd {

struct CheckParams {
  int n, p, left, cols_per_cta;
  long long batch;
  const double* A;  // [batch][p][n*n] storage layout, user factor order
  const double* T;
  const double* Z;
  double* out;      // [batch][p][4], zero on entry
};

extern __shared__ __align__(16) double chk_smem[];

template <int CB>
__device__ __forceinline__ void chk_matvec(const double* Mx, int n, const double* v, double* w, int tid, int nt) {
  // w[r][q] = sum_k M(r, k) v[k][q]   (vectors stored interleaved: index * CB + q)
  for (int r = tid; r < n; r += nt) {
    double acc[CB];
#pragma unroll
    for (int q = 0; q < CB; q++) acc[q] = 0.0;
    for (int k = 0; k < n; k++) {
      const double m = Mx[r + (size_t)k * n];
#pragma unroll
      for (int q = 0; q < CB; q++) acc[q] = fma(m, v[k * CB + q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < CB; q++) w[r * CB + q] = acc[q];
  }
}

__device__ __forceinline__ double chk_block_sum(double v, double* red, int tid, int nt) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (tid == 0)
    for (int w = 0; w < (nt + 31) / 32; w++) s += red[w];
  return s;  // valid in thread 0
}

__device__ __forceinline__ void chk_atomic_max_nonneg(double* addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template <int CB>
__global__ void __launch_bounds__(256) checkpsd_kernel(CheckParams P) {
  const int n = P.n, p = P.p, tid = threadIdx.x, nt = blockDim.x;
  const int l = blockIdx.x;        // factor, 0-based user order
  const long long b = blockIdx.y;  // problem
  const int c0 = blockIdx.z * P.cols_per_cta, c1 = min(n, c0 + P.cols_per_cta);
  const size_t nn = (size_t)n * n;
  const double* Al = P.A + ((size_t)b * p + l) * nn;
  const double* Tl = P.T + ((size_t)b * p + l) * nn;
  const double* Zl = P.Z + ((size_t)b * p + l) * nn;
  const double* Zn = P.Z + ((size_t)b * p + (l + 1) % p) * nn;
  const double* Za = P.left ? Zn : Zl;  // A_l = Za T_l Zb'
  const double* Zb = P.left ? Zl : Zn;
  double* out = P.out + ((size_t)b * p + l) * 4;
  double* v = chk_smem;               // [n][CB]
  double* w = chk_smem + (size_t)n * CB;
  __shared__ double red[8];
  const int js = P.left ? p - 1 : 0;
  double err2 = 0.0, orth2 = 0.0, tri2 = 0.0, a1 = 0.0;
  for (int cb = c0; cb < c1; cb += CB) {
    const int nc = min(CB, c1 - cb);
    // ---- residual: x = Za (T (row c of Zb)') ----
    for (int e = tid; e < n * CB; e += nt) {
      const int k = e / CB, q = e % CB;
      v[e] = (q < nc) ? Zb[(cb + q) + (size_t)k * n] : 0.0;
    }
    __syncthreads();
    chk_matvec<CB>(Tl, n, v, w, tid, nt);
    __syncthreads();
    chk_matvec<CB>(Za, n, w, v, tid, nt);
    __syncthreads();
    for (int q = 0; q < nc; q++) {
      double colabs = 0.0;
      for (int r = tid; r < n; r += nt) {
        const double a = Al[r + (size_t)(cb + q) * n];
        const double d = v[r * CB + q] - a;
        err2 = fma(d, d, err2);
        colabs += fabs(a);
        // structure of T_l: entries below the (quasi-)triangle
        if (r > (cb + q) + ((l == js) ? 1 : 0)) {
          const double t = Tl[r + (size_t)(cb + q) * n];
          tri2 = fma(t, t, tri2);
        }
      }
      const double cs = chk_block_sum(colabs, red, tid, nt);
      if (tid == 0) a1 = fmax(a1, cs);
    }
    __syncthreads();
    // ---- orthogonality: y = Z_l (row c of Z_l)' - e_c ----
    for (int e = tid; e < n * CB; e += nt) {
      const int k = e / CB, q = e % CB;
      w[e] = (q < nc) ? Zl[(cb + q) + (size_t)k * n] : 0.0;
    }
    __syncthreads();
    chk_matvec<CB>(Zl, n, w, v, tid, nt);
    __syncthreads();
    for (int q = 0; q < nc; q++)
      for (int r = tid; r < n; r += nt) {
        const double d = v[r * CB + q] - ((r == cb + q) ? 1.0 : 0.0);
        orth2 = fma(d, d, orth2);
      }
    __syncthreads();
  }
  const double e2 = chk_block_sum(err2, red, tid, nt);
  const double o2 = chk_block_sum(orth2, red, tid, nt);
  const double t2 = chk_block_sum(tri2, red, tid, nt);
  if (tid == 0) {
    atomicAdd(&out[0], e2);
    chk_atomic_max_nonneg(&out[1], a1);
    atomicAdd(&out[2], t2);
    atomicAdd(&out[3], o2);
  }
}

}  // namespace psd
