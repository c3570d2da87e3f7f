// FP64 tensor-core GEMM for the large-N blocked updates (compact-WY trailing updates of the
// periodic Hessenberg-triangular reduction, Schur-vector accumulation).
//
// These are the level-3 operations the reference performs one BLAS-1/2 call at a time
// (householder.jl:190-237: dot + axpy per column, gemv + ger per reflector); blocking p-fold
// dlahr2-style turns them into C <- alpha * op(A) * op(B) + beta * C.
//
// FP64 has no tcgen05 kind on sm_100a; the FP64 tensor path is the warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA).  CTA tile 128 x 64, k-tile 16, 8 warps as 4 x 2,
// warp tile 32 x 32 = 4 x 4 DMMA tiles (32 accumulator doubles per thread).  Operands are staged
// k-major in shared memory with leading dimensions = 4 (mod 16) doubles, which makes the
// fragment loads (8 rows x 4 k per half-warp) bank-conflict free; global loads of the next
// k-tile are issued into registers before the current tile is multiplied.  Optional split-K
// (grid.z) with atomicAdd for the tall-skinny inner products (V' * A, Q * V).
#pragma once
#include <cuda_runtime.h>

namespace psd {

struct GemmArgs {
  int M, N, K;
  const double* A;   // A(m,k) = A[m*rsA + k*csA]
  long long rsA, csA;
  const double* B;   // B(k,n) = B[k*rsB + n*csB]
  long long rsB, csB;
  double* C;         // column-major, leading dimension ldc
  long long ldc;
  double alpha, beta;
  int splitK;        // > 1: C must already hold beta*C; partial products are atomically added
};

constexpr int GM_BM = 128, GM_BN = 64, GM_BK = 16;
constexpr int GM_LDA = GM_BM + 4, GM_LDB = GM_BN + 4;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dgemm_dmma_kernel(GemmArgs g) {
  __shared__ double As[GM_BK][GM_LDA];
  __shared__ double Bs[GM_BK][GM_LDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int gq = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * GM_BN;
  // K range of this split
  const int ksz = (g.K + g.splitK - 1) / g.splitK;
  const int kbeg = blockIdx.z * ksz, kend = min(g.K, kbeg + ksz);
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // loader mappings: along the unit-stride dimension when there is one
  const bool a_mfast = (g.rsA == 1);  // consecutive m contiguous
  const bool b_nfast = (g.csB == 1);  // consecutive n contiguous
  double ra[8], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int m, k;
      if (a_mfast) { m = tid & 127; k = (tid >> 7) + 2 * i; }
      else { k = tid & 15; m = (tid >> 4) + 16 * i; }
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < kend) ? g.A[gm * g.rsA + gk * g.csA] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int n, k;
      if (b_nfast) { n = tid & 63; k = (tid >> 6) + 4 * i; }
      else { k = tid & 15; n = (tid >> 4) + 16 * i; }
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < g.N && gk < kend) ? g.B[gk * g.rsB + gn * g.csB] : 0.0;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int m, k;
      if (a_mfast) { m = tid & 127; k = (tid >> 7) + 2 * i; }
      else { k = tid & 15; m = (tid >> 4) + 16 * i; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int n, k;
      if (b_nfast) { n = tid & 63; k = (tid >> 6) + 4 * i; }
      else { k = tid & 15; n = (tid >> 4) + 16 * i; }
      Bs[k][n] = rb[i];
    }
  };

  if (kbeg < kend) load_tile(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += GM_BK) {
    __syncthreads();
    store_tile();
    __syncthreads();
    if (k0 + GM_BK < kend) load_tile(k0 + GM_BK);
#pragma unroll
    for (int ks = 0; ks < GM_BK; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[ks + tq][wm * 32 + i * 8 + gq];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[ks + tq][wn * 32 + j * 8 + gq];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // epilogue: C(row = gq, cols = 2*tq, 2*tq+1) of each 8x8 tile
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int m = m0 + wm * 32 + i * 8 + gq;
        const int n = n0 + wn * 32 + j * 8 + 2 * tq + e;
        if (m < g.M && n < g.N) {
          double* c = g.C + m + n * g.ldc;
          const double v = g.alpha * acc[i][j][e];
          if (g.splitK > 1)
            atomicAdd(c, v);
          else
            *c = (g.beta == 0.0) ? v : fma(g.beta, *c, v);
        }
      }
}

// C <- beta * C (used before a split-K accumulation when beta != 1)
__global__ void dscale_kernel(double* C, long long ldc, int M, int N, double beta) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * N;
       e += (long long)gridDim.x * blockDim.x) {
    double* c = C + (e % M) + (e / M) * ldc;
    *c = (beta == 0.0) ? 0.0 : beta * *c;
  }
}

// Enqueue C = alpha*A*B + beta*C.  Picks split-K when the output has too few tiles to fill the GPU.
inline cudaError_t dgemm_launch(cudaStream_t st, int sm_count, GemmArgs g) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  const int tm = (g.M + GM_BM - 1) / GM_BM, tn = (g.N + GM_BN - 1) / GM_BN;
  int split = 1;
  if (g.K >= 512 && tm * tn < sm_count) {
    split = (2 * sm_count + tm * tn - 1) / (tm * tn);
    split = max(1, min(split, g.K / 128));
  }
  g.splitK = split;
  if (g.K <= 0 || split > 1) {
    if (g.beta != 1.0) dscale_kernel<<<min(1024, (int)(((long long)g.M * g.N + 255) / 256)), 256, 0, st>>>(g.C, g.ldc, g.M, g.N, g.beta);
    if (g.K <= 0) return cudaGetLastError();
  }
  dim3 grid(tm, tn, split);
  dgemm_dmma_kernel<<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

}  // namespace psd
