// Blocked periodic Hessenberg-triangular reduction for large N (one problem on the whole GPU).
//
// Same mathematics as phessenberg! + explicit Q (PeriodicSchurDecompositions.jl:213-259,
// 136-140) - for each column i, factors p..2 get a QR-type reflector that is pushed into the
// right neighbour, then H_1 gets the Hessenberg reflector which is pushed into H_p - but
// organised as the p-fold analogue of LAPACK's dlahr2/dgehrd:
//
//   panel (nb columns, cooperative kernel, all SMs): for every column and every factor j the
//     current column is assembled lazily,  a = A_j[:,i] - Y_j V_{j+1}[i,:]'  (right transform by
//     the neighbour's reflectors generated so far) and  a <- (I - V_j T_j' V_j') a  (own earlier
//     reflectors); the new reflector v, its column of T_j and  Y_{j-1}[:,k] = tau (A_{j-1} v -
//     Y_{j-1} V_j' v)  follow.  The only O(n^2) work per reflector is the matrix-vector product
//     A_{j-1}[:, piv:] v, streamed from HBM by all CTAs (each owns a slab of rows for the whole
//     panel, so the row-parallel steps need no communication); two grid-wide barriers per
//     reflector carry the reductions (V_j' a; then |a|^2 and V_j' a below the pivot, from which
//     V_j' v follows without a third barrier).
//   trailing update (FP64 tensor-core GEMMs, psd_dgemm.cuh), per factor j:
//       A_j[:, c1:]   -= Y_j V_{j+1}[c1:, :]'
//       A_j[c0:, c1:] -= V_j (X_j' A_j[c0:, c1:]),   X_j = V_j T_j (formed by the panel kernel)
//       Q_j[:, c0:]   -= (Q_j[:, c0:] X_j) V_j'
//
// During the reduction the factors are kept TRANSPOSED in memory (A_j[r,c] at At[c + r*n]): the
// matrix-vector products then read each owned row as one contiguous run (sequential DRAM
// pages; with column-major storage a 28-row slab is a 224-byte fragment of every column and the
// products ran at ~30 % of the HBM rate), and every trailing update becomes an NN / NT GEMM
// with unit-stride operands.  The factors are transposed in place before and after.
//
// Reflector convention: dlarfg (householder.jl:66-108), H = I - tau w w', w = (1, v);
// compact WY: H_0 ... H_k = I - V T V', T upper triangular (dlarft, forward / columnwise).
// Sums of squares are formed without rescaling: the large-N path assumes entries whose squares
// are representable (|a| within 1e-150 .. 1e150).
#pragma once
#include <cooperative_groups.h>

#include "psd_device.cuh"
#include "psd_dgemm.cuh"

namespace psd {

namespace cg = cooperative_groups;

constexpr int LH_MAXP = 16;
constexpr int LH_THREADS = 512;
constexpr int LH_NW = LH_THREADS / 32;
constexpr int LH_RMAX = 64;  // rows per CTA slab (two 32-row chunks)

struct PanelParams {
  int n, p, c0, kb, nb;
  double* A[LH_MAXP];  // internal factor j at A[j-1], n x n, TRANSPOSED storage: A_j[r,c] at [c + r*n]
  double* V;           // [p][n*nb]
  double* Y;           // [p][n*nb]
  double* T;           // [p][nb*nb], upper triangular, zero-initialised per panel
  double* X;           // [p][n*nb]  X_j = V_j T_j (written at the end of the panel)
  double* abuf;        // [n]
  double* wbuf;        // [2][nb] (zero on entry)
  double* zbuf;        // [2][nb] (zero on entry)
  double* sc;          // [3]: sum of squares (two parities, zero on entry), alpha
  int R;               // rows per slab
  long long* prof;     // [8] cycle counters of CTA 0 (phase breakdown), or nullptr
};

// partial[r] over the warps of the CTA -> out[r], r < 64 (sred is [LH_NW][64])
__device__ __forceinline__ void lh_reduce_rows(double (*sred)[LH_RMAX], double v0, double v1, int warp, int tx) {
  sred[warp][tx] = v0;
  sred[warp][tx + 32] = v1;
  __syncthreads();
}
__device__ __forceinline__ double lh_sum_rows(double (*sred)[LH_RMAX], int r) {
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < LH_NW; w++) s += sred[w][r];
  return s;
}

extern __shared__ __align__(16) double lh_smem[];

__global__ void __launch_bounds__(LH_THREADS, 1) rphess_panel_kernel(PanelParams P) {
  cg::grid_group grid = cg::this_grid();
  const int n = P.n, p = P.p, nb = P.nb, kb = P.kb, c0 = P.c0;
  const int tid = threadIdx.x, tx = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * P.R, r1 = min(n, r0 + P.R);
  const bool have = r0 < n;
  double* Ts = lh_smem;  // [p][nb][nb], T(t,q) at Ts[(j-1)*nb*nb + t + q*nb]
  double* vsm = lh_smem + (size_t)p * nb * nb;  // [n] the reflector vector of the current step
  __shared__ double sred[LH_NW][LH_RMAX];
  __shared__ double a_s[LH_RMAX], y_s[LH_RMAX], u_s[64], z_s[64], w_s[64];
  for (int e = tid; e < p * nb * nb; e += LH_THREADS) Ts[e] = 0.0;
  __syncthreads();
  const size_t nnb = (size_t)n * nb;
  // reduction buffers, double-buffered by the parity of the step counter:
  //   wbuf[par][nb]  w = V_j' a           (accumulated before S2, read after it)
  //   zbuf[par][nb]  S = V_j[piv+1:,:]' a (accumulated before S3, read after it)
  //   sc[par]        |a[piv+1:]|^2        (likewise);  sc[2] = alpha
  int step = 0;
  long long tacc[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = P.prof && blockIdx.x == 0 && tid == 0;
  long long tl = prof ? clock64() : 0;
#define LH_TICK(slot)                 \
  if (prof) {                         \
    const long long tn_ = clock64();  \
    tacc[slot] += tn_ - tl;           \
    tl = tn_;                         \
  }

  for (int k = 0; k < kb; k++) {
    const int i = c0 + k;
    for (int jj = 0; jj < p; jj++, step++) {
      const int par = step & 1;
      double* wb = P.wbuf + par * nb;
      double* zb = P.zbuf + par * nb;
      const int j = (jj == p - 1) ? 1 : p - jj;
      const int piv = (j == 1) ? i + 1 : i;
      const int jn = (j == p) ? 1 : j + 1;
      const int kk = (j == p) ? k : k + 1;
      const int jm = (j == 1) ? p : j - 1;
      double* Aj = P.A[j - 1];
      const double* Vn = P.V + (size_t)(jn - 1) * nnb;
      double* Vj = P.V + (size_t)(j - 1) * nnb;
      const double* Yj = P.Y + (size_t)(j - 1) * nnb;
      double* Ym = P.Y + (size_t)(jm - 1) * nnb;
      const double* Am = P.A[jm - 1];
      double* Tj = Ts + (size_t)(j - 1) * nb * nb;
      const bool ra = have && (r0 + tx < r1), rb = have && (r0 + tx + 32 < r1);

      // ---- s1: a = A_j[:, i] - Y_j[:, 0:kk] V_jn[i, 0:kk]'  (own rows) ----
      // The newest column of V_jn was written by other CTAs since the last grid barrier, but
      // row i is its pivot row, so the value is known to be 1 without reading it: column k of
      // V_{j+1} (pivot i) for j < p, column k-1 of V_1 (pivot (i-1)+1) for j == p.
      {
        const int qnew = kk - 1;
        double p0 = 0.0, p1 = 0.0;
#pragma unroll 4
        for (int q = warp; q < kk; q += LH_NW) {
          const double vq = (q == qnew) ? 1.0 : Vn[i + (size_t)q * n];
          if (ra) p0 = fma(Yj[r0 + tx + (size_t)q * n], vq, p0);
          if (rb) p1 = fma(Yj[r0 + tx + 32 + (size_t)q * n], vq, p1);
        }
        lh_reduce_rows(sred, p0, p1, warp, tx);
        if (tid < LH_RMAX) {
          const int r = r0 + tid;
          a_s[tid] = (have && r < r1) ? Aj[i + (size_t)r * n] - lh_sum_rows(sred, tid) : 0.0;
        }
        __syncthreads();
      }
      // ---- s2: w = V_j[:, 0:k]' a  (partial over own rows) ----
      if (have && r1 > c0) {
#pragma unroll 4
        for (int q = warp; q < k; q += LH_NW) {
          double d = 0.0;
          if (ra) d = Vj[r0 + tx + (size_t)q * n] * a_s[tx];
          if (rb) d = fma(Vj[r0 + tx + 32 + (size_t)q * n], a_s[tx + 32], d);
          d = warp_sum(d);
          if (tx == 0 && d != 0.0) atomicAdd(&wb[q], d);
        }
      }
      LH_TICK(0)
      grid.sync();  // S2
      LH_TICK(1)
      // ---- s3: a <- a - V_j (T_j' w); publish a; |a[piv+1:]|^2, alpha, S = V_j[piv+1:,:]' a ----
      if (tid < k) w_s[tid] = wb[tid];
      __syncthreads();
      if (tid < k) {
        double u = 0.0;
        for (int t = 0; t <= tid; t++) u = fma(Tj[t + tid * nb], w_s[t], u);
        u_s[tid] = u;
      }
      if (blockIdx.x == 0 && tid < nb) {  // buffers of the next step (other parity)
        P.wbuf[(par ^ 1) * nb + tid] = 0.0;
        P.zbuf[(par ^ 1) * nb + tid] = 0.0;
        if (tid == 0) P.sc[par ^ 1] = 0.0;
      }
      __syncthreads();
      {
        double p0 = 0.0, p1 = 0.0;
        if (have && r1 > c0) {
  #pragma unroll 4
        for (int q = warp; q < k; q += LH_NW) {
            const double uq = u_s[q];
            if (ra) p0 = fma(Vj[r0 + tx + (size_t)q * n], uq, p0);
            if (rb) p1 = fma(Vj[r0 + tx + 32 + (size_t)q * n], uq, p1);
          }
        }
        lh_reduce_rows(sred, p0, p1, warp, tx);
        if (tid < LH_RMAX) {
          const int r = r0 + tid;
          double a = 0.0;
          if (have && r < r1) {
            a = a_s[tid] - lh_sum_rows(sred, tid);
            P.abuf[r] = a;
            if (r == piv) P.sc[2] = a;
          }
          a_s[tid] = a;
          double sq = (have && r < r1 && r > piv) ? a * a : 0.0;
          sq = warp_sum(sq);
          if (tx == 0 && sq != 0.0) atomicAdd(&P.sc[par], sq);
        }
        __syncthreads();
        if (have && r1 > piv + 1) {
  #pragma unroll 4
        for (int q = warp; q < k; q += LH_NW) {
            double d = 0.0;
            if (ra && r0 + tx > piv) d = Vj[r0 + tx + (size_t)q * n] * a_s[tx];
            if (rb && r0 + tx + 32 > piv) d = fma(Vj[r0 + tx + 32 + (size_t)q * n], a_s[tx + 32], d);
            d = warp_sum(d);
            if (tx == 0 && d != 0.0) atomicAdd(&zb[q], d);
          }
        }
      }
      LH_TICK(2)
      grid.sync();  // S3
      LH_TICK(3)
      // ---- s4: reflector; column i of A_j, V_j[:,k]; z = V_j' v; y = A_m[:, piv:] v;
      //          Y_m[:, k] = tau (y - Y_m[:, 0:k] z);  T_j[0:k, k] = -tau T_j[0:k,0:k] z ----
      const double ssq = P.sc[par], alpha = P.sc[2];
      double beta = alpha, tau = 0.0, scl = 0.0;
      if (ssq > 0.0) {
        beta = -copysign(sqrt(fma(alpha, alpha, ssq)), alpha);
        tau = (beta - alpha) / beta;
        scl = 1.0 / (alpha - beta);
      }
      if (tid < LH_RMAX) {
        const int r = r0 + tid;
        if (have && r < r1) {
          Vj[r + (size_t)k * n] = (r < piv) ? 0.0 : (r == piv) ? 1.0 : a_s[tid] * scl;
          Aj[i + (size_t)r * n] = (r < piv) ? a_s[tid] : (r == piv) ? beta : 0.0;
        }
      }
      // z[q] = V_j[piv, q] + scl * S[q]   (row piv of V_j is at least one barrier old)
      if (tid < k) z_s[tid] = fma(scl, zb[tid], Vj[piv + (size_t)tid * n]);
      {
        // y = A_m[own rows, piv:] v.  The vector goes to shared memory once; work items are
        // (row, 1024-column chunk) pairs dealt round-robin to the warps, 8 coalesced 256-byte
        // loads in flight per warp.
        if (tid < LH_RMAX) y_s[tid] = 0.0;
        if (tau != 0.0)
          for (int c = piv + tid; c < n; c += LH_THREADS) vsm[c] = (c == piv) ? 1.0 : P.abuf[c] * scl;
        __syncthreads();
        if (have && tau != 0.0) {
          const int nrow = r1 - r0;
          const int nch = (n - piv + 1023) >> 10;
          for (int it = warp; it < nrow * nch; it += LH_NW) {
            const int rr = it % nrow, ch = it / nrow;
            const double* arow = Am + (size_t)(r0 + rr) * n;
            const int cb = piv + (ch << 10), ce = min(n, cb + 1024);
            double acc = 0.0;
            int c = cb + tx;
            for (; c + 7 * 32 < ce; c += 8 * 32) {
              double m[8];
#pragma unroll
              for (int t = 0; t < 8; t++) m[t] = arow[c + t * 32];
#pragma unroll
              for (int t = 0; t < 8; t++) acc = fma(m[t], vsm[c + t * 32], acc);
            }
            for (; c < ce; c += 32) acc = fma(arow[c], vsm[c], acc);
            acc = warp_sum(acc);
            if (tx == 0) atomicAdd(&y_s[rr], acc);
          }
        }
        __syncthreads();
        LH_TICK(4)
      }
      {
        double p0 = 0.0, p1 = 0.0;
        if (have) {
  #pragma unroll 4
        for (int q = warp; q < k; q += LH_NW) {
            const double zq = z_s[q];
            if (ra) p0 = fma(Ym[r0 + tx + (size_t)q * n], zq, p0);
            if (rb) p1 = fma(Ym[r0 + tx + 32 + (size_t)q * n], zq, p1);
          }
        }
        lh_reduce_rows(sred, p0, p1, warp, tx);
        if (tid < LH_RMAX) {
          const int r = r0 + tid;
          if (have && r < r1) Ym[r + (size_t)k * n] = tau * (y_s[tid] - lh_sum_rows(sred, tid));
        }
        if (tid >= 64 && tid < 64 + k + 1) {
          const int t = tid - 64;
          double val;
          if (t == k) {
            val = tau;
          } else {
            double sacc = 0.0;
            for (int q = t; q < k; q++) sacc = fma(Tj[t + q * nb], z_s[q], sacc);
            val = -tau * sacc;
          }
          Tj[t + k * nb] = val;
          if (blockIdx.x == 0) P.T[(size_t)(j - 1) * nb * nb + t + (size_t)k * nb] = val;
        }
        __syncthreads();
        LH_TICK(5)
      }
    }
  }
  // X_j = V_j T_j for the own rows: the trailing updates then need no multiplication by T
  if (have) {
    const int nrow = r1 - r0;
    for (int e = tid; e < p * nrow * kb; e += LH_THREADS) {
      const int j = e / (nrow * kb), rem = e % (nrow * kb);
      const int q = rem / nrow, r = r0 + rem % nrow;
      const double* Vj = P.V + (size_t)j * nnb;
      const double* Tj = Ts + (size_t)j * nb * nb;
      double acc = 0.0;
      for (int t = 0; t <= q; t++) acc = fma(Vj[r + (size_t)t * n], Tj[t + q * nb], acc);
      P.X[(size_t)j * nnb + r + (size_t)q * n] = acc;
    }
  }
  if (prof)
    for (int t = 0; t < 6; t++) atomicAdd((unsigned long long*)&P.prof[t], (unsigned long long)tacc[t]);
#undef LH_TICK
}

__global__ void lh_identity_kernel(double* Q, int n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)n * n;
       e += (long long)gridDim.x * blockDim.x)
    Q[e] = ((e % n) == (e / n)) ? 1.0 : 0.0;
}

// In-place transpose of an n x n matrix: tile (bx, by) with bx >= by swaps with its mirror.
__global__ void lh_transpose_kernel(double* A, int n) {
  __shared__ double t0[32][33], t1[32][33];
  const int bx = blockIdx.x, by = blockIdx.y;
  if (bx < by) return;
  const int x = threadIdx.x;
  for (int yy = threadIdx.y; yy < 32; yy += 8) {
    const int r = by * 32 + x, c = bx * 32 + yy;  // tile (rows by, cols bx): element [r + c*n]
    if (r < n && c < n) t0[yy][x] = A[r + (size_t)c * n];
    const int r2 = bx * 32 + x, c2 = by * 32 + yy;  // mirror tile
    if (r2 < n && c2 < n) t1[yy][x] = A[r2 + (size_t)c2 * n];
  }
  __syncthreads();
  for (int yy = threadIdx.y; yy < 32; yy += 8) {
    // A[r, c] <- old A[c, r]
    const int r = by * 32 + x, c = bx * 32 + yy;
    if (r < n && c < n) A[r + (size_t)c * n] = t1[x][yy];
    if (bx != by) {
      const int r2 = bx * 32 + x, c2 = by * 32 + yy;
      if (r2 < n && c2 < n) A[r2 + (size_t)c2 * n] = t0[x][yy];
    }
  }
}

// Workspace of the large-N reduction (device memory, owned by the caller).
struct LargeHessWork {
  double *V = nullptr, *Y = nullptr, *T = nullptr, *W = nullptr, *W2 = nullptr, *small = nullptr;
  size_t cap = 0;
};

inline int lh_panel_width(int p) {
  int nb = 64;
  while (nb > 8 && (size_t)p * nb * nb * sizeof(double) > 160 * 1024) nb -= 8;
  return nb;
}
inline size_t lh_work_doubles(int n, int p) {
  const int nb = lh_panel_width(p);
  return 3 * (size_t)p * n * nb + (size_t)p * nb * nb + 2 * (size_t)n * nb + n + 4 * nb + 8;
}

// Reduce the p factors A[0..p-1] (internal rightwards order, device pointers) in place and, when
// Q != nullptr, accumulate the explicit Q_j (initialised to the identity here).  `flops`
// receives the GEMM flops issued (for the roofline report).
// `mark(kind, phase)` is called around the panel kernel (kind 2) and around each group of GEMMs
// (kind 3) with phase 0 = begin, 1 = end (kernel-level timing hooks of the caller).
template <class Mark>
inline cudaError_t rphess_large(cudaStream_t st, int sm_count, int n, int p, double* const* A, double* const* Q,
                                double* work, double* gemm_flops, Mark&& mark,
                                long long* prof_cycles = nullptr) {
  if (p > LH_MAXP) return cudaErrorInvalidValue;
  const int nb = lh_panel_width(p);
  double* V = work;
  double* Y = V + (size_t)p * n * nb;
  double* X = Y + (size_t)p * n * nb;
  double* T = X + (size_t)p * n * nb;
  double* W = T + (size_t)p * nb * nb;
  double* W2 = W + (size_t)n * nb;
  double* small = W2 + (size_t)n * nb;  // abuf[n], wbuf[nb], zbuf[nb], sc[2]
  cudaError_t e;
  if (Q)
    for (int j = 0; j < p; j++) lh_identity_kernel<<<sm_count * 4, 256, 0, st>>>(Q[j], n);
  const size_t smem = ((size_t)p * nb * nb + n) * sizeof(double);
  for (int j = 0; j < p; j++) lh_transpose_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, st>>>(A[j], n);
  if ((e = cudaFuncSetAttribute(rphess_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
      cudaSuccess)
    return e;
  int G = sm_count;
  int R = (n + G - 1) / G;
  R = (R + 3) & ~3;
  if (R > LH_RMAX) return cudaErrorInvalidValue;  // n too large for this slab scheme
  double fl = 0.0;
  for (int c0 = 0; c0 < n - 1; c0 += nb) {
    const int kb = std::min(nb, n - 1 - c0);
    const int c1 = c0 + kb;
    if ((e = cudaMemsetAsync(T, 0, (size_t)p * nb * nb * sizeof(double), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(small + n, 0, (4 * nb + 4) * sizeof(double), st)) != cudaSuccess) return e;
    PanelParams P;
    P.n = n; P.p = p; P.c0 = c0; P.kb = kb; P.nb = nb;
    for (int j = 0; j < p; j++) P.A[j] = A[j];
    P.V = V; P.Y = Y; P.T = T; P.X = X;
    P.abuf = small; P.wbuf = small + n; P.zbuf = small + n + 2 * nb; P.sc = small + n + 4 * nb;
    P.R = R;
    P.prof = prof_cycles;
    void* args[] = {&P};
    mark(2, 0);
    e = cudaLaunchCooperativeKernel((void*)rphess_panel_kernel, dim3(G), dim3(LH_THREADS), args, smem, st);
    mark(2, 1);
    if (e != cudaSuccess) return e;
    const int ncols = n - c1;
    mark(3, 0);
    for (int j = 1; j <= p; j++) {
      const int jn = (j == p) ? 1 : j + 1;
      double* Aj = A[j - 1];
      const double* Vj = V + (size_t)(j - 1) * n * nb;
      const double* Vn = V + (size_t)(jn - 1) * n * nb;
      const double* Yj = Y + (size_t)(j - 1) * n * nb;
      const double* Xj = X + (size_t)(j - 1) * n * nb;
      GemmArgs g;
      if (ncols > 0) {
        // transposed storage: At = A_j', an n x n column-major array with At[c, r] = A_j[r, c]
        // (1) A_j[:, c1:] -= Y_j V_jn[c1:, :]'        <=>  At[c1:, :] -= V_jn[c1:, :] Y_j'
        g = GemmArgs{ncols, n, kb, Vn + c1, 1, n, Yj, n, 1, Aj + c1, n, -1.0, 1.0, 1};
        if ((e = dgemm_launch(st, sm_count, g)) != cudaSuccess) return e;
        // (2) W2 = T_j' V_j[c0:, :]' A_j[c0:, c1:]    <=>  W2t = At[c1:, c0:] X_j[c0:, :]
        g = GemmArgs{ncols, kb, n - c0, Aj + c1 + (size_t)c0 * n, 1, n, Xj + c0, 1, n, W2, n, 1.0, 0.0, 1};
        if ((e = dgemm_launch(st, sm_count, g)) != cudaSuccess) return e;
        // (3) A_j[c0:, c1:] -= V_j[c0:, :] W2         <=>  At[c1:, c0:] -= W2t V_j[c0:, :]'
        g = GemmArgs{ncols, n - c0, kb, W2, 1, n, Vj + c0, n, 1, Aj + c1 + (size_t)c0 * n, n, -1.0, 1.0, 1};
        if ((e = dgemm_launch(st, sm_count, g)) != cudaSuccess) return e;
        fl += 2.0 * kb * ncols * ((double)n + 2.0 * (n - c0));
      }
      if (Q) {
        double* Qj = Q[j - 1];
        // (4) W2 = Q_j[:, c0:] X_j[c0:, :]   (n x kb, leading dimension n)
        g = GemmArgs{n, kb, n - c0, Qj + (size_t)c0 * n, 1, n, Xj + c0, 1, n, W2, n, 1.0, 0.0, 1};
        if ((e = dgemm_launch(st, sm_count, g)) != cudaSuccess) return e;
        // (5) Q_j[:, c0:] -= W2 V_j[c0:, :]'
        g = GemmArgs{n, n - c0, kb, W2, 1, n, Vj + c0, n, 1, Qj + (size_t)c0 * n, n, -1.0, 1.0, 1};
        if ((e = dgemm_launch(st, sm_count, g)) != cudaSuccess) return e;
        fl += 2.0 * kb * n * (2.0 * (n - c0));
      }
    }
    mark(3, 1);
  }
  for (int j = 0; j < p; j++) lh_transpose_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, st>>>(A[j], n);
  if (gemm_flops) *gemm_flops = fl;
  return cudaGetLastError();
}

}  // namespace psd
