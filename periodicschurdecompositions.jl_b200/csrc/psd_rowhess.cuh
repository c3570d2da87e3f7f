// Row-wise periodic Hessenberg reduction for the left orientation (Krylov restarts).
//
// Replaces _rphessenberg!(Ap, A, Q) (rhessx.jl:53-109) and the RHouseholder lmul!/rmul!
// (rhessx.jl:7-50): for i = n..2 and l = p-1..1 a reflector built from the reversed conjugated
// row A_l[i, i:-1:1] annihilates the leading i-1 entries of that row from the right
// (A_l <- A_l H', Q_l <- Q_l H') and is applied from the left to the first i rows of the
// neighbour (A_{l-1}, or Ap for l = 1); then row i of Ap is reduced the same way against
// A_{p-1} / Q_p.  Ap may carry one extra row (the Arnoldi foot), which is treated first
// (:67-79).  One CTA per problem, matrices in place in global memory.
#pragma once
#include "psd_gen_common.cuh"

namespace psd {

template <class T>
struct RowHessParams {
  int n, m, p, qrows;   // Ap is m x n (m = n or n+1); A_l n x n, l = 1..p-1; Q_l qrows x n, l = 1..p
  long long batch;
  T* Ap;                // [batch][m*n]
  T* A;                 // [batch][p-1][n*n]
  T* Q;                 // [batch][p][qrows*n] in/out, or nullptr
};

// One reflector: source row `k` of X (leading dimension ldx), pivot column kc, acting on columns
// 1..kc.  Right targets: X itself (rows 1..xrows except k), Qm (qrows rows).  Left target: L
// (rows 1..kc, ncolL columns).
template <class T>
PSD_DEV void rowhess_step(T* X, int ldx, int xrows, int k, int kc, T* Qm, int ldq, int qrows, T* L,
                          int ldl, int ncolL, int tid, int nt) {
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int m = kc;
  if (m < 1) return;
  const T* x = &PSD_GE(X, ldx, k, kc);
  const long long inc = -(long long)ldx;
  double beta;
  T tau, tv;
  if (!refl_vec<T, true>(x, inc, m, lane, beta, tau, tv)) return;
  extern __shared__ __align__(16) double rowhess_smem[];
  T* w = reinterpret_cast<T*>(rowhess_smem);  // w_0 = 1 (column kc), w_r = tv conj(x_r) (column kc-r)
  for (int r = tid; r < m; r += nt) w[r] = (r == 0) ? Scalar<T>::one() : tv * conj_(x[r * inc]);
  __syncthreads();
  const int nX = xrows, nQ = Qm ? qrows : 0;
  for (int c = tid; c < nX + nQ; c += nt) {
    if (c < nX) {
      if (c + 1 == k) continue;  // the source row is finalised below
      hh_right_one<T>(&PSD_GE(X, ldx, 1 + c, kc), -(long long)ldx, m, w, tau);
    } else {
      hh_right_one<T>(&PSD_GE(Qm, ldq, 1 + (c - nX), kc), -(long long)ldq, m, w, tau);
    }
  }
  if (L == X) __syncthreads();  // p == 1: both sides act on Ap, one after the other
  if (L)
    for (int c = warp; c < ncolL; c += nw) hh_left_warp<T>(&PSD_GE(L, ldl, kc, 1 + c), -1, m, w, tau, lane);
  __syncthreads();
  for (int r = tid; r < m; r += nt)
    PSD_GE(X, ldx, k, kc - r) = (r == 0) ? Scalar<T>::from_real(beta) : Scalar<T>::zero();
  __syncthreads();
}

template <class T>
__global__ void rowhess_kernel(RowHessParams<T> P) {
  const int n = P.n, m = P.m, p = P.p, tid = threadIdx.x, nt = blockDim.x;
  const size_t nn = (size_t)n * n;
  for (long long b = blockIdx.x; b < P.batch; b += gridDim.x) {
    T* Ap = P.Ap + (size_t)b * m * n;
    T* Ab = P.A + (size_t)b * (p - 1) * nn;
    T* Qb = P.Q ? P.Q + (size_t)b * p * (size_t)P.qrows * n : nullptr;
    auto Af = [&](int l) { return Ab + (size_t)(l - 1) * nn; };           // A_l, l = 1..p-1
    auto Qf = [&](int l) { return Qb ? Qb + (size_t)(l - 1) * P.qrows * n : (T*)nullptr; };
    // left neighbour of the Hessenberg factor: A_{p-1}, or Ap itself when p == 1
    T* Ax = (p == 1) ? Ap : Af(p - 1);
    const int ldax = (p == 1) ? m : n;
    if (m == n + 1)  // the extra (foot) row first (:67-79)
      rowhess_step<T>(Ap, m, m, n + 1, n, Qf(p), P.qrows, P.qrows, Ax, ldax, n, tid, nt);
    for (int i = n; i >= 2; i--) {
      for (int l = p - 1; l >= 1; l--) {
        T* L = (l == 1) ? Ap : Af(l - 1);
        rowhess_step<T>(Af(l), n, n, i, i, Qf(l), P.qrows, P.qrows, L, (l == 1) ? m : n, n, tid, nt);
      }
      rowhess_step<T>(Ap, m, m, i, i - 1, Qf(p), P.qrows, P.qrows, Ax, ldax, n, tid, nt);
    }
    __syncthreads();
  }
}

}  // namespace psd
