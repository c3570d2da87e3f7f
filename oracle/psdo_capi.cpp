// ORACLE — TEST INFRASTRUCTURE ONLY (see psdo_common.hpp header).
// extern "C" surface of the CPU restatement, loaded by oracle/oracle.py via ctypes.
// OpenMP parallel-for over independent problems is the "reference threaded across the
// batch on the host cores" baseline of BASELINE.md §4 (C++ restatement, not Julia).
#include <omp.h>

#include "psdo_real.hpp"

using namespace psdo;

extern "C" {

int psdo_max_threads() { return omp_get_max_threads(); }

// Fill A[batch][p][n][n] (col-major) with uniform [0,1) entries for problems
// first_b .. first_b+batch-1 (part = 0).
void psdo_gen_real(uint64_t seed, int n, int p, int64_t batch, int64_t first_b, double* A) {
  size_t nn = (size_t)n * n;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; b++)
    for (int j = 0; j < p; j++)
      for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++)
          A[((size_t)b * p + j) * nn + (size_t)c * n + r] =
              gen_uniform(seed, (uint64_t)(first_b + b), j, r, c, 0);
}

// Complex: re = part 0, im = part 1; interleaved complex128.
void psdo_gen_complex(uint64_t seed, int n, int p, int64_t batch, int64_t first_b, double* A) {
  size_t nn = (size_t)n * n;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; b++)
    for (int j = 0; j < p; j++)
      for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) {
          size_t o = 2 * (((size_t)b * p + j) * nn + (size_t)c * n + r);
          A[o] = gen_uniform(seed, (uint64_t)(first_b + b), j, r, c, 0);
          A[o + 1] = gen_uniform(seed, (uint64_t)(first_b + b), j, r, c, 1);
        }
}

// Real standard periodic Schur, batched.  orientation 0 = :R, 1 = :L.
// A in/out [batch][p][n][n]; Z out [batch][p][n][n] or NULL; eig [batch][n] complex128;
// info [batch]; iters [batch] (total QR iterations, may be NULL).
int psdo_rpschur_batched(int n, int p, int64_t batch, int orientation, int wantT, int wantZ,
                         int maxitfac, double* A, double* Z, double* eig, int32_t* info,
                         int32_t* iters, int nthreads) {
  if (n < 1 || p < 1 || batch < 0) return -1;
  size_t nn = (size_t)n * n;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t b = 0; b < batch; b++) {
    RealQRStats st;
    info[b] = rpschur(n, p, A + (size_t)b * p * nn, (wantZ && Z) ? Z + (size_t)b * p * nn : nullptr,
                      eig + (size_t)b * 2 * n, orientation != 0, wantT != 0, wantZ != 0, maxitfac,
                      &st);
    if (iters) iters[b] = st.niter;
  }
  return 0;
}

// Reduction only (phessenberg! + explicit Q), for testing the reduction kernels:
// A in/out -> H factors (zeros below the Hessenberg/triangular structure), Q out.
int psdo_rphess_batched(int n, int p, int64_t batch, double* A, double* Q, int nthreads) {
  size_t nn = (size_t)n * n;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t b = 0; b < batch; b++) {
    std::vector<Mat> H(p);
    for (int j = 0; j < p; j++) H[j] = Mat{A + ((size_t)b * p + j) * nn, n};
    std::vector<std::vector<double>> tau;
    phessenberg(n, p, H, tau);
    if (Q)
      for (int j = 0; j < p; j++)
        form_q(n, H[j], tau[j], j == 0 ? 1 : 0, Mat{Q + ((size_t)b * p + j) * nn, n});
    for (int j = 0; j < p; j++) {
      int keep = (j == 0) ? 1 : 0;
      for (int c = 1; c <= n; c++)
        for (int r = c + keep + 1; r <= n; r++) H[j](r, c) = 0.0;
    }
  }
  return 0;
}

// Inner solver on already Hessenberg/triangular input (the reference's
// pschur!(H1, Hs; ...) entry, PeriodicSchurDecompositions.jl:322), rightwards form.
// Z in/out (identity-initialised here when wantZ).
int psdo_rpschur_hessut(int n, int p, double* H, double* Z, double* eig, int wantT, int wantZ,
                        int maxitfac) {
  size_t nn = (size_t)n * n;
  std::vector<Mat> Hm(p), Zm(p);
  for (int j = 0; j < p; j++) {
    Hm[j] = Mat{H + (size_t)j * nn, n};
    if (wantZ) {
      Zm[j] = Mat{Z + (size_t)j * nn, n};
      for (int c = 1; c <= n; c++)
        for (int r = 1; r <= n; r++) Zm[j](r, c) = (r == c) ? 1.0 : 0.0;
    }
  }
  std::vector<double> lre(n), lim(n);
  int info = real_periodic_qr(n, p, Hm, Zm, wantT != 0, wantZ != 0, maxitfac, lre.data(),
                              lim.data());
  for (int k = 0; k < n; k++) {
    eig[2 * k] = lre[k];
    eig[2 * k + 1] = lim[k];
  }
  return info;
}

}  // extern "C"
