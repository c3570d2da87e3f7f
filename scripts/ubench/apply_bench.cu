// Micro-benchmark and cross-check of the window-update kernels in isolation: one synthetic round
// of the N = 4096, p = 4 pipeline (windows of order 56, T and Z wanted), timed with CUDA events.
// Checks the near / far split of a round against the unsplit update (bitwise).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
//        -I periodicschurdecompositions.jl_b200/csrc scripts/ubench/apply_bench.cu -o scripts/ubench/apply_bench
// Usage: apply_bench [n] [windows]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "psd_ms_kernels.cuh"

using namespace psd::ms;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Setup {
  int n = 4096, p = 4, W = 56, D = 28, cnt = 13;
  std::vector<double*> H, Z;
  double* U = nullptr;
  WinDesc* wins = nullptr;
  std::vector<double> h0;
  double flops = 0, bytes = 0;
};

static void launch_v1(const Setup& S, ApplyParams A, int tpb, int part, int phases) {
  const int tiles = (S.n + AP_T - 1) / AP_T;
  A.tpb = (part == 1) ? 1 : tpb;
  A.part = part;
  const int ch0 = (part == 1) ? 2 : (tiles + A.tpb - 1) / A.tpb, ch1 = (part == 1) ? 1 : ch0;
  if (phases & 1) { A.phase = 0; ms_apply_kernel<<<dim3(ch0, S.cnt * S.p * 2), 256, AP_SMEM>>>(A); }
  if (phases & 2) { A.phase = 1; ms_apply_kernel<<<dim3(ch1, S.cnt * S.p), 256, AP_SMEM>>>(A); }
}
template <class F>
static float time_rounds(F&& launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

static void reset(Setup& S) {
  const size_t nn = (size_t)S.n * S.n;
  for (int j = 0; j < S.p; j++) {
    CK(cudaMemcpy(S.H[j], S.h0.data(), nn * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S.Z[j], S.h0.data(), nn * 8, cudaMemcpyHostToDevice));
  }
}
static std::vector<double> snapshot(Setup& S) {
  const size_t nn = (size_t)S.n * S.n;
  std::vector<double> out(2 * S.p * nn);
  for (int j = 0; j < S.p; j++) {
    CK(cudaMemcpy(out.data() + (2 * j) * nn, S.H[j], nn * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data() + (2 * j + 1) * nn, S.Z[j], nn * 8, cudaMemcpyDeviceToHost));
  }
  return out;
}
static double maxdiff(const std::vector<double>& a, const std::vector<double>& b) {
  double d = 0;
  for (size_t i = 0; i < a.size(); i++) {
    const double e = std::fabs(a[i] - b[i]);
    if (!(e <= d)) d = e;  // NaN propagates
  }
  return d;
}

// windows: cnt of them spread over the matrix; odd = 1 shifts every other window to an odd start
// and shortens the last one to a length that is neither even nor a multiple of 4
static void make_windows(Setup& S, int odd) {
  const int n = S.n, p = S.p, W = S.W;
  std::vector<double> u((size_t)p * n * W, 0.0);
  std::vector<WinDesc> wd(S.cnt);
  unsigned long long seed = 12345;
  auto rnd = [&] { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; return (double)((seed >> 33) % 2001) / 1000.0 - 1.0; };
  S.flops = S.bytes = 0;
  for (int w = 0; w < S.cnt; w++) {
    WinDesc d{};
    d.s = (int)((long long)w * (n - W) / S.cnt / S.D) * S.D;
    if (S.cnt > 1 && w > 0 && wd[w - 1].s + W > d.s) d.s = wd[w - 1].s + W;
    d.wl = W;
    if (odd && (w & 1)) d.s += 1;
    if (odd && w == S.cnt - 1) d.wl = W - 3;
    if (d.s + d.wl > n) d.s = n - d.wl;
    d.ilo = 0; d.ihi = n - 1;
    wd[w] = d;
    for (int j = 0; j < p; j++)
      for (int c = 0; c < d.wl; c++)
        for (int r = 0; r < d.wl; r++)
          u[(size_t)j * n * W + (size_t)d.s * W + (size_t)c * W + r] = (r == c ? 1.0 : 0.0) + 0.05 * rnd();
    const double len = (n - (d.s + d.wl)) + d.s + n;
    S.flops += p * 2.0 * d.wl * d.wl * len;
    S.bytes += p * 16.0 * d.wl * len;
  }
  CK(cudaMemcpy(S.U, u.data(), u.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(S.wins, wd.data(), wd.size() * sizeof(WinDesc), cudaMemcpyHostToDevice));
}

int main(int argc, char** argv) {
  Setup S;
  if (argc > 1) S.n = atoi(argv[1]);
  if (argc > 2) S.cnt = atoi(argv[2]);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n = S.n, p = S.p, W = S.W;
  const size_t nn = (size_t)n * n;
  S.h0.resize(nn);
  for (size_t i = 0; i < nn; i++) S.h0[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  for (int j = 0; j < p; j++) {
    double *a, *z;
    CK(cudaMalloc(&a, nn * 8)); CK(cudaMalloc(&z, nn * 8));
    S.H.push_back(a); S.Z.push_back(z);
  }
  CK(cudaMalloc(&S.U, (size_t)p * n * W * 8));
  CK(cudaMalloc(&S.wins, S.cnt * sizeof(WinDesc)));
  ApplyParams A{};
  A.n = n; A.p = p; A.W = W; A.wantT = 1; A.wantZ = 1; A.nwin = S.cnt;
  A.do_scan = 0; A.scan_ctl = nullptr; A.scan_ticket = nullptr; A.prof = nullptr;
  for (int j = 0; j < p; j++) { A.H[j] = S.H[j]; A.Z[j] = S.Z[j]; }
  A.U = S.U; A.wins = S.wins;
  CK(cudaFuncSetAttribute(ms_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
  if (argc > 3) {
    // profiling mode (ncu): one round
    make_windows(S, 0);
    reset(S);
    launch_v1(S, A, 2, 0, 3);
    CK(cudaDeviceSynchronize());
    return 0;
  }
  // ---- cross-check: the near / far split against the unsplit update, on windows with odd starts
  // and an order that is neither even nor a multiple of 4 (odd = 1) ----
  for (int odd = 0; odd < 2; odd++) {
    make_windows(S, odd);
    reset(S);
    launch_v1(S, A, 2, 0, 3);
    CK(cudaDeviceSynchronize());
    const std::vector<double> ref = snapshot(S);
    reset(S);
    launch_v1(S, A, 1, 1, 1);  // near0
    launch_v1(S, A, 2, 2, 1);  // far0
    launch_v1(S, A, 1, 1, 2);  // near1
    launch_v1(S, A, 2, 2, 2);  // far1
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    printf("odd %d: near + far vs unsplit, max |diff| %.3e\n", odd, maxdiff(ref, snapshot(S)));
  }
  // ---- timing (U close to the identity keeps the data bounded over the repetitions) ----
  make_windows(S, 0);
  reset(S);
  printf("n %d p %d W %d windows %d: %.2f GFLOP, %.1f MB per round\n", n, p, W, S.cnt, S.flops * 1e-9, S.bytes * 1e-6);
  auto report = [&](const char* name, float ms) {
    printf("%-34s %8.1f us  %6.2f TFLOP/s  %6.2f TB/s\n", name, ms * 1e3, S.flops / ms * 1e-9, S.bytes / ms * 1e-9);
  };
  for (int tpb : {1, 2, 4, 8}) {
    char name[64];
    snprintf(name, sizeof(name), "unsplit, %d tiles per CTA", tpb);
    report(name, time_rounds([&] { launch_v1(S, A, tpb, 0, 3); }, 20));
  }
  report("near0, far0, near1, far1", time_rounds([&] {
           launch_v1(S, A, 1, 1, 1);
           launch_v1(S, A, 2, 2, 1);
           launch_v1(S, A, 1, 1, 2);
           launch_v1(S, A, 2, 2, 2);
         }, 20));
  report("far only", time_rounds([&] { launch_v1(S, A, 2, 2, 3); }, 20));
  return 0;
}
