// Micro-benchmark of the window-update GEMM (ms_apply_kernel) in isolation: one synthetic round of
// the N = 4096, p = 4 pipeline (13 windows of order 56, T and Z wanted) repeated, timed with CUDA
// events.  Variants of the staging path are compared against the product kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
//        -I periodicschurdecompositions.jl_b200/csrc scripts/ubench/apply_bench.cu -o scripts/ubench/apply_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#include <cuda_runtime.h>
#include "psd_ms_kernels.cuh"

using namespace psd::ms;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(double* sdst, const double* gsrc, int nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* sdst, const double* gsrc, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// LOAD: 0 = 8-byte cp.async (product), 1 = 16-byte cp.async, 2 = bulk copies per column + mbarrier
// MMA / STORE / GLOAD: switch the phase off (timing experiments only; results are then wrong)
template <int LOAD, bool MMA, bool STORE, bool GLOAD>
__global__ void __launch_bounds__(256, 2) apply_variant(ApplyParams P) {
  double* Us = ms_smem;
  double* Xb[2] = {ms_smem + 64 * AP_LD, ms_smem + 2 * 64 * AP_LD};
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = P.n, p = P.p;
  const int kinds = (P.phase == 0) ? 2 : 1;
  int it = blockIdx.y;
  const int w = it / (p * kinds);
  it -= w * p * kinds;
  const int j = it / kinds + 1;
  const int kind = (P.phase == 0) ? (it % kinds == 0 ? 0 : 2) : 1;
  const WinDesc d = P.wins[w];
  const int s = d.s, wl = d.wl;
  double* X;
  int lo, hi;
  if (kind == 0) { X = P.H[j - 1]; lo = s + wl; hi = P.wantT ? n : d.ihi + 1; }
  else if (kind == 1) { X = P.H[(j == 1) ? p - 1 : j - 2]; lo = P.wantT ? 0 : d.ilo; hi = s; }
  else { X = P.Z[j - 1]; lo = 0; hi = P.wantZ ? n : 0; }
  const int tfirst = blockIdx.x * P.tpb;
  if (!(lo + tfirst * AP_T < hi)) return;
  const double* Ug = P.U + (size_t)(j - 1) * n * P.W + (size_t)s * P.W;

  if (LOAD == 2) {
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    // zero everything once: the padding is never written by the bulk copies
    for (int e = tid; e < 3 * 64 * AP_LD; e += 256) ms_smem[e] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  auto fetch = [&](double* dst, int t0, int tl, int buf, bool withU) {
    if (!GLOAD) return;
    if (LOAD == 0) {
#pragma unroll 4
      for (int i = 0; i < 16; i++) {
        const int e = tid + 256 * i;
        const int r = e & 63, cc = e >> 6;
        bool in; const double* g;
        if (kind == 0) { in = (r < wl && cc < tl); g = X + (s + r) + (size_t)(t0 + cc) * n; }
        else { in = (r < tl && cc < wl); g = X + (t0 + r) + (size_t)(s + cc) * n; }
        ms_cp_async8(dst + cc * AP_LD + r, in ? g : X, in ? 8 : 0);
      }
    } else if (LOAD == 1) {
#pragma unroll 4
      for (int i = 0; i < 8; i++) {
        const int e = tid + 256 * i;
        const int r = (e & 31) * 2, cc = e >> 5;
        bool in; const double* g;
        if (kind == 0) { in = (r < wl && cc < tl); g = X + (s + r) + (size_t)(t0 + cc) * n; }
        else { in = (r < tl && cc < wl); g = X + (t0 + r) + (size_t)(s + cc) * n; }
        cp_async16(dst + cc * AP_LD + r, in ? g : X, in ? 16 : 0);
      }
    } else {
      const int ncol = (kind == 0) ? tl : wl, len = (kind == 0) ? wl : tl;
      if (tid == 0) mbar_expect_tx(&bar[buf], (unsigned)(ncol * len * 8 + (withU ? wl * wl * 8 : 0)));
      if (tid < ncol) {
        const double* g = (kind == 0) ? X + s + (size_t)(t0 + tid) * n : X + t0 + (size_t)(s + tid) * n;
        bulk_g2s(dst + tid * AP_LD, g, (unsigned)(len * 8), &bar[buf]);
      } else if (withU && tid >= 64 && tid < 64 + wl) {
        const int cc = tid - 64;
        bulk_g2s(Us + cc * AP_LD, Ug + (size_t)cc * wl, (unsigned)(wl * 8), &bar[buf]);
      }
    }
  };
  if (LOAD != 2) {
#pragma unroll 4
    for (int i = 0; i < 16; i++) {
      const int e = tid + 256 * i;
      const int r = e & 63, cc = e >> 6;
      const bool in = (r < wl && cc < wl);
      ms_cp_async8(Us + cc * AP_LD + r, in ? Ug + r + (size_t)cc * wl : Ug, in ? 8 : 0);
    }
  }
  int t0 = lo + tfirst * AP_T;
  int tl = min(AP_T, hi - t0);
  {
    const bool g = GLOAD;
    if (!g && LOAD == 2) { /* nothing to wait for */ }
    // the first fetch always happens (U must be there)
    if (LOAD == 2) {
      const int ncol = (kind == 0) ? tl : wl, len = (kind == 0) ? wl : tl;
      if (tid == 0) mbar_expect_tx(&bar[0], (unsigned)(ncol * len * 8 + wl * wl * 8));
      if (tid < ncol) {
        const double* gp = (kind == 0) ? X + s + (size_t)(t0 + tid) * n : X + t0 + (size_t)(s + tid) * n;
        bulk_g2s(Xb[0] + tid * AP_LD, gp, (unsigned)(len * 8), &bar[0]);
      } else if (tid >= 64 && tid < 64 + wl) {
        const int cc = tid - 64;
        bulk_g2s(Us + cc * AP_LD, Ug + (size_t)cc * wl, (unsigned)(wl * 8), &bar[0]);
      }
    } else {
      const bool keep = GLOAD;
      (void)keep;
      // first tile through the normal path even in the no-load experiment
#pragma unroll 4
      for (int i = 0; i < 16; i++) {
        const int e = tid + 256 * i;
        const int r = e & 63, cc = e >> 6;
        bool in; const double* gp;
        if (kind == 0) { in = (r < wl && cc < tl); gp = X + (s + r) + (size_t)(t0 + cc) * n; }
        else { in = (r < tl && cc < wl); gp = X + (t0 + r) + (size_t)(s + cc) * n; }
        ms_cp_async8(Xb[0] + cc * AP_LD + r, in ? gp : X, in ? 8 : 0);
      }
      ms_cp_commit();
    }
  }
  const int wm = warp & 1, wn = warp >> 1;
  const int gq = lane >> 2, tq = lane & 3;
  const int kmax = (wl + 3) & ~3;
  unsigned phase_bits = 0;
  for (int tt = 0; tt < P.tpb; tt++) {
    double* Xs = Xb[tt & 1];
    if (LOAD == 2) {
      if (GLOAD || tt == 0) { mbar_wait(&bar[tt & 1], (phase_bits >> (tt & 1)) & 1u); phase_bits ^= 1u << (tt & 1); }
    } else {
      ms_cp_wait_all();
    }
    __syncthreads();
    const int t0n = t0 + AP_T;
    const bool more = (tt + 1 < P.tpb) && (t0n < hi);
    const int tln = more ? min(AP_T, hi - t0n) : 0;
    if (more) {
      fetch(Xb[(tt + 1) & 1], t0n, tln, (tt + 1) & 1, false);
      if (LOAD != 2) ms_cp_commit();
    }
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int q = 0; q < 2; q++) acc[i][q][0] = acc[i][q][1] = 0.0;
    if (MMA) {
      if (kind == 0) {
        for (int k0 = 0; k0 < kmax; k0 += 4) {
          double a[4], b[2];
#pragma unroll
          for (int i = 0; i < 4; i++) a[i] = Us[(wm * 32 + i * 8 + gq) * AP_LD + k0 + tq];
#pragma unroll
          for (int q = 0; q < 2; q++) b[q] = Xs[(wn * 16 + q * 8 + gq) * AP_LD + k0 + tq];
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int q = 0; q < 2; q++) ms_dmma(acc[i][q][0], acc[i][q][1], a[i], b[q]);
        }
      } else {
        for (int k0 = 0; k0 < kmax; k0 += 4) {
          double a[4], b[2];
#pragma unroll
          for (int i = 0; i < 4; i++) a[i] = Xs[(k0 + tq) * AP_LD + wm * 32 + i * 8 + gq];
#pragma unroll
          for (int q = 0; q < 2; q++) b[q] = Us[(wn * 16 + q * 8 + gq) * AP_LD + k0 + tq];
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int q = 0; q < 2; q++) ms_dmma(acc[i][q][0], acc[i][q][1], a[i], b[q]);
        }
      }
    } else {
      // keep a dependence on the staged data
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
          for (int e = 0; e < 2; e++)
            acc[i][q][e] = (kind == 0) ? Xs[(wn * 16 + q * 8 + 2 * tq + e) * AP_LD + wm * 32 + i * 8 + gq]
                                       : Xs[(wn * 16 + q * 8 + 2 * tq + e) * AP_LD + wm * 32 + i * 8 + gq];
    }
    if (STORE) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int m = wm * 32 + i * 8 + gq;
            const int nn = wn * 16 + q * 8 + 2 * tq + e;
            if (kind == 0) { if (m < wl && nn < tl) X[(s + m) + (size_t)(t0 + nn) * n] = acc[i][q][e]; }
            else { if (m < tl && nn < wl) X[(t0 + m) + (size_t)(s + nn) * n] = acc[i][q][e]; }
          }
    } else {
      double sum = 0;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int q = 0; q < 2; q++) sum += acc[i][q][0] + acc[i][q][1];
      if (sum == 1.2345e-300) X[0] = sum;
    }
    if (!more) break;
    t0 = t0n;
    tl = tln;
  }
}

struct Setup {
  int n = 4096, p = 4, W = 56, D = 28, cnt = 13;
  std::vector<double*> H, Z;
  double* U = nullptr;
  WinDesc* wins = nullptr;
  double flops = 0, bytes = 0;
  int tiles_total = 0;
};

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
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main(int argc, char** argv) {
  Setup S;
  if (argc > 1) S.n = atoi(argv[1]);
  if (argc > 2) S.cnt = atoi(argv[2]);
  const int n = S.n, p = S.p, W = S.W;
  const size_t nn = (size_t)n * n;
  std::vector<double> h(nn);
  for (size_t i = 0; i < nn; i++) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  for (int j = 0; j < p; j++) {
    double *a, *z;
    CK(cudaMalloc(&a, nn * 8)); CK(cudaMalloc(&z, nn * 8));
    CK(cudaMemcpy(a, h.data(), nn * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(z, h.data(), nn * 8, cudaMemcpyHostToDevice));
    S.H.push_back(a); S.Z.push_back(z);
  }
  // U_j: identity blocks (results stay bounded over the repetitions)
  std::vector<double> u((size_t)p * n * W, 0.0);
  std::vector<WinDesc> wd(S.cnt);
  for (int w = 0; w < S.cnt; w++) {
    WinDesc d{};
    d.s = (int)((long long)w * (n - W) / S.cnt / S.D) * S.D;
    d.wl = W; d.ilo = 0; d.ihi = n - 1;
    wd[w] = d;
    for (int j = 0; j < p; j++)
      for (int c = 0; c < W; c++) u[(size_t)j * n * W + (size_t)d.s * W + (size_t)c * W + c] = 1.0;
  }
  CK(cudaMalloc(&S.U, u.size() * 8));
  CK(cudaMemcpy(S.U, u.data(), u.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&S.wins, wd.size() * sizeof(WinDesc)));
  CK(cudaMemcpy(S.wins, wd.data(), wd.size() * sizeof(WinDesc), cudaMemcpyHostToDevice));
  // work of one round
  for (auto& d : wd) {
    const double cols_left = n - (d.s + W), rows_right = d.s, rows_z = n;
    S.flops += p * 2.0 * W * W * (cols_left + rows_right + rows_z);
    S.bytes += p * 16.0 * W * (cols_left + rows_right + rows_z);
  }
  ApplyParams A{};
  A.n = n; A.p = p; A.W = W; A.wantT = 1; A.wantZ = 1; A.nwin = S.cnt;
  A.do_scan = 0; A.scan_ctl = nullptr; A.scan_ticket = nullptr; A.prof = nullptr;
  for (int j = 0; j < p; j++) { A.H[j] = S.H[j]; A.Z[j] = S.Z[j]; }
  A.U = S.U; A.wins = S.wins;
  const int tiles = (n + AP_T - 1) / AP_T;
  CK(cudaFuncSetAttribute(ms_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
#define SETATTR(...) CK(cudaFuncSetAttribute(apply_variant<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM))
  SETATTR(0, true, true, true); SETATTR(0, false, true, true); SETATTR(0, true, false, true); SETATTR(0, true, true, false);
  SETATTR(1, true, true, true); SETATTR(2, true, true, true); SETATTR(2, false, true, true); SETATTR(1, false, true, true);
  SETATTR(2, true, false, true);
  printf("n %d p %d W %d windows %d: %.2f GFLOP, %.1f MB per round\n", n, p, W, S.cnt, S.flops * 1e-9, S.bytes * 1e-6);
  for (int tpb : {1, 2, 4, 8}) {
    A.tpb = tpb;
    const int chunks = (tiles + tpb - 1) / tpb;
    auto run = [&](const char* name, auto kern) {
      auto launch = [&] {
        ApplyParams B = A;
        B.phase = 0;
        kern<<<dim3(chunks, S.cnt * p * 2), 256, AP_SMEM>>>(B);
        B.phase = 1;
        kern<<<dim3(chunks, S.cnt * p), 256, AP_SMEM>>>(B);
      };
      const float ms = time_rounds(launch, 20);
      CK(cudaGetLastError());
      printf("tpb %d %-28s %8.1f us  %6.2f TFLOP/s  %6.2f TB/s\n", tpb, name, ms * 1e3, S.flops / ms * 1e-9, S.bytes / ms * 1e-9);
    };
    run("product", ms_apply_kernel);
    run("cp8", apply_variant<0, true, true, true>);
    run("cp8 no-mma", apply_variant<0, false, true, true>);
    run("cp8 no-store", apply_variant<0, true, false, true>);
    run("cp8 no-load", apply_variant<0, true, true, false>);
    run("cp16", apply_variant<1, true, true, true>);
    run("cp16 no-mma", apply_variant<1, false, true, true>);
    run("bulk", apply_variant<2, true, true, true>);
    run("bulk no-mma", apply_variant<2, false, true, true>);
    run("bulk no-store", apply_variant<2, true, false, true>);
  }
  // sanity: with U = I the real variants must leave the data unchanged
  A.tpb = 2;
  auto check = [&](const char* name, auto kern) {
    for (int j = 0; j < p; j++) {
      CK(cudaMemcpy(S.H[j], h.data(), nn * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(S.Z[j], h.data(), nn * 8, cudaMemcpyHostToDevice));
    }
    ApplyParams B = A;
    const int chunks = (tiles + B.tpb - 1) / B.tpb;
    B.phase = 0;
    kern<<<dim3(chunks, S.cnt * p * 2), 256, AP_SMEM>>>(B);
    B.phase = 1;
    kern<<<dim3(chunks, S.cnt * p), 256, AP_SMEM>>>(B);
    CK(cudaDeviceSynchronize());
    std::vector<double> back(nn);
    size_t bad = 0;
    for (int j = 0; j < p; j++) {
      CK(cudaMemcpy(back.data(), S.H[j], nn * 8, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < nn; i++) bad += (back[i] != h[i]);
      CK(cudaMemcpy(back.data(), S.Z[j], nn * 8, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < nn; i++) bad += (back[i] != h[i]);
    }
    printf("%-10s elements changed by identity updates: %zu\n", name, bad);
  };
  check("product", ms_apply_kernel);
  check("cp16", apply_variant<1, true, true, true>);
  check("bulk", apply_variant<2, true, true, true>);
  return 0;
}
