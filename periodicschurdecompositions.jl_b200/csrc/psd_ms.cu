// CUDA backend + host entry of the large-N multishift periodic QR iteration.
// Algorithm: psd_ms_core.cuh (bulge chase in diagonal windows), psd_ms_driver.hpp (sweep loop),
// psd_ms_kernels.cuh (kernels).  Compiled as its own translation unit and linked into
// libpsd_b200.so; called from launch_real_large in psd_capi.cu.
#include "psd_ms.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "psd_ms_driver.hpp"
#include "psd_ms_kernels.cuh"

namespace psd {
namespace ms {

struct Workspace {
  double* dU = nullptr;  size_t capU = 0;        // [p][n * W]
  double* dPairs = nullptr; size_t capPairs = 0; // shift pairs
  WinDesc* dPlan = nullptr; size_t capPlan = 0;  // windows of the current sweep (or final block list)
  int* dCtl = nullptr;                           // control block (16 ints)
  double* dSc = nullptr;                         // [2 * MS_MAXP] scales
  unsigned long long* dMax = nullptr;            // [MS_MAXP]
  int* hCtl = nullptr;                           // pinned mirror of dCtl
  WinDesc* hPlan = nullptr; size_t hcapPlan = 0; // pinned staging of the plan
  cudaEvent_t evCopy = nullptr;
};

Workspace* ws_create() { return new Workspace(); }

void ws_destroy(Workspace* ws) {
  if (!ws) return;
  cudaFree(ws->dU); cudaFree(ws->dPairs); cudaFree(ws->dPlan); cudaFree(ws->dCtl); cudaFree(ws->dSc);
  cudaFree(ws->dMax);
  cudaFreeHost(ws->hCtl); cudaFreeHost(ws->hPlan);
  if (ws->evCopy) cudaEventDestroy(ws->evCopy);
  delete ws;
}

static_assert(kMaxPeriod == MS_MAXP, "psd_ms.h and psd_ms_core.cuh disagree");

bool supported(int n, int p) { return p >= 1 && p <= MS_MAXP && n >= 2 * geom_for(p).W; }

namespace {

#define MS_CHECK(call)                      \
  do {                                      \
    cudaError_t e__ = (call);               \
    if (e__ != cudaSuccess) return e__;     \
  } while (0)

template <class T>
cudaError_t grow(T*& ptr, size_t& cap, size_t bytes) {
  if (bytes <= cap) return cudaSuccess;
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  cudaError_t e = cudaMalloc((void**)&ptr, bytes);
  if (e == cudaSuccess) cap = bytes;
  return e;
}

cudaError_t ws_basic(Workspace* ws) {
  if (!ws->dCtl) MS_CHECK(cudaMalloc((void**)&ws->dCtl, 16 * sizeof(int)));
  if (!ws->dSc) MS_CHECK(cudaMalloc((void**)&ws->dSc, 2 * MS_MAXP * sizeof(double)));
  if (!ws->dMax) MS_CHECK(cudaMalloc((void**)&ws->dMax, MS_MAXP * sizeof(unsigned long long)));
  if (!ws->hCtl) MS_CHECK(cudaHostAlloc((void**)&ws->hCtl, 16 * sizeof(int), cudaHostAllocDefault));
  if (!ws->evCopy) MS_CHECK(cudaEventCreateWithFlags(&ws->evCopy, cudaEventDisableTiming));
  return cudaSuccess;
}

struct Timer {  // optional per-launch timing (profile mode only)
  bool on = false;
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  void begin(int k) {
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    ev.push_back(a);
    ev.push_back(b);
    kind.push_back(k);
  }
  void end() {
    if (!on) return;
    cudaEventRecord(ev.back(), st);
  }
  void collect(double* ms) {
    for (size_t i = 0; i < kind.size(); i++) {
      float f = 0.f;
      cudaEventSynchronize(ev[2 * i + 1]);
      cudaEventElapsedTime(&f, ev[2 * i], ev[2 * i + 1]);
      ms[kind[i]] += f;
      cudaEventDestroy(ev[2 * i]);
      cudaEventDestroy(ev[2 * i + 1]);
    }
    ev.clear();
    kind.clear();
  }
};

struct CudaBackend {
  cudaStream_t st;
  int sm_count;
  Workspace* ws;
  int n, p, wantT, wantZ, maxitfac;
  Geom g;
  double* H[MS_MAXP];
  double* Z[MS_MAXP];
  double* dEig;
  int* dInfo;
  size_t chase_smem, shift_smem, block_smem;
  cudaError_t err = cudaSuccess;
  long long launches = 0;
  Timer tm;

  bool ok() const { return err == cudaSuccess; }
  void note(cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  }

  void scan(int nmin, int& ilo, int& ihi, int& done, int& nzero) {
    if (!ok()) { done = 1; return; }
    tm.begin(3);
    ms_scan_kernel<<<1, 1024, 0, st>>>(H[0], n, nmin, ws->dCtl);
    tm.end();
    launches++;
    note(cudaGetLastError());
    note(cudaMemcpyAsync(ws->hCtl, ws->dCtl, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
    if (!ok()) { done = 1; return; }
    ilo = ws->hCtl[0]; ihi = ws->hCtl[1]; done = ws->hCtl[2]; nzero = ws->hCtl[3];
  }

  int shifts(int lo, int m, double perturb) {
    if (!ok()) return 0;
    ShiftParams P;
    P.n = n; P.p = p; P.lo = lo; P.m = m;
    for (int j = 0; j < p; j++) P.H[j] = H[j];
    P.pairs = ws->dPairs; P.ctl = ws->dCtl; P.perturb = perturb;
    tm.begin(2);
    ms_shifts_kernel<<<1, 256, shift_smem, st>>>(P);
    tm.end();
    launches++;
    note(cudaGetLastError());
    note(cudaMemcpyAsync(ws->hCtl, ws->dCtl, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
    if (!ok()) return 0;
    return ws->hCtl[4];
  }

  void upload_plan(const std::vector<WinDesc>& plan) {
    if (!ok() || plan.empty()) return;
    const size_t bytes = plan.size() * sizeof(WinDesc);
    // the previous sweep's copy has long completed (scan/shifts synchronised the stream)
    if (bytes > ws->hcapPlan) {
      if (ws->hPlan) cudaFreeHost(ws->hPlan);
      ws->hPlan = nullptr;
      ws->hcapPlan = 0;
      note(cudaHostAlloc((void**)&ws->hPlan, bytes * 2, cudaHostAllocDefault));
      if (!ok()) return;
      ws->hcapPlan = bytes * 2;
    }
    note(grow(ws->dPlan, ws->capPlan, bytes));
    if (!ok()) return;
    std::copy(plan.begin(), plan.end(), ws->hPlan);
    note(cudaMemcpyAsync(ws->dPlan, ws->hPlan, bytes, cudaMemcpyHostToDevice, st));
  }

  void apply(const WinDesc* wins, int cnt) {
    ApplyParams A;
    A.n = n; A.p = p; A.W = g.W; A.wantT = wantT; A.wantZ = wantZ; A.nwin = cnt;
    for (int j = 0; j < p; j++) { A.H[j] = H[j]; A.Z[j] = Z[j]; }
    A.U = ws->dU; A.wins = wins;
    const int tiles = (n + AP_T - 1) / AP_T;
    tm.begin(1);
    A.phase = 0;
    ms_apply_kernel<<<dim3(tiles, cnt * p * 2), 256, AP_SMEM, st>>>(A);
    A.phase = 1;
    ms_apply_kernel<<<dim3(tiles, cnt * p), 256, AP_SMEM, st>>>(A);
    tm.end();
    launches += 2;
    note(cudaGetLastError());
  }

  void round(int off, int cnt) {
    if (!ok()) return;
    ChaseParams C;
    C.n = n; C.p = p; C.g = g;
    for (int j = 0; j < p; j++) C.H[j] = H[j];
    C.U = ws->dU; C.shifts = ws->dPairs; C.wins = ws->dPlan + off;
    tm.begin(0);
    ms_chase_kernel<<<cnt, 64 * g.NB, chase_smem, st>>>(C);
    tm.end();
    launches++;
    note(cudaGetLastError());
    apply(ws->dPlan + off, cnt);
  }

  void finish(int& nblocks) {
    nblocks = 0;
    if (!ok()) return;
    note(grow(ws->dPlan, ws->capPlan, (size_t)(n / 2 + 1) * sizeof(WinDesc)));
    if (!ok()) return;
    BlockParams B;
    B.n = n; B.p = p; B.W = g.W; B.wantT = wantT; B.wantZ = wantZ; B.maxitfac = maxitfac;
    for (int j = 0; j < p; j++) B.H[j] = H[j];
    B.U = ws->dU; B.eig = dEig; B.info = dInfo; B.list = ws->dPlan; B.ctl = ws->dCtl;
    tm.begin(4);
    ms_blocklist_kernel<<<1, 1024, 0, st>>>(B);
    launches++;
    note(cudaGetLastError());
    note(cudaMemcpyAsync(ws->hCtl, ws->dCtl, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
    if (!ok()) return;
    nblocks = ws->hCtl[5];
    if (nblocks > 0) {
      ms_blocks_kernel<<<std::min(nblocks, sm_count), 256, block_smem, st>>>(B, nblocks);
      launches++;
      note(cudaGetLastError());
    }
    tm.end();
    if (nblocks > 0 && (wantT || wantZ)) {
      // grid.y is limited to 65535: apply in slices of blocks
      const int per = std::max(1, 60000 / (2 * p));
      for (int o = 0; o < nblocks; o += per) apply(ws->dPlan + o, std::min(per, nblocks - o));
    }
  }
};

}  // namespace

cudaError_t prescale(cudaStream_t st, Workspace* ws, int n, int p, double* const* A) {
  if (p > MS_MAXP) return cudaSuccess;
  MS_CHECK(ws_basic(ws));
  MS_CHECK(cudaMemsetAsync(ws->dMax, 0, MS_MAXP * sizeof(unsigned long long), st));
  const long long cnt = (long long)n * n;
  for (int j = 0; j < p; j++) ms_maxabs_kernel<<<296, 256, 0, st>>>(A[j], cnt, ws->dMax + j);
  ms_scales_kernel<<<1, 32, 0, st>>>(ws->dMax, p, ws->dSc, ws->dCtl + 8);
  for (int j = 0; j < p; j++) ms_scale_kernel<<<592, 256, 0, st>>>(A[j], cnt, ws->dSc + j);
  return cudaGetLastError();
}

cudaError_t postscale(cudaStream_t st, Workspace* ws, int n, int p, double* const* A, int wantT, double* dEig) {
  if (p > MS_MAXP) return cudaSuccess;
  const long long cnt = (long long)n * n;
  if (wantT)
    for (int j = 0; j < p; j++) ms_scale_kernel<<<592, 256, 0, st>>>(A[j], cnt, ws->dSc + p + j);
  if (dEig) ms_scale_eig_kernel<<<32, 256, 0, st>>>(dEig, n, ws->dCtl + 8);
  return cudaGetLastError();
}

cudaError_t iterate(cudaStream_t st, int sm_count, Workspace* ws, int n, int p, double* const* H,
                    double* const* Z, int wantT, int wantZ, int maxitfac, double* dEig, int* dInfo,
                    int profile, Result* res) {
  MS_CHECK(ws_basic(ws));
  CudaBackend be;
  be.st = st; be.sm_count = sm_count; be.ws = ws;
  be.n = n; be.p = p; be.wantT = wantT; be.wantZ = (wantZ && Z) ? 1 : 0;
  be.maxitfac = maxitfac > 0 ? maxitfac : 30;
  be.g = geom_for(p);
  for (int j = 0; j < p; j++) {
    be.H[j] = H[j];
    be.Z[j] = be.wantZ ? Z[j] : nullptr;
  }
  be.dEig = dEig; be.dInfo = dInfo;
  be.tm.on = profile != 0;
  be.tm.st = st;
  const Geom g = be.g;
  MS_CHECK(grow(ws->dU, ws->capU, (size_t)p * n * g.W * sizeof(double)));
  DriverConfig cfg;
  cfg.n = n; cfg.p = p; cfg.wantT = wantT; cfg.wantZ = be.wantZ;
  // shift window: limited by the shared memory of one CTA
  int nsw = 64;
  while (nsw > 16 && ((size_t)((rp_small_doubles(nsw, p) + 1) & ~1LL) + (size_t)p * (nsw + 1) * nsw) * 8 > 200 * 1024) nsw -= 8;
  cfg.nsw = nsw;
  if (const char* ev = getenv("PSD_MS_REP")) cfg.rep_max = std::max(1, atoi(ev));
  if (const char* ev = getenv("PSD_MS_NSW")) cfg.nsw = std::max(2, std::min(nsw, atoi(ev)));
  MS_CHECK(grow(ws->dPairs, ws->capPairs, (size_t)(cfg.nsw + 2) * 4 * sizeof(double)));
  be.chase_smem = (size_t)2 * p * g.W * g.LD * sizeof(double);
  be.shift_smem = ((size_t)((rp_small_doubles(cfg.nsw, p) + 1) & ~1LL) + (size_t)p * (cfg.nsw + 1) * cfg.nsw) * 8;
  be.block_smem = ((size_t)((rp_small_doubles(g.W, p) + 1) & ~1LL) + (size_t)2 * p * (g.W + 1) * g.W) * 8;
  MS_CHECK(cudaFuncSetAttribute(ms_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be.chase_smem));
  MS_CHECK(cudaFuncSetAttribute(ms_shifts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be.shift_smem));
  MS_CHECK(cudaFuncSetAttribute(ms_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)be.block_smem));
  MS_CHECK(cudaFuncSetAttribute(ms_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AP_SMEM));
  DriverStats ds;
  const int status = drive(be, cfg, ds);
  if (!be.ok()) return be.err;
  if (res) {
    res->status = status;
    res->sweeps = ds.sweeps; res->rounds = ds.rounds; res->windows = ds.windows;
    res->shift_pairs = ds.shift_pairs; res->exceptional = ds.exceptional;
    res->final_blocks = ds.final_blocks; res->launches = be.launches; res->apply_flops = ds.apply_flops;
    if (profile) {
      double ms[5] = {0, 0, 0, 0, 0};
      be.tm.collect(ms);
      res->ms_chase = ms[0]; res->ms_apply = ms[1]; res->ms_shifts = ms[2]; res->ms_scan = ms[3]; res->ms_final = ms[4];
    }
  }
  if (getenv("PSD_MS_VERBOSE"))
    fprintf(stderr, "[psd ms] n %d p %d W %d NB %d nsw %d: status %d, %d sweeps, %lld rounds, %lld windows, %lld pairs, %d exceptional, %d final blocks, %.3f TFLOP applied\n",
            n, p, g.W, g.NB, cfg.nsw, status, ds.sweeps, ds.rounds, ds.windows, ds.shift_pairs, ds.exceptional,
            ds.final_blocks, ds.apply_flops * 1e-12);
  return cudaSuccess;
}

}  // namespace ms
}  // namespace psd
