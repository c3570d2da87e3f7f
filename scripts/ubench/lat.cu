// Dependent-chain latency microbenchmark for the FP64 path on B200 (one warp, one CTA).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = fma(x, b, a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double y = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) y = y * b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  double z = y + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z)); z = r + 1.5; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  double w = z;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w)); w = r + 1.5; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 4 independent DFMA chains (ILP)
  double p0 = w, p1 = w + 1, p2 = w + 2, p3 = w + 3;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { p0 = fma(p0, b, a); p1 = fma(p1, b, a); p2 = fma(p2, b, a); p3 = fma(p3, b, a); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // shfl chain
  double q = p0 + p1 + p2 + p3;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) q = __shfl_sync(0xffffffffu, q, (threadIdx.x + 1) & 31) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // smem store->load round trip chain
  __shared__ double sh[64];
  double s = q;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { sh[threadIdx.x] = s; __syncwarp(); s = sh[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // sqrt + div (library)
  double d = s;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 64; i++) d = sqrt(d) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 64; i++) d = 1.0 / d + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[threadIdx.x] = x + y + z + w + d;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 16 * 8);
  for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
  printf("DFMA dep %.2f cyc | DMUL dep %.2f | rsqrt.approx+DADD %.2f | rcp.approx+DADD %.2f | 4xDFMA ILP per-iter %.2f | shfl64+DADD %.2f | STS+sync+LDS+DADD %.2f | sqrt+DADD %.2f | div+DADD %.2f\n",
    cyc[0]/256.0, cyc[1]/256.0, cyc[2]/64.0, cyc[3]/64.0, cyc[4]/256.0, cyc[5]/64.0, cyc[6]/64.0, cyc[7]/64.0, cyc[8]/64.0);
  return 0;
}
