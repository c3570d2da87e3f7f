// Counter-based synthetic-input generator: splitmix64 keyed by
// (seed, problem, factor, row, col, part) -> uniform [0,1) double.  Bit-identical to the
// oracle's generator (oracle/psdo_common.hpp) and to tests/psd_rng.py; mirrors rand(T,n,n)
// with a fixed seed in the reference's tests (test/testfuncs.jl:12, test/runtests.jl:20,93).
#pragma once
#include <cstdint>

namespace psd {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__host__ __device__ inline double gen_uniform(uint64_t seed, uint64_t b, uint64_t j, uint64_t r,
                                              uint64_t c, uint64_t part) {
  uint64_t k = splitmix64(seed);
  k = splitmix64(k ^ (b * 0x9E3779B97F4A7C15ull + 0x1234567ull));
  k = splitmix64(k ^ (j * 0xC2B2AE3D27D4EB4Full + 0x89ABCDEull));
  k = splitmix64(k ^ ((r << 32) | (c << 1) | part));
  return (double)(k >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace psd
