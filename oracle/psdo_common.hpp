// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, scalar, sequential per problem) of the elementary
// transformations of RalphAS/PeriodicSchurDecompositions.jl.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call anything in oracle/.  The product (the CUDA library in
// periodicschurdecompositions.jl_b200/csrc) never includes or links this.
//
// Parity pinning: the reference is pure Julia and cannot run in this image
// (no julia binary), so this restatement is pinned by the reference's own test
// predicates (test/testfuncs.jl:56-145, 155-382), its one known-answer family
// (expsplit, test/testfuncs.jl:412-421) and eigvals of the explicit product.
//
// Citations are file:line relative to /root/reference.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

namespace psdo {

// Column-major matrix view with 1-based (i,j) access, mirroring Julia indexing so
// the restatement can be compared against the reference line by line.
template <class T>
struct MatT {
  T* d;
  int ld;
  inline T& operator()(int i, int j) const { return d[(i - 1) + (size_t)(j - 1) * ld]; }
};
using Mat = MatT<double>;
using CMat = MatT<std::complex<double>>;

// ---------------------------------------------------------------------------
// Counter-based input generator shared (bit-identically) by the oracle, the CUDA
// library tests and the Python harness: splitmix64 of a key built from
// (seed, problem, factor, row, col, part) -> uniform [0,1) double.
// Mirrors rand(T,n,n) with Random.seed!(1234) in spirit (test/testfuncs.jl:12).
// ---------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline double gen_uniform(uint64_t seed, uint64_t b, uint64_t j, uint64_t r, uint64_t c,
                                 uint64_t part) {
  uint64_t k = splitmix64(seed);
  k = splitmix64(k ^ (b * 0x9E3779B97F4A7C15ull + 0x1234567ull));
  k = splitmix64(k ^ (j * 0xC2B2AE3D27D4EB4Full + 0x89ABCDEull));
  k = splitmix64(k ^ ((r << 32) | (c << 1) | part));
  return (double)(k >> 11) * (1.0 / 9007199254740992.0);
}

// householder.jl:5-24  (_norm2, real): scaled sum of squares.
static inline double norm2(const double* x, int n, int inc = 1) {
  if (n < 1) return 0.0;
  if (n == 1) return std::fabs(x[0]);
  double scale = 0.0, ssq = 0.0;
  for (int i = 0; i < n; i++) {
    double xi = x[(size_t)i * inc];
    if (xi != 0.0) {
      double a = std::fabs(xi);
      if (scale < a) {
        double q = scale / a;
        ssq = 1.0 + ssq * q * q;
        scale = a;
      } else {
        double q = a / scale;
        ssq += q * q;
      }
    }
  }
  return scale * std::sqrt(ssq);
}

// householder.jl:26-54  (_norm2, complex): real and imaginary parts separately.
static inline double norm2(const std::complex<double>* x, int n, int inc = 1) {
  if (n < 1) return 0.0;
  if (n == 1) return std::abs(x[0]);
  double scale = 0.0, ssq = 0.0;
  auto acc = [&](double v) {
    if (v != 0.0) {
      double a = std::fabs(v);
      if (scale < a) {
        double q = scale / a;
        ssq = 1.0 + ssq * q * q;
        scale = a;
      } else {
        double q = a / scale;
        ssq += q * q;
      }
    }
  };
  for (int i = 0; i < n; i++) {
    acc(x[(size_t)i * inc].real());
    acc(x[(size_t)i * inc].imag());
  }
  return scale * std::sqrt(ssq);
}

// householder.jl:161-169  (_hypot3 = dlapy3)
static inline double hypot3(double x, double y, double z) {
  double xa = std::fabs(x), ya = std::fabs(y), za = std::fabs(z);
  double w = std::max(xa, std::max(ya, za));
  if (w == 0.0) return 0.0;  // (reference would give NaN; never reached with w==0 on its call path)
  double rw = 1.0 / w;
  return w * std::sqrt((rw * xa) * (rw * xa) + (rw * ya) * (rw * ya) + (rw * za) * (rw * za));
}

// householder.jl:66-108  (_xreflector!, real = dlarfg).  x has n entries with stride inc.
// On return x[0] = beta, x[1:] = v (essential part), returns tau.
static inline double reflector(double* x, int n, int inc = 1) {
  if (n <= 1) return 0.0;
  const double sfmin = 2.0 * DBL_MIN / DBL_EPSILON;
  double alpha = x[0];
  double xnorm = norm2(x + inc, n - 1, inc);
  if (xnorm == 0.0) return 0.0;
  double beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
  int kount = 0;
  bool smallb = std::fabs(beta) < sfmin;
  if (smallb) {
    const double rsfmin = 1.0 / sfmin;
    while (smallb) {
      kount++;
      for (int j = 1; j < n; j++) x[(size_t)j * inc] *= rsfmin;
      beta *= rsfmin;
      alpha *= rsfmin;
      smallb = (std::fabs(beta) < sfmin) && (kount < 20);
    }
    xnorm = norm2(x + inc, n - 1, inc);
    beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
  }
  double tau = (beta - alpha) / beta;
  double t = 1.0 / (alpha - beta);
  for (int j = 1; j < n; j++) x[(size_t)j * inc] *= t;
  for (int j = 0; j < kount; j++) beta *= sfmin;
  x[0] = beta;
  return tau;
}

// householder.jl:110-156  (_xreflector!, complex = zlarfg; n==1 is non-trivial).
static inline std::complex<double> reflector(std::complex<double>* x, int n, int inc = 1) {
  using C = std::complex<double>;
  if (n < 1) return C(0.0);
  const double sfmin = DBL_MIN / DBL_EPSILON;
  C alpha = x[0];
  double ar = alpha.real(), ai = alpha.imag();
  double xnorm = norm2(x + inc, n - 1, inc);
  if (xnorm == 0.0 && ai == 0.0) return C(0.0);
  double beta = -std::copysign(hypot3(ar, ai, xnorm), ar);
  int kount = 0;
  bool smallb = std::fabs(beta) < sfmin;
  if (smallb) {
    const double rsfmin = 1.0 / sfmin;
    while (smallb) {
      kount++;
      for (int j = 1; j < n; j++) x[(size_t)j * inc] *= rsfmin;
      beta *= rsfmin;
      ar *= rsfmin;
      ai *= rsfmin;
      smallb = (std::fabs(beta) < sfmin) && (kount < 20);
    }
    xnorm = norm2(x + inc, n - 1, inc);
    alpha = C(ar, ai);
    beta = -std::copysign(hypot3(ar, ai, xnorm), ar);
  }
  C tau((beta - ar) / beta, -ai / beta);
  C t = C(1.0) / (alpha - beta);
  for (int j = 1; j < n; j++) x[(size_t)j * inc] *= t;
  for (int j = 0; j < kount; j++) beta *= sfmin;
  x[0] = C(beta, 0.0);
  return tau;
}

// householder.jl:222-237  lmul!(H', A) on rows r0..r0+m-1, columns c0..c1 of A;
// v = essential part (m-1 entries, stride vinc).  For real tau' == tau (also :190-205).
template <class T>
static inline void hh_lmul_adj(const MatT<T>& A, int r0, int m, int c0, int c1, const T* v,
                               int vinc, T tau) {
  T tc = tau;
  if constexpr (!std::is_same<T, double>::value) tc = std::conj(tau);
  for (int j = c0; j <= c1; j++) {
    T va = A(r0, j);
    for (int i = 1; i < m; i++) {
      T vi = v[(size_t)(i - 1) * vinc];
      if constexpr (std::is_same<T, double>::value)
        va += vi * A(r0 + i, j);
      else
        va += std::conj(vi) * A(r0 + i, j);
    }
    va = tc * va;
    A(r0, j) -= va;
    for (int i = 1; i < m; i++) A(r0 + i, j) -= va * v[(size_t)(i - 1) * vinc];
  }
}

// householder.jl:207-220  rmul!(A, H) on rows r0..r1, columns c0..c0+m-1.
template <class T>
static inline void hh_rmul(const MatT<T>& A, int r0, int r1, int c0, int m, const T* v, int vinc,
                           T tau) {
  for (int i = r0; i <= r1; i++) {
    T x = A(i, c0);
    for (int k = 1; k < m; k++) x += A(i, c0 + k) * v[(size_t)(k - 1) * vinc];
    A(i, c0) -= tau * x;
    for (int k = 1; k < m; k++) {
      T vk = v[(size_t)(k - 1) * vinc];
      if constexpr (std::is_same<T, double>::value)
        A(i, c0 + k) -= tau * x * vk;
      else
        A(i, c0 + k) -= tau * x * std::conj(vk);
    }
  }
}

// householder.jl:269-304  HH2: 2-vector reflector with the whole vector (v1,v2) stored.
static inline void hh2_rmul(const Mat& A, int r0, int r1, int c0, double v1, double v2,
                            double tau) {
  double t1 = v1 * tau, t2 = v2 * tau;
  for (int i = r0; i <= r1; i++) {
    double s = A(i, c0) * v1 + A(i, c0 + 1) * v2;
    A(i, c0) -= s * t1;
    A(i, c0 + 1) -= s * t2;
  }
}
static inline void hh2_lmul_adj(const Mat& A, int r0, int c0, int c1, double v1, double v2,
                                double tau) {
  double t1 = tau * v1, t2 = tau * v2;
  for (int j = c0; j <= c1; j++) {
    double s = v1 * A(r0, j) + v2 * A(r0 + 1, j);
    A(r0, j) -= s * t1;
    A(r0 + 1, j) -= s * t2;
  }
}

// Julia stdlib LinearAlgebra.givensAlgorithm(f::Float64, g::Float64): translation of
// LAPACK dlartg (not under /root/reference).
// [c s; -s c] * [f; g] = [r; 0].
static inline void givens_real(double f, double g, double& cs, double& sn, double& r) {
  const double safmn2 = std::ldexp(1.0, -485);  // Julia floatmin2(Float64)
  const double safmx2 = 1.0 / safmn2;
  if (g == 0.0) {
    cs = 1.0; sn = 0.0; r = f;
  } else if (f == 0.0) {
    cs = 0.0; sn = 1.0; r = g;
  } else {
    double f1 = f, g1 = g;
    double scale = std::max(std::fabs(f1), std::fabs(g1));
    if (scale >= safmx2) {
      int count = 0;
      do {
        count++;
        f1 *= safmn2; g1 *= safmn2;
        scale = std::max(std::fabs(f1), std::fabs(g1));
      } while (scale >= safmx2 && count < 20);
      r = std::sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
      for (int i = 0; i < count; i++) r *= safmx2;
    } else if (scale <= safmn2) {
      int count = 0;
      do {
        count++;
        f1 *= safmx2; g1 *= safmx2;
        scale = std::max(std::fabs(f1), std::fabs(g1));
      } while (scale <= safmn2);
      r = std::sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
      for (int i = 0; i < count; i++) r *= safmn2;
    } else {
      r = std::sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
    }
    if (std::fabs(f) > std::fabs(g) && cs < 0.0) {
      cs = -cs; sn = -sn; r = -r;
    }
  }
}

// lmul!(Givens(i1,i2,c,s), A) restricted to columns c0..c1 (real).
static inline void rot_rows(const Mat& A, int i1, int i2, int c0, int c1, double c, double s) {
  for (int j = c0; j <= c1; j++) {
    double a1 = A(i1, j), a2 = A(i2, j);
    A(i1, j) = c * a1 + s * a2;
    A(i2, j) = -s * a1 + c * a2;
  }
}
// rmul!(A, Givens(j1,j2,c,s)') restricted to rows r0..r1 (real).
static inline void rot_cols_adj(const Mat& A, int j1, int j2, int r0, int r1, double c, double s) {
  for (int i = r0; i <= r1; i++) {
    double a1 = A(i, j1), a2 = A(i, j2);
    A(i, j1) = a1 * c + a2 * s;
    A(i, j2) = -a1 * s + a2 * c;
  }
}

// rschur2x2.jl:9-96  (_gs2x2! = dlanv2).  In: a,b,c,d.  Out: standardised a,b,c,d,
// rotation (cs,sn) and eigenvalues (w1r,w1i),(w2r,w2i); positive imaginary part first.
static inline void gs2x2(double& a, double& b, double& c, double& d, double& cs, double& sn,
                         double& w1r, double& w1i, double& w2r, double& w2i) {
  auto sgn = [](double x) { return x < 0 ? -1.0 : 1.0; };
  const double half = 0.5, small = 4.0 * DBL_EPSILON;
  if (c == 0.0) {
    cs = 1.0; sn = 0.0;
  } else if (b == 0.0) {
    cs = 0.0; sn = 1.0;
    double ta = a;
    a = d; b = -c; c = 0.0; d = ta;
  } else if ((a - d) == 0.0 && (b * c < 0.0)) {
    cs = 1.0; sn = 0.0;
  } else {
    double asubd = a - d;
    double p = half * asubd;
    double bcmax = std::max(std::fabs(b), std::fabs(c));
    double bcmis = std::min(std::fabs(b), std::fabs(c)) * sgn(b) * sgn(c);
    double scale = std::max(std::fabs(p), bcmax);
    double z = (p / scale) * p + (bcmax / scale) * bcmis;
    if (z >= small) {
      z = p + std::sqrt(scale) * std::sqrt(z) * sgn(p);
      a = d + z;
      d -= (bcmax / z) * bcmis;
      double tau = std::hypot(c, z);
      cs = z / tau; sn = c / tau;
      b -= c; c = 0.0;
    } else {
      double sigma = b + c;
      double tau = std::hypot(sigma, asubd);
      cs = std::sqrt(half * (1.0 + std::fabs(sigma) / tau));
      sn = -(p / (tau * cs)) * sgn(sigma);
      double aa = a * cs + b * sn, bb = -a * sn + b * cs;
      double cc = c * cs + d * sn, dd = -c * sn + d * cs;
      a = aa * cs + cc * sn; b = bb * cs + dd * sn;
      c = -aa * sn + cc * cs; d = -bb * sn + dd * cs;
      double midad = half * (a + d);
      a = midad; d = a;
      if (c != 0.0) {
        if (b != 0.0) {
          if (b * c >= 0.0) {
            double sab = std::sqrt(std::fabs(b)), sac = std::sqrt(std::fabs(c));
            p = sab * sac * sgn(c);
            tau = 1.0 / std::sqrt(std::fabs(b + c));
            a = midad + p; d = midad - p;
            b -= c; c = 0.0;
            double cs1 = sab * tau, sn1 = sac * tau;
            double ncs = cs * cs1 - sn * sn1, nsn = cs * sn1 + sn * cs1;
            cs = ncs; sn = nsn;
          }
        } else {
          b = -c; c = 0.0;
          double t = cs;
          cs = -sn; sn = t;
        }
      }
    }
  }
  if (c == 0.0) {
    w1r = a; w1i = 0.0; w2r = d; w2i = 0.0;
  } else {
    double rti = std::sqrt(std::fabs(b)) * std::sqrt(std::fabs(c));
    w1r = a; w1i = rti; w2r = d; w2i = -rti;
  }
}

// opnorm(view(H, r0:r1, c0:c1), 1): max absolute column sum.
static inline double opnorm1(const Mat& A, int r0, int r1, int c0, int c1) {
  double m = 0.0;
  for (int j = c0; j <= c1; j++) {
    double s = 0.0;
    for (int i = r0; i <= r1; i++) s += std::fabs(A(i, j));
    m = std::max(m, s);
  }
  return m;
}

}  // namespace psdo
