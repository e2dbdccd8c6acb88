// Small FP64 device helpers shared by the assembly kernels.  Operation order follows the Fortran
// expressions of the reference (no FMA contraction: the library is compiled with -fmad=false).
#pragma once
#include <cstdint>

namespace cfdl {

__device__ __forceinline__ double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// vec_weight, src/modules/mod_util.f90:729-743 — weight on the neighbour r2
__device__ __forceinline__ double vec_weight(const double r0[3], const double r1[3], const double r2[3]) {
  double ra0 = r1[0] - r0[0], ra1 = r1[1] - r0[1], ra2 = r1[2] - r0[2];
  double rb0 = r2[0] - r0[0], rb1 = r2[1] - r0[1], rb2 = r2[2] - r0[2];
  double la = sqrt(ra0 * ra0 + ra1 * ra1 + ra2 * ra2);
  double lb = sqrt(rb0 * rb0 + rb1 * rb1 + rb2 * rb2);
  return (la + lb > 0.0) ? la / (la + lb) : 0.5;
}

// a / b from y = RN(1/b): one multiply and two fused multiply-adds give the correctly rounded
// quotient (Markstein's correction: q0 = RN(a*y), r = a - b*q0 exactly, q = RN(q0 + r*y)) — the
// same bits as the division, for 3 FP64 instructions instead of the ~25 of the division routine.
// Checked against IEEE division on 4e8 random and adversarial operand pairs (all-ones significands,
// powers of two, near-ties): tools/check_fast_div.cpp.  Needs |a| >= 2^-969 or a == 0 so that r is
// exact (a == -0 returns +0).  FAST = false is the plain division.
template <bool FAST>
__device__ __forceinline__ double quot(double a, double b, double y) {
  if (FAST) {
    const double q0 = a * y;
    const double r = fma(-b, q0, a);
    return fma(r, y, q0);
  }
  return a / b;
}

// L2 residency hint for data that every pass of a solve reads again (the matrix coefficients): a
// fraction `frac` of the accesses made through the policy is marked evict_last, so that much of the
// array tends to stay in the 126 MB L2 across passes while everything else streams through.
// Caching hint only: the loaded values are the same.
__device__ __forceinline__ unsigned long long l2_keep_policy(float frac) {
#ifdef __CUDA_ARCH__
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(frac));
  return pol;
#else
  return 0ull;
#endif
}
__device__ __forceinline__ double ld_keep(const double* p, unsigned long long pol) {
#ifdef __CUDA_ARCH__
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
#else
  return *p;
#endif
}

__device__ __forceinline__ void load3(const double* __restrict__ a, int64_t i, double v[3]) {
  v[0] = a[3 * i]; v[1] = a[3 * i + 1]; v[2] = a[3 * i + 2];
}

}  // namespace cfdl
