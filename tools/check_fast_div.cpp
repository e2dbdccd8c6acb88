// Empirical check of quot<true> (cfd-lite_b200/csrc/device_math.cuh): the quotient formed from a
// correctly rounded reciprocal with one multiply and two fused multiply-adds (Markstein's correction)
// against IEEE division, on random operands and on adversarial significands (all ones, nearly all
// ones, just above a power of two, powers of two).  Exit code 1 on any mismatch of the
// one-correction form (the one the kernels use).
//   g++ -O2 -ffp-contract=off -o check_fast_div tools/check_fast_div.cpp && ./check_fast_div 400000000
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
static inline double div1(double a, double b, double y) { double q0 = a * y; double r0 = std::fma(-b, q0, a); return std::fma(r0, y, q0); }
static inline double div2(double a, double b, double y) { double q1 = div1(a, b, y); double r1 = std::fma(-b, q1, a); return std::fma(r1, y, q1); }
static inline double from_bits(uint64_t u) { double d; std::memcpy(&d, &u, 8); return d; }
int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 200000000LL;
  std::mt19937_64 rng(12345);
  long long bad1 = 0, bad2 = 0;
  for (long long i = 0; i < n; ++i) {
    uint64_t ma = rng() & ((1ull << 52) - 1), mb = rng() & ((1ull << 52) - 1);
    int mode = i & 7;
    if (mode == 1) mb = (1ull << 52) - 1;                       // significand all ones
    if (mode == 2) mb = ((1ull << 52) - 1) ^ (rng() & 0xff);    // nearly all ones
    if (mode == 3) mb = rng() & 0xff;                           // just above a power of two
    if (mode == 4) ma = (1ull << 52) - 1 - (rng() & 0xf);
    if (mode == 5) { mb = 0; }                                  // power of two
    int ea = 1023 + (int)(rng() % 61) - 30, eb = 1023 + (int)(rng() % 61) - 30;
    double a = from_bits(((uint64_t)ea << 52) | ma), b = from_bits(((uint64_t)eb << 52) | mb);
    if (rng() & 1) a = -a;
    const double y = 1.0 / b, q = a / b;
    if (div1(a, b, y) != q) ++bad1;
    if (div2(a, b, y) != q) { if (bad2 < 5) printf("div2 mismatch a=%a b=%a got %a want %a\n", a, b, div2(a, b, y), q); ++bad2; }
  }
  printf("n=%lld  one-correction mismatches=%lld  two-correction mismatches=%lld\n", n, bad1, bad2);
  return bad1 ? 1 : 0;
}
