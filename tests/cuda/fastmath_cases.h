// Shared by tests/cuda/fastmath_check.cu (device, real MUFU seeds) and
// tests/cpp/fastmath_host.cpp (CPU, emulated seeds): compares the branch-free
// division / square root of csrc/fastmath.cuh bit for bit against the IEEE
// operators wherever their validity flag is clean.
#pragma once
#include "../../finitevolumesolver_b200/csrc/fastmath.cuh"

namespace fmcheck {
using namespace fvs;

#ifdef __CUDACC__
#define FM_FN __device__ inline
FM_FN long long bits_of(double x) { return __double_as_longlong(x); }
FM_FN double from_bits(long long b) { return __longlong_as_double(b); }
#else
#define FM_FN inline
FM_FN long long bits_of(double x) {
  long long b;
  std::memcpy(&b, &x, 8);
  return b;
}
FM_FN double from_bits(long long b) {
  double x;
  std::memcpy(&x, &b, 8);
  return x;
}
#endif

struct Counters {
  unsigned long long bad_div, bad_sqrt, fl_div, fl_sqrt, n, tol_div, tol_sqrt, bad_divz, fl_divz,
      zero_divz;
};

FM_FN unsigned long long splitmix(unsigned long long& s) {
  s += 0x9E3779B97F4A7C15ull;
  unsigned long long z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// mode 0: arbitrary bit patterns; 1: moderate exponents (|x| in 2^-40..2^40);
// 2: same as 1; 3: wide exponents (2^-1074..2^1023) with zero low words now and then
FM_FN double draw(unsigned long long& s, int mode) {
  unsigned long long bits = splitmix(s);
  if (mode == 0) return from_bits((long long)bits);
  unsigned long long mant = bits & 0x000fffffffffffffull;
  unsigned long long sign = bits & 0x8000000000000000ull;
  if (mode == 3) {
    const unsigned long long k = splitmix(s);
    int e = (int)(k % 2047);
    if ((k >> 20) % 7 == 0) mant &= 0xfffff00000000ull; // zero low word
    if ((k >> 24) % 11 == 0) mant = 0;
    return from_bits((long long)(sign | ((unsigned long long)e << 52) | mant));
  }
  int e = 1023 + (int)((bits >> 52) % 81) - 40;
  return from_bits((long long)(sign | ((unsigned long long)e << 52) | mant));
}

FM_FN unsigned long long ulp_distance(double got, double ref) {
  const long long a = bits_of(got), b = bits_of(ref);
  if ((a < 0) != (b < 0)) return ~0ull;
  return (unsigned long long)(a > b ? a - b : b - a);
}
FM_FN bool same_bits(double a, double b) {
  return bits_of(a) == bits_of(b) || (a != a && b != b);
}

FM_FN void run_cases(unsigned long long seed, int iters, Counters& c) {
  unsigned long long s = seed;
  const double specials[16] = {0.0,
                               -0.0,
                               1.0,
                               -1.0,
                               2.2250738585072014e-308,
                               4.9e-324,
                               1.7976931348623157e308,
                               1e-300,
                               1e300,
                               0.5,
                               3.0,
                               from_bits(0x7ff0000000000000ll),
                               from_bits(0x0000000100000000ll),  // denormal, high word 1
                               from_bits(0x00000000ffffffffll),  // denormal, high word 0
                               from_bits(0x8000000000000001ll),  // -min denormal
                               from_bits(0x7ff8000000000000ll)}; // NaN
  for (int it = 0; it < iters; ++it) {
    const int mode = it % 4;
    double a = draw(s, mode), b = draw(s, mode);
    if ((it & 63) == 0) a = specials[(it >> 6) % 16];
    if ((it & 127) == 1) b = specials[(it >> 7) % 16];
    // shared reciprocal with a second numerator
    double a2 = draw(s, mode);
    {
      Flag f;
      const Rcp rc = FastMath::rcp(b);
      FastMath::check(rc, f);
      const double q1 = FastMath::div(a, rc, f);
      Flag f2;
      FastMath::check(rc, f2);
      const double q2 = FastMath::div(a2, rc, f2);
      const double r1 = a / b, r2 = a2 / b;
      if (!f.any()) {
        c.bad_div += !same_bits(q1, r1);
      } else {
        c.fl_div++;
      }
      if (!f2.any()) {
        c.bad_div += !same_bits(q2, r2);
      } else {
        c.fl_div++;
      }
      // zero-tolerant division over a divisor of either sign
      Flag f3;
      FastMath::check(rc, f3);
      const double az = (it & 3) == 2 ? ((it & 4) ? 0.0 : -0.0) : a;
      const double q3 = FastMath::div_any(az, rc, f3);
      if (!f3.any()) {
        c.bad_div += !same_bits(q3, az / b);
      } else {
        c.fl_div++;
      }
    }
    // zero-tolerant division over the positive window; every 4th numerator is +-0,
    // every 16th a denormal
    {
      double bz = fabs(b);
      double az = a;
      if ((it & 3) == 2) az = (it & 4) ? 0.0 : -0.0;
      if ((it & 15) == 5) az = from_bits((long long)(splitmix(s) & 0x800fffffffffffffull));
      if ((it & 31) == 9) az = from_bits((long long)(splitmix(s) & 0x80000000ffffffffull));
      Flag f;
      const Rcp rc = FastMath::rcp(bz);
      FastMath::check_window(rc, f);
      const double q = FastMath::divz(az, rc, f);
      if (!f.any()) {
        c.bad_divz += !same_bits(q, az / bz);
        c.zero_divz += az == 0.0;
      } else {
        c.fl_divz++;
      }
    }
    // division over a windowed divisor with only the quotient tested (divw), the
    // two-argument form, and the `used` overloads (an unused operation never flags)
    {
      const double bw = fabs(b);
      Flag f;
      const Rcp rc = FastMath::rcp(bw);
      FastMath::check_window(rc, f);
      const double q = FastMath::divw(a, rc, f);
      if (!f.any()) {
        c.bad_div += !same_bits(q, a / bw);
      } else {
        c.fl_div++;
      }
      Flag g2;
      const double q2 = FastMath::div(a, b, g2);
      if (!g2.any()) c.bad_div += !same_bits(q2, a / b);
      Flag g3;
      const bool used = (it & 8) != 0;
      const double q3 = FastMath::div(a, b, g3, used);
      const double s3 = FastMath::sqrt(a2, g3, used);
      const Rcp rc3 = FastMath::rcp(a2);
      FastMath::check(rc3, g3, used);
      const double q4 = FastMath::div(a, rc3, g3, used);
      if (!used) {
        c.bad_div += g3.any();
      } else if (!g3.any()) {
        c.bad_div += !same_bits(q3, a / b) + !same_bits(q4, a / a2);
        c.bad_sqrt += !same_bits(s3, sqrt(a2));
      }
    }
    {
      Flag g;
      const double x = fabs(a);
      const double sq = FastMath::sqrt(x, g);
      const double sr = sqrt(x);
      if (!g.any()) {
        c.bad_sqrt += bits_of(sq) != bits_of(sr);
      } else {
        c.fl_sqrt++;
      }
      if ((mode == 1 || mode == 2) && (it & 63) != 0 && (it & 127) != 1) {
        // moderate operands: inside the domain, no over- / underflow
        Flag t;
        const Rcp rt = ToleranceMath::rcp(b);
        ToleranceMath::check(rt, t);
        const unsigned long long ud = ulp_distance(ToleranceMath::div(a, rt, t), a / b);
        const unsigned long long us = ulp_distance(ToleranceMath::sqrt(x, t), sr);
        if (!t.any()) {
          c.tol_div = ud > c.tol_div ? ud : c.tol_div;
          c.tol_sqrt = us > c.tol_sqrt ? us : c.tol_sqrt;
        }
      }
    }
    c.n++;
  }
}

} // namespace fmcheck
