// csrc/fastmath.cuh on the CPU: the same algorithm with hardware FMA and emulated
// 20-bit MUFU seeds (the refinement, the negated-remainder zero handling and the
// validity tests are seed independent).  Usage: fastmath_host [iters]
#include <cstdio>
#include <cstdlib>
#define FVS_DEV inline
#include "../cuda/fastmath_cases.h"

// Host only: a deterministic sweep over the exponents around the thresholds of the folded
// validity tests -- numerators and quotients from 2^-1074 to 2^-860 (the 2^-969 / 2^-940 /
// 2^-890 lines), divisors around the edges of the window 2^-74 .. 2^74, huge operands near
// 2^1017 -- with random mantissas: wherever the flag is clean the result must equal the IEEE
// operator bit for bit, and the clean cases must exist on the valid side of every line.
static int threshold_sweep(unsigned long long& clean, unsigned long long& tested) {
  using namespace fmcheck;
  unsigned long long s = 7;
  int bad = 0;
  const int eb[] = {-80, -76, -75, -74, -73, -72, -40, -1, 0, 1, 40, 72, 73, 74, 75, 76, 80};
  for (int ea = -1074; ea <= 1023; ea += (ea < -850 || ea > 1000) ? 1 : 37) {
    for (int ib = 0; ib < (int)(sizeof(eb) / sizeof(eb[0])); ++ib) {
      for (int rep = 0; rep < 24; ++rep) {
        const unsigned long long ma = splitmix(s) & 0x000fffffffffffffull;
        const unsigned long long mb = splitmix(s) & 0x000fffffffffffffull;
        double a = ea >= -1022 ? from_bits((long long)(((unsigned long long)(ea + 1023) << 52) | ma))
                               : from_bits((long long)((ma >> (-1022 - ea)) | 1ull));
        if (rep & 1) a = -a;
        const double b = from_bits((long long)(((unsigned long long)(eb[ib] + 1023) << 52) | mb));
        const Rcp rc = FastMath::rcp(b);
        {
          Flag f;
          FastMath::check_window(rc, f);
          const double q = FastMath::divw(a, rc, f);
          ++tested;
          if (!f.any()) {
            ++clean;
            bad += !same_bits(q, a / b);
          }
        }
        {
          Flag f;
          FastMath::check_window(rc, f);
          const double q = FastMath::divz(a, rc, f);
          ++tested;
          if (!f.any()) {
            ++clean;
            bad += !same_bits(q, a / b);
          }
        }
        {
          Flag f;
          FastMath::check(rc, f);
          const double q = FastMath::div(a, rc, f);
          ++tested;
          if (!f.any()) {
            ++clean;
            bad += !same_bits(q, a / b);
          }
        }
        {
          Flag f;
          const double x = fabs(a);
          const double r = FastMath::sqrt(x, f);
          ++tested;
          if (!f.any()) {
            ++clean;
            bad += !same_bits(r, sqrt(x));
          }
        }
      }
    }
  }
  return bad;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000000;
  fmcheck::Counters c{};
  fmcheck::run_cases(42ull, iters, c);
  unsigned long long clean = 0, tested = 0;
  const int sweep_bad = threshold_sweep(clean, tested);
  printf("threshold_sweep_mismatch %d threshold_sweep_clean %llu threshold_sweep_tested %llu ", sweep_bad,
         clean, tested);
  if (sweep_bad || clean == 0) return 1;
  printf("mismatch_div %llu mismatch_sqrt %llu flagged_div %llu flagged_sqrt %llu tested %llu "
         "tol_div_ulp %llu tol_sqrt_ulp %llu mismatch_divz %llu flagged_divz %llu zero_divz %llu\n",
         c.bad_div, c.bad_sqrt, c.fl_div, c.fl_sqrt, c.n, c.tol_div, c.tol_sqrt, c.bad_divz,
         c.fl_divz, c.zero_divz);
  return (c.bad_div || c.bad_sqrt || c.bad_divz || c.tol_div > 2 || c.tol_sqrt > 2) ? 1 : 0;
}
