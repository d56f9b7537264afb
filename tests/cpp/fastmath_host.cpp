// csrc/fastmath.cuh on the CPU: the same algorithm with hardware FMA and emulated
// 20-bit MUFU seeds (the refinement, the negated-remainder zero handling and the
// validity tests are seed independent).  Usage: fastmath_host [iters]
#include <cstdio>
#include <cstdlib>
#define FVS_DEV inline
#include "../cuda/fastmath_cases.h"

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000000;
  fmcheck::Counters c{};
  fmcheck::run_cases(42ull, iters, c);
  printf("mismatch_div %llu mismatch_sqrt %llu flagged_div %llu flagged_sqrt %llu tested %llu "
         "tol_div_ulp %llu tol_sqrt_ulp %llu mismatch_divz %llu flagged_divz %llu zero_divz %llu\n",
         c.bad_div, c.bad_sqrt, c.fl_div, c.fl_sqrt, c.n, c.tol_div, c.tol_sqrt, c.bad_divz,
         c.fl_divz, c.zero_divz);
  return (c.bad_div || c.bad_sqrt || c.bad_divz || c.tol_div > 2 || c.tol_sqrt > 2) ? 1 : 0;
}
