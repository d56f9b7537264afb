// Compares FastMath::div / divz / sqrt (finitevolumesolver_b200/csrc/fastmath.cuh)
// bit for bit against the IEEE operators on random and edge-case operands, on the
// device.  ToleranceMath (the opt-in arithmetic) is checked for its error bound
// instead: the largest error of its quotients / square roots in ulps of the IEEE
// result over moderate operands.
// Usage: fastmath_check [n_blocks]   prints "name value" pairs on one line
#include <cstdio>
#include <cstdlib>
#include "fastmath_cases.h"

using namespace fmcheck;

__global__ void check(unsigned long long seed, int iters, unsigned long long* out) {
  Counters c{};
  run_cases(seed + 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x), iters, c);
  atomicAdd(out + 0, c.bad_div);
  atomicAdd(out + 1, c.bad_sqrt);
  atomicAdd(out + 2, c.fl_div);
  atomicAdd(out + 3, c.fl_sqrt);
  atomicAdd(out + 4, c.n);
  atomicMax(out + 5, c.tol_div);
  atomicMax(out + 6, c.tol_sqrt);
  atomicAdd(out + 7, c.bad_divz);
  atomicAdd(out + 8, c.fl_divz);
  atomicAdd(out + 9, c.zero_divz);
}

int main(int argc, char** argv) {
  const int blocks = argc > 1 ? atoi(argv[1]) : 148 * 8;
  unsigned long long* d;
  cudaMalloc(&d, 10 * sizeof(unsigned long long));
  cudaMemset(d, 0, 10 * sizeof(unsigned long long));
  check<<<blocks, 256>>>(42ull, 20000, d);
  unsigned long long h[10];
  if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("cuda error\n");
    return 2;
  }
  printf("mismatch_div %llu mismatch_sqrt %llu flagged_div %llu flagged_sqrt %llu tested %llu "
         "tol_div_ulp %llu tol_sqrt_ulp %llu mismatch_divz %llu flagged_divz %llu zero_divz %llu\n",
         h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  return (h[0] || h[1] || h[7] || h[5] > 2 || h[6] > 2) ? 1 : 0;
}
