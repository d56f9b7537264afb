// Compares FastMath::div / FastMath::sqrt (finitevolumesolver_b200/csrc/fastmath.cuh)
// bit for bit against the IEEE operators on random and edge-case operands.
// ToleranceMath (the opt-in arithmetic) is checked for its error bound instead: the largest error
// of its quotients / square roots in ulps of the IEEE result over moderate operands.
// Usage: fastmath_check [n_blocks]   prints "mismatch_div M mismatch_sqrt M flagged_div F flagged_sqrt F tested N tol_div_ulp U tol_sqrt_ulp U"
#include <cstdio>
#include <cstdlib>
#include "../../finitevolumesolver_b200/csrc/fastmath.cuh"

using namespace fvs;

__device__ unsigned long long splitmix(unsigned long long& s) {
  s += 0x9E3779B97F4A7C15ull;
  unsigned long long z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// mode 0: arbitrary bit patterns; 1: moderate exponents (|x| in 2^-40..2^40); 2: same-magnitude pairs
__device__ double draw(unsigned long long& s, int mode) {
  unsigned long long bits = splitmix(s);
  if (mode == 0) return __longlong_as_double((long long)bits);
  unsigned long long mant = bits & 0x000fffffffffffffull;
  unsigned long long sign = bits & 0x8000000000000000ull;
  int e = 1023 + (int)((bits >> 52) % 81) - 40;
  return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
}

__device__ unsigned long long ulp_distance(double got, double ref) {
  const long long a = __double_as_longlong(got), b = __double_as_longlong(ref);
  if ((a < 0) != (b < 0)) return ~0ull;
  return (unsigned long long)(a > b ? a - b : b - a);
}

__global__ void check(unsigned long long seed, int iters, unsigned long long* out) {
  unsigned long long s = seed + 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x);
  unsigned long long bad_div = 0, bad_sqrt = 0, fl_div = 0, fl_sqrt = 0, n = 0;
  unsigned long long tol_div = 0, tol_sqrt = 0;
  const double specials[12] = {0.0, -0.0, 1.0, -1.0, 2.2250738585072014e-308, 4.9e-324,
                               1.7976931348623157e308, 1e-300, 1e300, 0.5, 3.0,
                               __longlong_as_double(0x7ff0000000000000ll)};
  for (int it = 0; it < iters; ++it) {
    const int mode = it % 3;
    double a = draw(s, mode), b = draw(s, mode);
    if ((it & 63) == 0) a = specials[(it >> 6) % 12];
    if ((it & 127) == 1) b = specials[(it >> 7) % 12];
    // shared reciprocal with a second numerator
    double a2 = draw(s, mode);
    Flag f;
    const Rcp rc = FastMath::rcp(b);
    FastMath::check(rc, f);
    const double q1 = FastMath::div(a, rc, f);
    Flag f2;
    FastMath::check(rc, f2);
    const double q2 = FastMath::div(a2, rc, f2);
    const double r1 = a / b, r2 = a2 / b;
    if (!f.bad) {
      const bool same = __double_as_longlong(q1) == __double_as_longlong(r1) || (q1 != q1 && r1 != r1);
      bad_div += !same;
    } else {
      fl_div++;
    }
    if (!f2.bad) {
      const bool same = __double_as_longlong(q2) == __double_as_longlong(r2) || (q2 != q2 && r2 != r2);
      bad_div += !same;
    } else {
      fl_div++;
    }
    Flag g;
    const double x = fabs(a);
    const double sq = FastMath::sqrt(x, g);
    const double sr = sqrt(x);
    if (!g.bad) {
      bad_sqrt += __double_as_longlong(sq) != __double_as_longlong(sr);
    } else {
      fl_sqrt++;
    }
    if (mode != 0 && (it & 63) != 0 && (it & 127) != 1) {
      // moderate operands: inside the domain, no over- / underflow
      Flag t;
      const Rcp rt = ToleranceMath::rcp(b);
      ToleranceMath::check(rt, t);
      const unsigned long long ud = ulp_distance(ToleranceMath::div(a, rt, t), r1);
      const unsigned long long us = ulp_distance(ToleranceMath::sqrt(x, t), sr);
      if (!t.bad) {
        tol_div = ud > tol_div ? ud : tol_div;
        tol_sqrt = us > tol_sqrt ? us : tol_sqrt;
      }
    }
    n++;
  }
  atomicMax(out + 5, tol_div);
  atomicMax(out + 6, tol_sqrt);
  atomicAdd(out + 0, bad_div);
  atomicAdd(out + 1, bad_sqrt);
  atomicAdd(out + 2, fl_div);
  atomicAdd(out + 3, fl_sqrt);
  atomicAdd(out + 4, n);
}

int main(int argc, char** argv) {
  const int blocks = argc > 1 ? atoi(argv[1]) : 148 * 8;
  unsigned long long* d;
  cudaMalloc(&d, 7 * sizeof(unsigned long long));
  cudaMemset(d, 0, 7 * sizeof(unsigned long long));
  check<<<blocks, 256>>>(42ull, 20000, d);
  unsigned long long h[7];
  if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("cuda error\n");
    return 2;
  }
  printf("mismatch_div %llu mismatch_sqrt %llu flagged_div %llu flagged_sqrt %llu tested %llu "
         "tol_div_ulp %llu tol_sqrt_ulp %llu\n",
         h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  return (h[0] || h[1] || h[5] > 2 || h[6] > 2) ? 1 : 0;
}
