// Compares FastMath::div / FastMath::sqrt (finitevolumesolver_b200/csrc/fastmath.cuh)
// bit for bit against the IEEE operators on random and edge-case operands.
// Usage: fastmath_check [n_blocks]   prints "mismatch_div M mismatch_sqrt M flagged_div F flagged_sqrt F tested N"
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

__global__ void check(unsigned long long seed, int iters, unsigned long long* out) {
  unsigned long long s = seed + 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x);
  unsigned long long bad_div = 0, bad_sqrt = 0, fl_div = 0, fl_sqrt = 0, n = 0;
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
    n++;
  }
  atomicAdd(out + 0, bad_div);
  atomicAdd(out + 1, bad_sqrt);
  atomicAdd(out + 2, fl_div);
  atomicAdd(out + 3, fl_sqrt);
  atomicAdd(out + 4, n);
}

int main(int argc, char** argv) {
  const int blocks = argc > 1 ? atoi(argv[1]) : 148 * 8;
  unsigned long long* d;
  cudaMalloc(&d, 5 * sizeof(unsigned long long));
  cudaMemset(d, 0, 5 * sizeof(unsigned long long));
  check<<<blocks, 256>>>(42ull, 20000, d);
  unsigned long long h[5];
  if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("cuda error\n");
    return 2;
  }
  printf("mismatch_div %llu mismatch_sqrt %llu flagged_div %llu flagged_sqrt %llu tested %llu\n",
         h[0], h[1], h[2], h[3], h[4]);
  return (h[0] || h[1]) ? 1 : 0;
}
