// Does a non-fp64 instruction issue "for free" next to a DFMA on B200?  Times a loop of 4
// independent DFMA chains per thread with K extra instructions of another pipe per DFMA
// (K = 0, 1, 2; ALU: LOP3/IADD, FMA pipe: FFMA, predicate: FSETP) at 12 and 16 warps per SM.
// Result (profiles/r02_tuning.md): cycles per DFMA per scheduler ~ 2 + K, i.e. a DFMA holds the
// dispatch port for two cycles and every other instruction costs one more.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int KIND> __global__ void mix(double* out, double a, double b, int iters, int seed) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  int i0 = seed + threadIdx.x, i1 = seed * 3, i2 = seed * 5, i3 = seed * 7;
  float f0 = seed * 0.5f, f1 = seed * 0.25f, f2 = 1.5f, f3 = 2.5f;
  bool p = true;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      if (K >= 1) {
        if (KIND == 0) { i0 = (i0 ^ i1) + 0x9e37; i1 = (i1 & i2) | 0x55; i2 = (i2 ^ i3) + 3; i3 = (i3 | i0) ^ 0x77; }
        if (KIND == 1) { f0 = fmaf(f0, 1.0001f, f1); f1 = fmaf(f1, 0.9999f, f2); f2 = fmaf(f2, 1.0001f, f3); f3 = fmaf(f3, 0.9999f, f0); }
        if (KIND == 2) { p = p && (f0 < __int_as_float(i0)); p = p && (f1 < __int_as_float(i1)); p = p && (f2 < __int_as_float(i2)); p = p && (f3 < __int_as_float(i3)); i0 += 1; }
      }
      if (K >= 2) {
        if (KIND == 0) { i0 = (i0 ^ i2) + 0x1234; i1 = (i1 & i3) | 0x11; i2 = (i2 ^ i0) + 7; i3 = (i3 | i1) ^ 0x33; }
        if (KIND == 1) { f0 = fmaf(f0, 1.0002f, f2); f1 = fmaf(f1, 0.9998f, f3); f2 = fmaf(f2, 1.0002f, f0); f3 = fmaf(f3, 0.9998f, f1); }
        if (KIND == 2) { p = p && (f0 > __int_as_float(i1)); p = p && (f1 > __int_as_float(i2)); p = p && (f2 > __int_as_float(i3)); p = p && (f3 > __int_as_float(i0)); i1 += 1; }
      }
    }
  }
  double s = x0 + x1 + x2 + x3 + i0 + i1 + i2 + i3 + f0 + f1 + f2 + f3 + (p ? 1.0 : 0.0);
  if (s == 12345.678) out[0] = s;
}

template <int K, int KIND> double run(int blocks, int threads, int iters, double* d, int clock_khz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mix<K, KIND><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters, 12345);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mix<K, KIND><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters, 12345);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // cycles per DFMA warp-instruction per scheduler: one block per SM, threads/32 warps over 4 schedulers
  const double dfma_per_sched = (double)(threads / 32) / 4.0 * iters * 16.0;
  return ms * 1e-3 * clock_khz * 1e3 / dfma_per_sched;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  double* d;
  cudaMalloc(&d, 8);
  const int sms = p.multiProcessorCount, iters = 1 << 13;
  for (int threads : {384, 512, 1024}) {
    printf("warps/SM %2d: cycles per DFMA per scheduler  K=0 %.2f | ALU K=1 %.2f K=2 %.2f | FFMA K=1 %.2f K=2 %.2f | FSETP K=1 %.2f K=2 %.2f\n",
           threads / 32, run<0, 0>(sms, threads, iters, d, clock_khz),
           run<1, 0>(sms, threads, iters, d, clock_khz), run<2, 0>(sms, threads, iters, d, clock_khz),
           run<1, 1>(sms, threads, iters, d, clock_khz), run<2, 1>(sms, threads, iters, d, clock_khz),
           run<1, 2>(sms, threads, iters, d, clock_khz), run<2, 2>(sms, threads, iters, d, clock_khz));
  }
  return 0;
}
