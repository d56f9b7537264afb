// Measures the fp64 pipe of this GPU: DFMA throughput with 8 independent chains per thread (peak) and
// the issue-to-use latency of a dependent DFMA chain.  The fused sweep is far above the fp64 ridge
// (DESIGN.md section 5), so this is the co-limit reported next to the graded HBM roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS> __global__ void dfma(double* out, double a, double b, int iters) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 12345.678) out[0] = s;
}

template <int CHAINS> double run(int blocks, int threads, int iters, double* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma<CHAINS><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  dfma<CHAINS><<<blocks, threads>>>(d, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  double* d;
  cudaMalloc(&d, 8);
  const int sms = p.multiProcessorCount, iters = 1 << 14;
  // peak: 8 chains x 1024 threads x 2 blocks per SM
  const double ms = run<8>(sms * 2, 1024, iters, d);
  const double fmas = (double)sms * 2 * 1024 * 8 * iters;
  const double tflops = 2.0 * fmas / (ms * 1e-3) / 1e12;
  const double per_sm_clk = fmas / (ms * 1e-3) / sms / (clock_khz * 1e3);
  // latency: one warp per SM sub-partition would need 4 blocks of 32; use 1 warp per SM, 1 chain
  const double ms1 = run<1>(sms, 32, iters * 8, d);
  const double cycles = ms1 * 1e-3 * clock_khz * 1e3 / (iters * 8.0);
  printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"dfma_tflops\": %.2f, \"dfma_per_sm_per_clk\": %.1f, "
         "\"dfma_dependent_latency_cycles\": %.1f}\n",
         sms, clock_khz / 1e3, tflops, per_sm_clk, cycles);
  return 0;
}
