// fp64 operand-form costs on B200: cycles per warp instruction per scheduler (16 warps/SM, 4 chains)
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND> __global__ void k(double* out, const double* in, int iters) {
  double x0 = in[threadIdx.x], x1 = in[threadIdx.x + 32], x2 = in[threadIdx.x + 64], x3 = in[threadIdx.x + 96];
  double y0 = in[threadIdx.x + 128], y1 = in[threadIdx.x + 160], y2 = in[threadIdx.x + 192], y3 = in[threadIdx.x + 224];
  double z0 = in[threadIdx.x + 256], z1 = in[threadIdx.x + 288], z2 = in[threadIdx.x + 320], z3 = in[threadIdx.x + 352];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (KIND == 0) { x0 = fma(x0, y0, y0); x1 = fma(x1, y1, y1); x2 = fma(x2, y2, y2); x3 = fma(x3, y3, y3); }      // a*b+b
      if (KIND == 1) { x0 = fma(x0, y0, x0); x1 = fma(x1, y1, x1); x2 = fma(x2, y2, x2); x3 = fma(x3, y3, x3); }      // a*b+a
      if (KIND == 2) { x0 = fma(x0, y0, 1.5); x1 = fma(x1, y1, 1.5); x2 = fma(x2, y2, 1.5); x3 = fma(x3, y3, 1.5); }  // a*b+imm
      if (KIND == 3) { x0 = fma(x0, x0, x0); x1 = fma(x1, x1, x1); x2 = fma(x2, x2, x2); x3 = fma(x3, x3, x3); }      // a*a+a
      if (KIND == 4) { x0 = x0 * y0; x1 = x1 * y1; x2 = x2 * y2; x3 = x3 * y3; }                                      // DMUL 2 regs
      if (KIND == 5) { x0 = x0 + y0; x1 = x1 + y1; x2 = x2 + y2; x3 = x3 + y3; }                                      // DADD 2 regs
      if (KIND == 6) { x0 = fma(y0, z0, x0); x1 = fma(y1, z1, x1); x2 = fma(y2, z2, x2); x3 = fma(y3, z3, x3); }      // accumulate: y*z+x
      if (KIND == 7) { x0 = fma(x0, 1.25, y0); x1 = fma(x1, 1.25, y1); x2 = fma(x2, 1.25, y2); x3 = fma(x3, 1.25, y3); } // a*imm+b
    }
  }
  double s = x0 + x1 + x2 + x3 + y0 + z0 + z1 + z2 + z3;
  if (s == 12345.678) out[0] = s;
}
template <int KIND> double run(int blocks, int threads, int iters, double* d, double* in, int clock_khz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<KIND><<<blocks, threads>>>(d, in, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<KIND><<<blocks, threads>>>(d, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3 * clock_khz * 1e3 / ((double)(threads / 32) / 4.0 * iters * 16.0);
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  double *d, *in;
  cudaMalloc(&d, 8);
  cudaMalloc(&in, 8 * 2048);
  cudaMemset(in, 0, 8 * 2048);
  const int sms = p.multiProcessorCount, iters = 1 << 13, threads = 512;
  printf("a*b+b %.2f | a*b+a %.2f | a*b+imm %.2f | a*a+a %.2f | DMUL %.2f | DADD %.2f | y*z+x %.2f | a*imm+b %.2f\n",
         run<0>(sms, threads, iters, d, in, clock_khz), run<1>(sms, threads, iters, d, in, clock_khz),
         run<2>(sms, threads, iters, d, in, clock_khz), run<3>(sms, threads, iters, d, in, clock_khz),
         run<4>(sms, threads, iters, d, in, clock_khz), run<5>(sms, threads, iters, d, in, clock_khz),
         run<6>(sms, threads, iters, d, in, clock_khz), run<7>(sms, threads, iters, d, in, clock_khz));
  return 0;
}
