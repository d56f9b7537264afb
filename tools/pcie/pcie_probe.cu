// Measures what the e2e path of bench.py is bounded by: contiguous vs pitched (cudaMemcpy3DAsync,
// valid cells of a ghosted 528^3 array) host<->device copies, one direction and both at once.
//   nvcc -O3 -o pcie_probe pcie_probe.cu && ./pcie_probe [n=512] [g=8] [ncomp=7]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

static void copy3d(double* dev, double* hst, int n, int g, int ncomp, bool to_dev, cudaStream_t s) {
  const size_t N = n + 2 * g;
  for (int c = 0; c < ncomp; ++c) {
    cudaMemcpy3DParms p{};
    cudaPitchedPtr d = make_cudaPitchedPtr(dev + N * N * N * c, N * 8, N, N);
    cudaPitchedPtr h = make_cudaPitchedPtr(hst + (size_t)n * n * n * c, (size_t)n * 8, n, n);
    cudaPos dp = make_cudaPos((size_t)g * 8, g, g), hp = make_cudaPos(0, 0, 0);
    p.srcPtr = to_dev ? h : d; p.srcPos = to_dev ? hp : dp;
    p.dstPtr = to_dev ? d : h; p.dstPos = to_dev ? dp : hp;
    p.extent = make_cudaExtent((size_t)n * 8, n, n);
    p.kind = to_dev ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpy3DAsync(&p, s));
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 512, g = argc > 2 ? atoi(argv[2]) : 8, nc = argc > 3 ? atoi(argv[3]) : 7;
  const size_t N = n + 2 * g, vbytes = (size_t)n * n * n * nc * 8, gbytes = N * N * N * nc * 8;
  double *h_in, *h_out, *d_a, *d_b;
  CK(cudaMallocHost(&h_in, vbytes)); CK(cudaMallocHost(&h_out, vbytes));
  CK(cudaMalloc(&d_a, gbytes)); CK(cudaMalloc(&d_b, gbytes));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto fn, double bytes) {
    fn(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s1));
      fn();
      CK(cudaStreamSynchronize(s2));
      CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
    }
    printf("%-34s %8.2f ms  %6.1f GB/s\n", name, best, bytes / best / 1e6);
  };
  timeit("H2D contiguous", [&] { CK(cudaMemcpyAsync(d_a, h_in, vbytes, cudaMemcpyHostToDevice, s1)); }, vbytes);
  timeit("D2H contiguous", [&] { CK(cudaMemcpyAsync(h_out, d_b, vbytes, cudaMemcpyDeviceToHost, s1)); }, vbytes);
  timeit("H2D pitched (valid cells)", [&] { copy3d(d_a, h_in, n, g, nc, true, s1); }, vbytes);
  timeit("D2H pitched (valid cells)", [&] { copy3d(d_b, h_out, n, g, nc, false, s1); }, vbytes);
  timeit("both contiguous (2 streams)", [&] {
    CK(cudaMemcpyAsync(d_a, h_in, vbytes, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(h_out, d_b, vbytes, cudaMemcpyDeviceToHost, s2)); }, 2.0 * vbytes);
  timeit("both pitched (2 streams)", [&] {
    copy3d(d_a, h_in, n, g, nc, true, s1); copy3d(d_b, h_out, n, g, nc, false, s2); }, 2.0 * vbytes);
  return 0;
}
