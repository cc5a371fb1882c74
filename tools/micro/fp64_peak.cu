// fp64 SIMT peak of the device: independent DFMA chains, no memory traffic.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// usage: ./fp64_peak   -> TFLOP/s for several (warps per SM, chains per thread) combinations
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 123.456) out[0] = s;
}
template <int CH> static void run(int threads, int blocks_per_sm, int sms) {
  double *out; cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<sms * blocks_per_sm, threads>>>(out, 100, 0.999, 1e-3);
  cudaEventRecord(e0);
  k<CH><<<sms * blocks_per_sm, threads>>>(out, iters, 0.999, 1e-3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * CH * iters * (double)threads * blocks_per_sm * sms;
  printf("chains %2d  threads/block %4d  blocks/SM %d  (%2d warps/SM): %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", CH, threads, blocks_per_sm,
         threads * blocks_per_sm / 32, flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<8>(64, 1, sms); run<8>(128, 1, sms); run<8>(256, 1, sms); run<8>(256, 2, sms); run<8>(1024, 2, sms);
  run<16>(64, 1, sms); run<16>(256, 1, sms); run<16>(256, 4, sms);
  run<64>(64, 1, sms); run<64>(256, 1, sms); run<64>(256, 2, sms);
  return 0;
}
