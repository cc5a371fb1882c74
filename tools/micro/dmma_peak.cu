// fp64 tensor-core (mma.sync.m8n8k4.f64) peak of the device: independent accumulator tiles per warp, operands in registers.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NT>
__global__ void k(double *out, int iters) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1e-3 * threadIdx.x, b = 1.0 - 1e-6 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template <int NT> static void run(int threads, int blocks_per_sm, int sms) {
  double *out; cudaMalloc(&out, 8);
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NT><<<sms * blocks_per_sm, threads>>>(out, 50);
  cudaEventRecord(e0);
  k<NT><<<sms * blocks_per_sm, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 256.0 * NT * iters * (double)(threads / 32) * blocks_per_sm * sms;
  printf("tiles/warp %2d  threads/block %4d  blocks/SM %d (%2d warps/SM): %.2f TFLOP/s\n", NT, threads, blocks_per_sm, threads * blocks_per_sm / 32, flops / ms / 1e9);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<8>(128, 1, sms); run<8>(256, 1, sms); run<8>(256, 2, sms); run<32>(128, 1, sms); run<32>(256, 1, sms); run<32>(256, 2, sms); run<32>(512, 2, sms);
  return 0;
}
