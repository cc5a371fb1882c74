// Deterministic grid-wide reduction: per-block partial sums, summed in a fixed order by
// the last block to finish (threadfence + atomic ticket).  Bitwise reproducible for a
// fixed grid size.  Blocks of at most 256 threads.
#pragma once
#include <cuda_runtime.h>

// ---------------------------------------------------------------------------------
// deterministic block reduction + last-block finalisation
// ---------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ bool reduce_publish(double (&v)[NV], double *partials, unsigned *counter,
                                               double *out) {
  __shared__ double sm[NV][8];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[i][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double s = 0.0;
      for (int w = 0; w < nw; w++) s += sm[i][w];
      partials[(size_t)blockIdx.x * NV + i] = s;
    }
    __threadfence();
    unsigned t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double s = 0.0;
    for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += ((volatile double *)partials)[(size_t)b * NV + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) sm[i][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double s = 0.0;
      for (int w = 0; w < nw; w++) s += sm[i][w];
      out[i] = s;
    }
    *counter = 0u;
    __threadfence();
  }
  __syncthreads();
  return true;
}

