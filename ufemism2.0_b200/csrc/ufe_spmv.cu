// CSR SpMV / multi-layer SpMM kernels (fp64 values, int32 1-based indices kept exactly as
// in type_sparse_matrix_CSR_dp; kernels subtract 1).
// Replaces multiply_CSR_matrix_with_vector_1D/_2D
// (src/UPSY/basic/CSR_matrix_algebra/CSR_matrix_vector_multiplication.f90:198-364): the
// reference streams the matrix once per layer (:355-359); here the nz layers of a 3-D
// field are accumulated in registers during a single pass over the matrix.
#include "ufe_internal.cuh"

// T threads cooperate on one row (T = 2..32, chosen from the mean row length); NL layers.
template <int T, int NL>
__global__ void __launch_bounds__(256)
k_spmv_rows(int m_loc, const int *__restrict__ ptr, const int *__restrict__ ind,
            const double *__restrict__ val, const double *__restrict__ x, long long ldx,
            double *__restrict__ y, long long ldy) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gt / T, lane = gt % T;
  const bool valid = row < m_loc;
  double acc[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) acc[l] = 0.0;
  if (valid) {
    const int k0 = ptr[row] - 1, k1 = ptr[row + 1] - 1;
    for (int k = k0 + lane; k < k1; k += T) {
      const double a = __ldg(val + k);
      const int c = __ldg(ind + k) - 1;
#pragma unroll
      for (int l = 0; l < NL; l++) acc[l] += a * __ldg(x + c + l * ldx);
    }
  }
#pragma unroll
  for (int off = T / 2; off > 0; off >>= 1) {
#pragma unroll
    for (int l = 0; l < NL; l++) acc[l] += __shfl_down_sync(0xffffffffu, acc[l], off, T);
  }
  if (valid && lane == 0) {
#pragma unroll
    for (int l = 0; l < NL; l++) y[row + l * ldy] = acc[l];
  }
}

template <int NL>
static int launch_T(cudaStream_t st, int T, int m_loc, const int *ptr, const int *ind, const double *val,
                    const double *x, long long ldx, double *y, long long ldy) {
  const int threads = 256;
  const long long total = (long long)m_loc * T;
  const int blocks = ufe_div_up(total, threads);
  if (blocks == 0) return UFE_OK;
  switch (T) {
    case 1: k_spmv_rows<1, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
    case 2: k_spmv_rows<2, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
    case 4: k_spmv_rows<4, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
    case 8: k_spmv_rows<8, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
    case 16: k_spmv_rows<16, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
    default: k_spmv_rows<32, NL><<<blocks, threads, 0, st>>>(m_loc, ptr, ind, val, x, ldx, y, ldy); break;
  }
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

// y(:, l) = A x(:, l) for l < nlayers.  x is indexed by global column (0-based = ind-1).
int ufe_spmv_launch(cudaStream_t st, int m_loc, int nnz, const int *ptr, const int *ind,
                    const double *val, const double *x, long long ldx, double *y, long long ldy,
                    int nlayers) {
  if (m_loc <= 0) return UFE_OK;
  const double mean = (double)nnz / (double)m_loc;
  int T = 1;
  while (T < 32 && T * 3 < mean) T *= 2;      // ~3..6 entries per thread
  int l = 0;
  while (nlayers - l >= 12) {
    UFE_TRY((launch_T<12>(st, T, m_loc, ptr, ind, val, x + l * ldx, ldx, y + l * ldy, ldy)));
    l += 12;
  }
  while (nlayers - l >= 4) {
    UFE_TRY((launch_T<4>(st, T, m_loc, ptr, ind, val, x + l * ldx, ldx, y + l * ldy, ldy)));
    l += 4;
  }
  for (; l < nlayers; l++)
    UFE_TRY((launch_T<1>(st, T, m_loc, ptr, ind, val, x + l * ldx, ldx, y + l * ldy, ldy)));
  return UFE_OK;
}
