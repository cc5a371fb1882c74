// Stiffness-matrix assembly into the reference's CSR layout.
// Replaces the assembly part of solve_SSA_DIVA_linearised
// (src/UFEMISM/ice_dynamics/conservation_of_momentum/SSA_DIVA/solve_linearised_SSA_DIVA.f90
// :23-153) and its row builders _row_free (:180-329), _sans_ (:331-479), _row_BC (:481-641).
//
// The reference rebuilds the matrix from scratch with append-only add_entry_CSR_dist on
// every Picard iteration although the sparsity pattern is constant; here the pattern
// (ptr, ind -- identical to what the append sequence produces) is built once per solve
// call and only the values are rewritten per iteration, one thread per triangle writing
// the row pair (2ti-1, 2ti).  The same pass also writes the left-preconditioned copy
// valS = B*A, bS = B*b (B = point Jacobi or 2x2 u-v block Jacobi) that the Krylov loop
// runs on.  Compiled with -fmad=false (values match the CPU evaluation order exactly).
#include "ufe_closures.cuh"

int ufe_counts_to_ptr(cudaStream_t st, int m_loc, int *counts, int *ptr, int *nnz_out);

enum { RK_FREE = 0, RK_PRESCR = 1, RK_INFINITE = 2, RK_ZERO = 3, RK_COPY = 4 };

__device__ __forceinline__ int bc_side(int tribi) {
  switch (tribi) {
    case 1: case 2: return 0;
    case 3: case 4: return 1;
    case 5: case 6: return 2;
    default: return 3;
  }
}

__device__ __forceinline__ int row_kind(const AssemblyParams &A, int tribi, int prescr, int uv) {
  if (prescr == 1) return RK_PRESCR;
  if (tribi > 0) {
    const int side = bc_side(tribi);
    const int choice = uv == 0 ? A.bc_u[side] : A.bc_v[side];
    if (choice == UFE_BC_INFINITE) return RK_INFINITE;
    if (choice == UFE_BC_ZERO) return RK_ZERO;
    return RK_COPY;
  }
  return RK_FREE;
}

__global__ void k_count_rows(int t0, int nt, int nTri, AssemblyParams A, DevFamilyView M2,
                             const int *__restrict__ TriBI, const int *__restrict__ TriC,
                             const int *__restrict__ bc_mask, int *__restrict__ counts, int *__restrict__ rowkind) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  const int prescr = bc_mask ? bc_mask[ti] : 0;
  for (int uv = 0; uv < 2; uv++) {
    const int kind = row_kind(A, TriBI[ti], prescr, uv);
    int c = 1;
    if (kind == RK_FREE) c = 2 * (M2.ptr[tl + 1] - M2.ptr[tl]);
    else if (kind == RK_INFINITE) {
      int nn = 0;
      for (int n = 0; n < 3; n++) if (TriC[(size_t)n * nTri + ti] != 0) nn++;
      c = nn + 1;
    }
    counts[2 * tl + uv] = c;
    rowkind[2 * tl + uv] = kind;
  }
}

__global__ void k_fill_ind(int t0, int nt, int nTri, DevFamilyView M2, const int *__restrict__ TriC,
                           const int *__restrict__ rowkind, const int *__restrict__ ptrA, int *__restrict__ indA) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  for (int uv = 0; uv < 2; uv++) {
    const int kind = rowkind[2 * tl + uv];
    int k = ptrA[2 * tl + uv] - 1;
    const int row = 2 * ti + uv + 1;           // 1-based global row, tiuv2n
    if (kind == RK_FREE) {
      for (int j = M2.ptr[tl] - 1; j < M2.ptr[tl + 1] - 1; j++) {
        const int tj = M2.ind[j];               // 1-based
        indA[k++] = 2 * (tj - 1) + 1;
        indA[k++] = 2 * (tj - 1) + 2;
      }
    } else if (kind == RK_INFINITE) {
      for (int n = 0; n < 3; n++) {
        const int tj = TriC[(size_t)n * nTri + ti];
        if (tj == 0) continue;
        indA[k++] = 2 * (tj - 1) + uv + 1;
      }
      indA[k++] = row;
    } else {
      indA[k++] = row;
    }
  }
}

__global__ void k_colrange(int nnz, const int *__restrict__ ind, int *mn, int *mx) {
  int lo = 0x7fffffff, hi = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += gridDim.x * blockDim.x) {
    const int c = ind[i];
    lo = min(lo, c); hi = max(hi, c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(mn, lo); atomicMax(mx, hi); }
}

int ufe_colrange(cudaStream_t st, int nnz, const int *ind, int *jmin, int *jmax) {
  int *d = nullptr;
  UFE_CUDA(cudaMalloc(&d, 2 * sizeof(int)));
  int init[2] = {0x7fffffff, 0};
  UFE_CUDA(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, st));
  if (nnz > 0) { k_colrange<<<296, 256, 0, st>>>(nnz, ind, d, d + 1); UFE_LAUNCH_CHECK(); }
  int out[2];
  UFE_CUDA(cudaMemcpyAsync(out, d, sizeof out, cudaMemcpyDeviceToHost, st));
  UFE_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  *jmin = out[0]; *jmax = out[1];
  return UFE_OK;
}

// ---- blocked sliced-ELL copy for the Krylov loop (layout: DevSystem in ufe_internal.cuh) ----
__device__ __forceinline__ int bell_row_width(int tl, int ti, int nTri, const int *__restrict__ rowkind,
                                              const DevFamilyView &M2, const int *__restrict__ TriC) {
  const int ku = rowkind[2 * tl], kv = rowkind[2 * tl + 1];
  if (ku == RK_FREE) return M2.ptr[tl + 1] - M2.ptr[tl];
  if (ku == RK_INFINITE || kv == RK_INFINITE) {
    int nn = 0;
    for (int n = 0; n < 3; n++) if (TriC[(size_t)n * nTri + ti] != 0) nn++;
    return nn + 1;
  }
  return 1;
}

__global__ void k_bell_slice_width(int t0, int nt, int nTri, DevFamilyView M2, const int *__restrict__ TriC,
                                   const int *__restrict__ rowkind, int *__restrict__ slice_w) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= (nt + 31) / 32) return;
  const int tl = s * 32 + lane;
  int w = tl < nt ? bell_row_width(tl, t0 + tl, nTri, rowkind, M2, TriC) : 0;
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if (lane == 0) slice_w[s] = w;
}

__global__ void k_bell_fill_cols(int t0, int nt, int nTri, DevFamilyView M2, const int *__restrict__ TriC,
                                 const int *__restrict__ rowkind, const int *__restrict__ bell_off,
                                 int *__restrict__ bell_col) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= (nt + 31) / 32) return;
  const int tl = s * 32 + lane;
  const int off = bell_off[s] - 1, w = bell_off[s + 1] - 1 - off;
  const int ti = tl < nt ? t0 + tl : t0;
  int e = 0;
  if (tl < nt) {
    const int ku = rowkind[2 * tl], kv = rowkind[2 * tl + 1];
    if (ku == RK_FREE) {
      for (int j = M2.ptr[tl] - 1; j < M2.ptr[tl + 1] - 1; j++, e++) bell_col[(size_t)(off + e) * 32 + lane] = M2.ind[j] - 1;
    } else {
      if (ku == RK_INFINITE || kv == RK_INFINITE)
        for (int n = 0; n < 3; n++) {
          const int tj = TriC[(size_t)n * nTri + ti];
          if (tj == 0) continue;
          bell_col[(size_t)(off + e) * 32 + lane] = tj - 1; e++;
        }
      bell_col[(size_t)(off + e) * 32 + lane] = ti; e++;
    }
  }
  for (; e < w; e++) bell_col[(size_t)(off + e) * 32 + lane] = ti;     // padding: own triangle, zero values
}

// Builds S.ptr / S.ind for rows 2*t0 .. 2*(t0+nt)-1 (0-based) and allocates val / valS /
// bb / bS (S.x is allocated by the caller).  *rowkind_io is (re)allocated.
int ufe_build_stiffness_pattern(cudaStream_t st, int t0, int nt, int nTri, const AssemblyParams &A,
                                DevFamilyView M2, const int *TriBI, const int *TriC, const int *bc_mask,
                                DevSystem &S, int **rowkind_io) {
  const int m_loc = 2 * nt;
  int *counts = nullptr;
  cudaFree(S.ptr); cudaFree(S.ind); cudaFree(S.val); cudaFree(S.valS); cudaFree(S.bb); cudaFree(S.bS);
  cudaFree(S.bell_off); cudaFree(S.bell_col); cudaFree(S.bell_val);
  cudaFree(*rowkind_io);
  S.ptr = S.ind = nullptr; S.val = S.valS = S.bb = S.bS = nullptr; *rowkind_io = nullptr;
  S.bell_off = S.bell_col = nullptr; S.bell_val = nullptr; S.nslices = 0; S.bell_entries = 0;
  UFE_CUDA(cudaMalloc(&counts, sizeof(int) * (m_loc > 0 ? m_loc : 1)));
  UFE_CUDA(cudaMalloc(rowkind_io, sizeof(int) * (m_loc > 0 ? m_loc : 1)));
  UFE_CUDA(cudaMalloc(&S.ptr, sizeof(int) * (m_loc + 1)));
  if (nt > 0) {
    k_count_rows<<<ufe_div_up(nt, 256), 256, 0, st>>>(t0, nt, nTri, A, M2, TriBI, TriC, bc_mask, counts, *rowkind_io);
    UFE_LAUNCH_CHECK();
  }
  UFE_TRY(ufe_counts_to_ptr(st, m_loc, counts, S.ptr, &S.nnz));
  const size_t nz = S.nnz > 0 ? S.nnz : 1, mb = m_loc > 0 ? m_loc : 1;
  UFE_CUDA(cudaMalloc(&S.ind, sizeof(int) * nz));
  UFE_CUDA(cudaMalloc(&S.val, sizeof(double) * nz));
  UFE_CUDA(cudaMalloc(&S.bb, sizeof(double) * mb));
  UFE_CUDA(cudaMalloc(&S.bS, sizeof(double) * mb));
  if (nt > 0) {
    k_fill_ind<<<ufe_div_up(nt, 256), 256, 0, st>>>(t0, nt, nTri, M2, TriC, *rowkind_io, S.ptr, S.ind);
    UFE_LAUNCH_CHECK();
  }
  S.N = 2 * nTri; S.m_loc = m_loc; S.r1 = 2 * t0 + 1;
  UFE_TRY(ufe_colrange(st, S.nnz, S.ind, &S.jmin, &S.jmax));
  cudaFree(counts);
  // blocked sliced-ELL pattern for the Krylov loop (values are written by k_assemble)
  if (nt > 0) {
    const int nsl = (nt + 31) / 32;
    int *slice_w = nullptr, total = 0;
    UFE_CUDA(cudaMalloc(&slice_w, sizeof(int) * nsl));
    UFE_CUDA(cudaMalloc(&S.bell_off, sizeof(int) * (nsl + 1)));
    k_bell_slice_width<<<ufe_div_up((long long)nsl * 32, 256), 256, 0, st>>>(t0, nt, nTri, M2, TriC, *rowkind_io, slice_w);
    UFE_LAUNCH_CHECK();
    UFE_TRY(ufe_counts_to_ptr(st, nsl, slice_w, S.bell_off, &total));     // 1-based offsets, in entries
    cudaFree(slice_w);
    S.nslices = nsl; S.bell_entries = total;
    UFE_CUDA(cudaMalloc(&S.bell_col, sizeof(int) * (size_t)(total > 0 ? total : 1) * 32));
    UFE_CUDA(cudaMalloc(&S.bell_val, sizeof(double) * (size_t)(total > 0 ? total : 1) * 128));
    UFE_CUDA(cudaMemsetAsync(S.bell_val, 0, sizeof(double) * (size_t)(total > 0 ? total : 1) * 128, st));
    k_bell_fill_cols<<<ufe_div_up((long long)nsl * 32, 256), 256, 0, st>>>(t0, nt, nTri, M2, TriC, *rowkind_io, S.bell_off, S.bell_col);
    UFE_LAUNCH_CHECK();
  }
  return UFE_OK;
}

// one thread per owned triangle -> rows (2ti-1, 2ti)
__global__ void __launch_bounds__(128)
k_assemble(int t0, int nt, int nTri, AssemblyParams A, DevFamilyView M2, const int *__restrict__ TriC,
           const int *__restrict__ rowkind, BCTables T, const int *__restrict__ bc_mask,
           const double *__restrict__ bc_u, const double *__restrict__ bc_v, DivaFields F,
           const int *__restrict__ ptrA, double *__restrict__ val, const int *__restrict__ bell_off,
           double *__restrict__ bval, double *__restrict__ bb, double *__restrict__ bS,
           double *__restrict__ xg, int write_x) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  // this block row's first entry in the blocked sliced-ELL value array (stride 128 per entry)
  double *const brow = bval + (size_t)(bell_off[tl >> 5] - 1) * 128 + (tl & 31);
  if (write_x) {                  // uv_buv interleave (:74-84): initial guess for the Krylov solve
    xg[2 * (size_t)ti] = F.u_vav_b[ti];
    xg[2 * (size_t)ti + 1] = F.v_vav_b[ti];
  }
  const int kind_u = rowkind[2 * tl], kind_v = rowkind[2 * tl + 1];
  if (kind_u == RK_FREE) {        // both rows are free together
    const double N = F.N_b[ti], Nx = F.dN_dx_b[ti], Ny = F.dN_dy_b[ti], beta = F.beta_eff_b[ti];
    const double tdx = F.tau_dx_b[ti], tdy = F.tau_dy_b[ti];
    const int j0 = M2.ptr[tl] - 1, j1 = M2.ptr[tl + 1] - 1;
    const int ku = ptrA[2 * tl] - 1, kv = ptrA[2 * tl + 1] - 1;
    double b_u, b_v;
    if (A.crossterms) { b_u = -tdx; b_v = -tdy; } else { b_u = -tdx / N; b_v = -tdy / N; }
    // pass 1: the 2x2 diagonal block (entry tj == ti)
    double Duu = 1.0, Duv = 0.0, Dvu = 0.0, Dvv = 1.0;
    for (int pass = 0; pass < 2; pass++) {
      double B00 = 1.0, B01 = 0.0, B10 = 0.0, B11 = 1.0;
      if (pass == 1) {
        if (A.pc == UFE_PC_BJACOBI2) {
          const double det = Duu * Dvv - Duv * Dvu;
          B00 = Dvv / det; B01 = -Duv / det; B10 = -Dvu / det; B11 = Duu / det;
        } else {
          B00 = 1.0 / Duu; B11 = 1.0 / Dvv;
        }
      }
      for (int j = j0; j < j1; j++) {
        const int tj = M2.ind[j] - 1;
        if (pass == 0 && tj != ti) continue;
        const double dx = M2.v0[j], dy = M2.v1[j], xx = M2.v2[j], xy = M2.v3[j], yy = M2.v4[j];
        double Au_u, Av_u, Au_v, Av_v;      // row u: (Au_u, Av_u); row v: (Au_v, Av_v)
        if (A.crossterms) {
          Au_u = 4.0 * N * xx + 4.0 * Nx * dx + N * yy + Ny * dy;
          if (tj == ti) Au_u = Au_u - beta;
          Av_u = 3.0 * N * xy + 2.0 * Nx * dy + Ny * dx;
          Av_v = 4.0 * N * yy + 4.0 * Ny * dy + N * xx + Nx * dx;
          if (tj == ti) Av_v = Av_v - beta;
          Au_v = 3.0 * N * xy + 2.0 * Ny * dx + Nx * dy;
        } else {
          Au_u = 4.0 * xx + yy;
          if (tj == ti) Au_u = Au_u - beta / N;
          Av_u = 3.0 * xy;
          Av_v = 4.0 * yy + xx;
          if (tj == ti) Av_v = Av_v - beta / N;
          Au_v = 3.0 * xy;
        }
        if (pass == 0) { Duu = Au_u; Duv = Av_u; Dvu = Au_v; Dvv = Av_v; break; }
        const int o = 2 * (j - j0);
        val[ku + o] = Au_u; val[ku + o + 1] = Av_u;
        val[kv + o] = Au_v; val[kv + o + 1] = Av_v;
        double *const be = brow + (size_t)(j - j0) * 128;
        be[0] = B00 * Au_u + B01 * Au_v; be[32] = B00 * Av_u + B01 * Av_v;
        be[64] = B10 * Au_u + B11 * Au_v; be[96] = B10 * Av_u + B11 * Av_v;
      }
      if (pass == 1) {
        bb[2 * tl] = b_u; bb[2 * tl + 1] = b_v;
        bS[2 * tl] = B00 * b_u + B01 * b_v; bS[2 * tl + 1] = B10 * b_u + B11 * b_v;
      }
    }
    return;
  }
  // boundary / prescribed rows.  Blocked copy: neighbours first (if either row is 'infinite'), then
  // the diagonal block; every scaled diagonal is 1.
  int nn = 0;
  for (int n = 0; n < 3; n++) if (TriC[(size_t)n * nTri + ti] != 0) nn++;
  const double dinf = -1.0 * (double)nn;
  int e = 0;
  if (kind_u == RK_INFINITE || kind_v == RK_INFINITE)
    for (; e < nn; e++) {
      double *const be = brow + (size_t)e * 128;
      be[0] = kind_u == RK_INFINITE ? 1.0 / dinf : 0.0; be[32] = 0.0; be[64] = 0.0;
      be[96] = kind_v == RK_INFINITE ? 1.0 / dinf : 0.0;
    }
  { double *const be = brow + (size_t)e * 128; be[0] = 1.0; be[32] = 0.0; be[64] = 0.0; be[96] = 1.0; }
  for (int uv = 0; uv < 2; uv++) {
    const int kind = uv == 0 ? kind_u : kind_v;
    int k = ptrA[2 * tl + uv] - 1;
    const int r = 2 * tl + uv;
    if (kind == RK_PRESCR) {
      val[k] = 1.0;
      const double b = uv == 0 ? bc_u[ti] : bc_v[ti];
      bb[r] = b; bS[r] = b;
    } else if (kind == RK_INFINITE) {
      for (int n = 0; n < nn; n++) { val[k] = 1.0; k++; }
      val[k] = dinf;
      bb[r] = 0.0; bS[r] = 0.0;
    } else if (kind == RK_ZERO) {
      val[k] = 1.0; bb[r] = 0.0; bS[r] = 0.0;
    } else {   // RK_COPY: periodic_ISMIP-HOM / infinite_SSA_icestream (:555-600)
      val[k] = 1.0;
      const double *prev = uv == 0 ? F.u_b_prev : F.v_b_prev;
      const int slot = T.slot[ti];
      double fixed = 0.0;
      for (int n = 0; n < T.nC_mem; n++) {
        const int tj = T.copy_ti[(size_t)slot * T.nC_mem + n];
        if (tj == 0) continue;
        fixed = fixed + T.copy_w[(size_t)slot * T.nC_mem + n] * prev[tj - 1];
      }
      fixed = (A.visc_it_relax * fixed) + ((1.0 - A.visc_it_relax) * prev[ti]);
      bb[r] = fixed; bS[r] = fixed;
    }
  }
}

int ufe_launch_assemble(cudaStream_t st, int t0, int nt, int nTri, const AssemblyParams &A, DevFamilyView M2,
                        const int *TriC, const int *rowkind, const BCTables &T, const int *bc_mask,
                        const double *bc_u, const double *bc_v, const DivaFields &F, const DevSystem &S,
                        int write_x) {
  if (nt <= 0) return UFE_OK;
  k_assemble<<<ufe_div_up(nt, 128), 128, 0, st>>>(t0, nt, nTri, A, M2, TriC, rowkind, T, bc_mask, bc_u, bc_v, F,
                                                  S.ptr, S.val, S.bell_off, S.bell_val, S.bb, S.bS, S.x, write_x);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

// generic point-Jacobi scaling for an arbitrary square CSR system (L0 entry point)
__global__ void k_scale_generic(int m_loc, int r0, const int *__restrict__ ptr, const int *__restrict__ ind,
                                const double *__restrict__ val, double *__restrict__ valS,
                                const double *__restrict__ bb, double *__restrict__ bS) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_loc) return;
  double d = 1.0;
  for (int k = ptr[r] - 1; k < ptr[r + 1] - 1; k++) if (ind[k] - 1 == r0 + r) d = val[k];
  if (d == 0.0) d = 1.0;
  for (int k = ptr[r] - 1; k < ptr[r + 1] - 1; k++) valS[k] = val[k] / d;
  bS[r] = bb[r] / d;
}

int ufe_launch_scale_generic(cudaStream_t st, const DevSystem &S) {
  if (S.m_loc <= 0) return UFE_OK;
  k_scale_generic<<<ufe_div_up(S.m_loc, 256), 256, 0, st>>>(S.m_loc, S.r1 - 1, S.ptr, S.ind, S.val, S.valS, S.bb, S.bS);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
