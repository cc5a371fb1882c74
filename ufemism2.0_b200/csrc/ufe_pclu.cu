// Block-Jacobi preconditioner with exact sub-domain solves (block cyclic reduction).
//
// The reference preconditions its KSP with PETSc's default block-Jacobi -- one block per MPI
// rank = one contiguous row range = one vertical strip of the x-sorted mesh -- with ILU(0)
// inside each block (src/UPSY/basic/petsc_basic.f90:106-119, no options set).  Here the block
// of a GPU is its own row range as well, but it is solved exactly.  Because the mesh is
// x-sorted, the (already Jacobi-scaled) matrix is banded; with a dense block size
// g >= bandwidth the local matrix is block tridiagonal with K = n_loc/g blocks,
//        L_i x_{i-1} + D_i x_i + U_i x_{i+1} = r_i ,
// and is factorised by block cyclic reduction: at level l the active nodes are
// i_q = (q+1) 2^l - 1; the even positions q are eliminated, the odd ones kept,
//   eliminated j :  Dinv_j = D_j^-1,  P_j = Dinv_j L_j,  Q_j = Dinv_j U_j
//   kept i (a = i - 2^l, b = i + 2^l):
//        D_i <- D_i - L_i Q_a - U_i P_b,   L_i' = -L_i P_a,   U_i' = -U_i Q_b   (next level's couplings)
// so every level is a handful of batched dense launches over all its nodes (log2 K levels in
// all; the couplings L, U of every level are kept), and applying the preconditioner is
// 3 log2 K batched GEMV launches:
//   down:  y_j = Dinv_j r_j,   r_i <- r_i - L_i y_a - U_i y_b        up:  x_j = y_j - P_j x_{j-s} - Q_j x_{j+s}.
// With one GPU the preconditioner is the exact inverse (the Krylov loop then converges in one or
// two steps and acts as iterative refinement).  With several GPUs there are two modes:
//   strip mode     : block Jacobi over the ranks' strips (couplings to other ranks' columns are
//                    dropped), each rank factorises its own strip -- PETSc's bjacobi structure;
//   replicated mode: when the whole system's dense blocks fit one GPU, every rank assembles the
//                    global block-tridiagonal matrix (its own rows + one all-reduce), factorises it
//                    redundantly and applies it to the all-gathered residual, so the preconditioner
//                    stays exact and the Krylov count does not grow with the rank count (small
//                    systems cannot amortise more communication than that).
// Inverses: blocked in-place Gauss-Jordan (32-wide panels, partial pivoting inside the 32 x 32
// pivot block, static perturbation of vanishing pivots; the Krylov iteration absorbs it).
#include <stdlib.h>

#include "ufe_internal.cuh"

#define GT 64          // GEMM tile
#define GK 16
#define NB 32          // Gauss-Jordan panel width

struct PcLevel { int s, n, nE, nK, base; };    // base: first slot of this level's couplings in LS / US

struct PcLU {
  int n_loc = 0, g = 0, K = 0;               // n_loc: rows the factorisation covers (all N in replicated mode)
  int replicated = 0, own_n = 0, own_r0 = 0; // replicated mode: this rank's rows [own_r0, own_r0 + own_n)
  const Comm *comm = nullptr;
  HaloPlan gather;                           // all-gather plan in triangle units (mult 2)
  double *gvec = nullptr;                    // global-length staging vector
  std::vector<PcLevel> lev;
  double *D = nullptr;                       // K blocks: D_i, overwritten by Dinv_i when i is eliminated
  double *LS = nullptr, *US = nullptr;       // couplings of the active nodes of every level: slot base_l + position (< 2K slots)
  double *Pm = nullptr, *Qm = nullptr;       // K blocks, per eliminated node
  double *ipp = nullptr, *colbuf = nullptr;  // Gauss-Jordan scratch: [items][32*32], [items][g][32]
  double *c = nullptr;                       // 2*K*g work vectors (rhs -> solution, staging)
  size_t bytes = 0;
  ufe_nd_solver *nd = nullptr;               // UFE_PC_ND_LU: the multifrontal solver replaces everything above
  bool exact = false;                        // the factorisation covers the whole matrix (not one strip block per rank)
  bool fresh = false;                        // factorised with the current matrix values (cleared when a factorisation is reused)
};

bool ufe_pclu_exact_and_fresh(const PcLU *pc) { return pc && pc->exact && pc->fresh; }
void ufe_pclu_mark_reused(PcLU *pc) { if (pc) pc->fresh = false; }

// operand addressing for the batched kernels: item z of a level works on node
//   node(z) = (2z + 1 + kept) * s - 1 ;   operand block = node + off   (mode 0)
//                                                        = off + z      (mode 1: next level's slot)
//                                                        = off + 2z     (mode 2: this level's slot)
struct Opnd { const double *p; int mode, off; };
__device__ __forceinline__ long long opnd_block(const Opnd &o, int node, int z, int K) {
  if (o.mode == 1) return o.off + z;
  if (o.mode == 2) return o.off + 2 * z;
  const int b = node + o.off;
  return (b < 0 || b >= K) ? -1 : b;
}

// ------------------------------------------------------------------------------------
// matrix -> dense block-tridiagonal storage
// ------------------------------------------------------------------------------------
// rows: local block row t -> t + tshift; columns: bcol - cshift, kept when inside [0, ncols)
__global__ void k_bell_bandwidth(int nt_loc, int tshift, int cshift, int ncols, int nslices, const int *__restrict__ bell_off,
                                 const int *__restrict__ bcol, int *bw) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int off = bell_off[s] - 1, w = bell_off[s + 1] - 1 - off;
  const int rl = s * 32 + lane, r = rl + tshift;
  int b = 0;
  if (rl < nt_loc)
    for (int e = 0; e < w; e++) {
      const int c = bcol[(size_t)(off + e) * 32 + lane] - cshift;
      if (c < 0 || c >= ncols) continue;         // strip mode: another rank's column, dropped by block Jacobi
      const int d = c > r ? c - r : r - c;
      b = max(b, 2 * d + 1);
    }
  for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if (lane == 0) atomicMax(bw, b);
}

__global__ void k_bell_to_blocks(int nt_loc, int tshift, int cshift, int ncols, int nslices, const int *__restrict__ bell_off,
                                 const int *__restrict__ bcol, const double *__restrict__ bval, int g,
                                 double *__restrict__ D, double *__restrict__ L, double *__restrict__ U) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int off = bell_off[s] - 1, w = bell_off[s + 1] - 1 - off;
  const int tl = s * 32 + lane, t = tl + tshift;
  if (tl >= nt_loc) return;
  const size_t gg = (size_t)g * g;
  for (int e = 0; e < w; e++) {
    const int ct = bcol[(size_t)(off + e) * 32 + lane] - cshift;
    if (ct < 0 || ct >= ncols) continue;
    const double *p = bval + (size_t)(off + e) * 128 + lane;
    const double a[4] = {p[0], p[32], p[64], p[96]};
    for (int q = 0; q < 4; q++) {
      if (a[q] == 0.0) continue;                 // also skips the zero padding entries of the slice
      const int r = 2 * t + (q >> 1), c = 2 * ct + (q & 1);
      const int kr = r / g, kc = c / g;
      const size_t o = (size_t)kr * gg + (size_t)(r - kr * g) * g + (c - kc * g);
      if (kc == kr) D[o] = a[q];
      else if (kc == kr - 1) L[o] = a[q];
      else if (kc == kr + 1) U[o] = a[q];
    }
  }
}

// the same from a scaled CSR matrix (generic L0 path); D == nullptr: bandwidth only
__global__ void k_csr_to_blocks(int m_loc, int r0, const int *__restrict__ ptr, const int *__restrict__ ind,
                                const double *__restrict__ val, int g, double *__restrict__ D,
                                double *__restrict__ L, double *__restrict__ U, int *bw) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_loc) return;
  const size_t gg = (size_t)g * g;
  int b = 0;
  for (int k = ptr[r] - 1; k < ptr[r + 1] - 1; k++) {
    const int c = ind[k] - 1 - r0;
    if (c < 0 || c >= m_loc) continue;
    b = max(b, c > r ? c - r : r - c);
    if (!D) continue;
    const int kr = r / g, kc = c / g;
    const size_t o = (size_t)kr * gg + (size_t)(r - kr * g) * g + (c - kc * g);
    if (kc == kr) atomicAdd(D + o, val[k]);
    else if (kc == kr - 1) atomicAdd(L + o, val[k]);
    else if (kc == kr + 1) atomicAdd(U + o, val[k]);
  }
  if (bw) atomicMax(bw, b);
}

__global__ void k_pad_identity(int n_loc, int g, int K, double *__restrict__ D) {
  const int i = n_loc + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * g) return;
  const int k = i / g, li = i - k * g;
  D[(size_t)k * g * g + (size_t)li * g + li] = 1.0;
}

// ------------------------------------------------------------------------------------
// batched dense GEMM on g x g row-major blocks: C = beta*C + alpha*A*B for every item of a
// level (blockIdx.z).  128x64 tile, 256 threads, 8x4 register micro-tile, register-staged
// prefetch of the next k-tile; g multiple of 64.
// Items whose A or B operand does not exist (no neighbour on that side) are skipped.
// ------------------------------------------------------------------------------------
// tile = (16 MI) x (16 NI); MI x NI register micro-tile per thread (rows ty + 16 p, columns tx + 16 q)
// `pair`: two independent products per item in one launch (blockIdx.z & 1 selects the triple)
struct GemmOps { Opnd A, B, C; };
template <int MI, int NI>
__global__ void __launch_bounds__(256)
k_bgemm(int g, int K, int s, int kept, GemmOps o0, GemmOps o1, int pair, double alpha, double beta) {
  constexpr int TMv = 16 * MI, TNv = 16 * NI, BROWS = 256 / TNv;     // B rows staged per pass
  const int z = pair ? (int)blockIdx.z >> 1 : (int)blockIdx.z, node = (2 * z + 1 + kept) * s - 1;
  const GemmOps &o = (pair && (blockIdx.z & 1)) ? o1 : o0;
  const Opnd &A = o.A, &B = o.B, &C = o.C;
  const long long ia = opnd_block(A, node, z, K), ib = opnd_block(B, node, z, K), ic = opnd_block(C, node, z, K);
  if (ia < 0 || ib < 0 || ic < 0) return;
  const size_t gg = (size_t)g * g;
  const double *__restrict__ Ab = A.p + ia * gg, *__restrict__ Bb = B.p + ib * gg;
  double *Cb = const_cast<double *>(C.p) + ic * gg;
  __shared__ double sA[GK][TMv + 1], sB[GK][TNv + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * TMv, j0 = blockIdx.x * TNv;
  // global -> register staging of the next k-tile (hides the load latency behind the FMAs)
  const int a_kk = threadIdx.x & 15, a_ii = threadIdx.x >> 4;             // A: rows a_ii + 16 q, column k0 + a_kk
  const int b_jj = threadIdx.x & (TNv - 1), b_kk = threadIdx.x / TNv;     // B: rows k0 + b_kk + BROWS q, column j0 + b_jj
  double ra[MI], rb[NI];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < MI; q++) { const int i = i0 + a_ii + 16 * q; ra[q] = i < g ? Ab[(size_t)i * g + k0 + a_kk] : 0.0; }
#pragma unroll
    for (int q = 0; q < NI; q++) rb[q] = Bb[(size_t)(k0 + b_kk + BROWS * q) * g + j0 + b_jj];
  };
  double acc[MI][NI] = {};
  fetch(0);
  for (int k0 = 0; k0 < g; k0 += GK) {
#pragma unroll
    for (int q = 0; q < MI; q++) sA[a_kk][a_ii + 16 * q] = ra[q];
#pragma unroll
    for (int q = 0; q < NI; q++) sB[b_kk + BROWS * q][b_jj] = rb[q];
    __syncthreads();
    if (k0 + GK < g) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; kk++) {
      double a[MI], b[NI];
#pragma unroll
      for (int q = 0; q < MI; q++) a[q] = sA[kk][ty + 16 * q];       // two addresses per warp: broadcast
#pragma unroll
      for (int q = 0; q < NI; q++) b[q] = sB[kk][tx + 16 * q];       // 16 consecutive doubles: conflict-free
#pragma unroll
      for (int p = 0; p < MI; p++)
#pragma unroll
        for (int q = 0; q < NI; q++) acc[p][q] += a[p] * b[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < MI; p++) {
    const int i = i0 + ty + 16 * p;
    if (i >= g) continue;
#pragma unroll
    for (int q = 0; q < NI; q++) {
      const size_t o = (size_t)i * g + j0 + tx + 16 * q;
      Cb[o] = (beta == 0.0 ? 0.0 : beta * Cb[o]) + alpha * acc[p][q];
    }
  }
}

// picks the tile so that small batches (upper reduction levels) still fill the GPU
static void launch_bgemm2(cudaStream_t st, int items, int g, int K, int s, int kept, const GemmOps &o0, const GemmOps &o1, int pair,
                          double alpha, double beta) {
  if (items <= 0) return;
  const int nz = pair ? 2 * items : items;
  const long long big = (long long)nz * ((g + 127) / 128) * (g / 64), mid = (long long)nz * (g / 64) * (g / 64);
  if (big >= 296) k_bgemm<8, 4><<<dim3(g / 64, (g + 127) / 128, nz), 256, 0, st>>>(g, K, s, kept, o0, o1, pair, alpha, beta);
  else if (mid >= 296) k_bgemm<4, 4><<<dim3(g / 64, g / 64, nz), 256, 0, st>>>(g, K, s, kept, o0, o1, pair, alpha, beta);
  else k_bgemm<2, 2><<<dim3(g / 32, g / 32, nz), 256, 0, st>>>(g, K, s, kept, o0, o1, pair, alpha, beta);
  g_launch_count++;
}
// picks the tile so that small batches (upper reduction levels) still fill the GPU
static void launch_bgemm(cudaStream_t st, int items, int g, int K, int s, int kept, Opnd A, Opnd B, Opnd C, double alpha,
                         double beta) {
  const GemmOps o{A, B, C};
  launch_bgemm2(st, items, g, K, s, kept, o, o, 0, alpha, beta);
}

// zero the blocks of a level's items (for outputs whose producing GEMM may be skipped)
__global__ void k_bzero(int g, int K, int s, int kept, Opnd C) {
  const int z = blockIdx.z, node = (2 * z + 1 + kept) * s - 1;
  const long long ic = opnd_block(C, node, z, K);
  if (ic < 0) return;
  double *Cb = const_cast<double *>(C.p) + ic * (size_t)g * g;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)g * g; i += (size_t)gridDim.x * blockDim.x) Cb[i] = 0.0;
}

// ------------------------------------------------------------------------------------
// Blocked in-place Gauss-Jordan inversion of D_j for all eliminated nodes of a level.
// For pivot block P:  Ipp = inv(A_PP);  A_PJ <- Ipp A_PJ (J != P);  A_IJ <- A_IJ - A_IP A_PJ
// (I,J != P);  A_IP <- -A_IP Ipp;  A_PP <- Ipp.
// ------------------------------------------------------------------------------------
// 32 x 32 inverse in shared memory by 1024 threads (i, j): Gauss-Jordan on [S | I2] with partial
// pivoting tracked as a row permutation (no physical swaps, three barriers per pivot).  On return
// Ip = S^-1.  The inverse is unique, so the pivoting leaves no trace outside this routine.
__device__ __forceinline__ void invert32(double (*S)[NB + 1], double (*I2)[NB + 1], double (*Ip)[NB + 1], int *perm,
                                         int i, int j) {
  I2[i][j] = (i == j) ? 1.0 : 0.0;
  if (i == 0) perm[j] = j;
  __syncthreads();
  for (int p = 0; p < NB; p++) {
    // threshold pivoting: the matrix is Jacobi-scaled (unit diagonal), so the natural pivot is
    // almost always of order one; the arg-max search (one warp, ~40 % of a step) and its barrier
    // are only paid when it is not.  The test reads one shared value, so it is uniform.
    if (fabs(S[perm[p]][p]) < 0.05) {
      __syncthreads();                   // everybody has evaluated the test before perm changes
      if (i == 0) {                      // warp 0: arg max over logical rows r >= p of |S[perm[r]][p]|
        double v = (j >= p) ? fabs(S[perm[j]][p]) : -1.0;
        int r = j;
        for (int o = 16; o > 0; o >>= 1) {
          const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
          const int r2 = __shfl_xor_sync(0xffffffffu, r, o);
          if (v2 > v || (v2 == v && r2 < r)) { v = v2; r = r2; }
        }
        if (j == 0) { const int t = perm[p]; perm[p] = perm[r]; perm[r] = t; }
      }
      __syncthreads();
    }
    const int P = perm[p];              // physical pivot row
    double piv = S[P][p];
    if (fabs(piv) < 1e-13) piv = piv < 0.0 ? -1e-13 : 1e-13;      // static perturbation
    const double d = 1.0 / piv;
    const double f = S[i][p];
    const double sp = S[P][j] * d, ip = I2[P][j] * d;
    __syncthreads();
    if (i == P) { S[i][j] = sp; I2[i][j] = ip; }
    else { S[i][j] -= f * sp; I2[i][j] -= f * ip; }
    __syncthreads();
  }
  Ip[i][j] = I2[perm[i]][j];
  __syncthreads();
}

// pivot block inverse, one CTA per eliminated node: ipp <- inv(A_PP).  (Measured alternatives: a
// single warp holding the rows in registers with shuffle broadcasts, 71 us; a single warp on shared
// memory, 47 us; this 1024-thread version, three barriers per pivot: fastest of the three.)
__global__ void __launch_bounds__(1024)
k_gjb_pivot(int g, int s, int b, const double *__restrict__ D, double *__restrict__ ipp) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  __shared__ double W1[NB][NB + 1], W2[NB][NB + 1], Ip[NB][NB + 1];
  __shared__ int perm[NB];
  const int j = threadIdx.x & 31, i = threadIdx.x >> 5;
  W1[i][j] = D[(size_t)node * g * g + (size_t)(b * NB + i) * g + b * NB + j];
  __syncthreads();
  invert32(W1, W2, Ip, perm, i, j);
  ipp[(size_t)z * NB * NB + i * NB + j] = Ip[i][j];
}

// Panel step for pivot block b:  y = 0: row panel tile (P, J = blockIdx.x): A_PJ <- Ipp A_PJ;
// y = 1: column panel tile (I = blockIdx.x, P): colbuf <- A_IP (old), A_IP <- -A_IP Ipp.
// A_PP itself is replaced by Ipp in the update kernel.
__global__ void __launch_bounds__(1024)
k_gjb_panel(int g, int s, int b, double *__restrict__ D, const double *__restrict__ ipp, double *__restrict__ colbuf) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  __shared__ double X[NB][NB + 1], Ip[NB][NB + 1];
  const int j = threadIdx.x & 31, i = threadIdx.x >> 5, q = blockIdx.x;
  if (q == b) return;
  double *A = D + (size_t)node * g * g;
  Ip[i][j] = ipp[(size_t)z * NB * NB + i * NB + j];
  if (blockIdx.y == 0) {
    double *T = A + (size_t)(b * NB) * g + q * NB;             // tile (P, J=q)
    X[i][j] = T[(size_t)i * g + j];
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NB; k++) sum += Ip[i][k] * X[k][j];
    T[(size_t)i * g + j] = sum;
  } else {
    double *T = A + (size_t)(q * NB) * g + b * NB;             // tile (I=q, P)
    X[i][j] = T[(size_t)i * g + j];
    colbuf[((size_t)z * g + q * NB + i) * NB + j] = X[i][j];   // old A_IP for the trailing update
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NB; k++) sum += X[i][k] * Ip[k][j];
    T[(size_t)i * g + j] = -sum;
  }
}

// The same panel step with the pivot-block inverse folded in: every CTA inverts A_PP itself.
// Redundant work, but on the small upper levels (a handful of nodes) the GPU is idle anyway and
// it removes a serial launch per panel; used when the whole grid is resident at once.
__global__ void __launch_bounds__(1024)
k_gjb_panel_inv(int g, int s, int b, double *__restrict__ D, double *__restrict__ ipp, double *__restrict__ colbuf) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  __shared__ double X[NB][NB + 1], Ip[NB][NB + 1], W1[NB][NB + 1], W2[NB][NB + 1];
  __shared__ int perm[NB];
  const int j = threadIdx.x & 31, i = threadIdx.x >> 5, q = blockIdx.x;
  if (blockIdx.y == 1 && q == b) return;
  double *A = D + (size_t)node * g * g;
  W1[i][j] = A[(size_t)(b * NB + i) * g + b * NB + j];     // A_PP is not written in this kernel
  __syncthreads();
  invert32(W1, W2, Ip, perm, i, j);
  if (blockIdx.y == 0) {
    if (q == b) { ipp[(size_t)z * NB * NB + i * NB + j] = Ip[i][j]; return; }   // the update kernel copies it into place
    double *T = A + (size_t)(b * NB) * g + q * NB;             // tile (P, J=q)
    X[i][j] = T[(size_t)i * g + j];
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NB; k++) sum += Ip[i][k] * X[k][j];
    T[(size_t)i * g + j] = sum;
  } else {
    double *T = A + (size_t)(q * NB) * g + b * NB;             // tile (I=q, P)
    X[i][j] = T[(size_t)i * g + j];
    colbuf[((size_t)z * g + q * NB + i) * NB + j] = X[i][j];   // old A_IP for the trailing update
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NB; k++) sum += X[i][k] * Ip[k][j];
    T[(size_t)i * g + j] = -sum;
  }
}

// trailing update, 64 x 64 tile per CTA (256 threads, 4 x 4 micro-tile), inner dimension 32
__global__ void __launch_bounds__(256)
k_gjb_update(int g, int s, int b, double *__restrict__ D, const double *__restrict__ colbuf,
             const double *__restrict__ ipp) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  __shared__ double Cb[GT][NB + 1], R[NB][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  double *A = D + (size_t)node * g * g;
  for (int q = threadIdx.x; q < GT * NB; q += 256) {
    const int ii = q / NB, kk = q % NB;
    Cb[ii][kk] = ((i0 + ii) / NB == b) ? 0.0 : colbuf[((size_t)z * g + i0 + ii) * NB + kk];
    const int k2 = q / GT, jj = q % GT;
    R[k2][jj] = A[(size_t)(b * NB + k2) * g + j0 + jj];
  }
  __syncthreads();
  if ((int)blockIdx.y == (b * NB) / GT && (int)blockIdx.x == (b * NB) / GT)       // A_PP <- Ipp
    for (int q = threadIdx.x; q < NB * NB; q += 256)
      A[(size_t)(b * NB + q / NB) * g + b * NB + q % NB] = ipp[(size_t)z * NB * NB + q];
  double acc[4][4] = {};
#pragma unroll
  for (int kk = 0; kk < NB; kk++) {
    double a[4], r[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { a[q] = Cb[ty + 16 * q][kk]; r[q] = R[kk][tx + 16 * q]; }
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[p][q] += a[p] * r[q];
  }
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int i = i0 + ty + 16 * p;
    if (i / NB == b) continue;                 // pivot rows were handled by the panel kernel
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int j = j0 + tx + 16 * q;
      if (j / NB == b) continue;               // pivot columns likewise
      A[(size_t)i * g + j] -= acc[p][q];
    }
  }
}

// ------------------------------------------------------------------------------------
// solve kernels: one warp per row of a g x g block
// ------------------------------------------------------------------------------------
// y_j = Dinv_j r_j for the eliminated nodes of a level
__global__ void __launch_bounds__(256)
k_bcr_y(int g, int s, const double *__restrict__ D, const double *__restrict__ c, double *__restrict__ y) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= g) return;
  const double *Mr = D + (size_t)node * g * g + (size_t)row * g, *r = c + (size_t)node * g;
  double sum = 0.0;
  for (int j = lane; j < g; j += 32) sum += Mr[j] * r[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) y[(size_t)node * g + row] = sum;
}

// down sweep, kept nodes:  r_i <- r_i - L_i y_{i-s} - U_i y_{i+s}
__global__ void __launch_bounds__(256)
k_bcr_down(int g, int K, int s, int base, const double *__restrict__ LS, const double *__restrict__ US,
           const double *__restrict__ y, double *__restrict__ c) {
  const int z = blockIdx.z, node = (2 * z + 2) * s - 1;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= g) return;
  const size_t mo = (size_t)(base + 2 * z + 1) * g * g + (size_t)row * g;
  const double *ya = y + (size_t)(node - s) * g;
  const bool hasb = node + s < K;
  const double *yb = y + (size_t)(hasb ? node + s : node) * g;
  double sum = 0.0;
  for (int j = lane; j < g; j += 32) {
    sum += LS[mo + j] * ya[j];
    if (hasb) sum += US[mo + j] * yb[j];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) c[(size_t)node * g + row] -= sum;
}

// up sweep, eliminated nodes:  x_j = y_j - P_j x_{j-s} - Q_j x_{j+s}   (solutions live in c)
__global__ void __launch_bounds__(256)
k_bcr_up(int g, int K, int s, const double *__restrict__ Pm, const double *__restrict__ Qm,
         const double *__restrict__ y, double *__restrict__ c) {
  const int z = blockIdx.z, node = (2 * z + 1) * s - 1;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= g) return;
  const size_t mo = (size_t)node * g * g + (size_t)row * g;
  const bool hasa = node - s >= 0, hasb = node + s < K;
  const double *xa = c + (size_t)(hasa ? node - s : node) * g, *xb = c + (size_t)(hasb ? node + s : node) * g;
  double sum = 0.0;
  if (hasa || hasb)
    for (int j = lane; j < g; j += 32) {
      if (hasa) sum += Pm[mo + j] * xa[j];
      if (hasb) sum += Qm[mo + j] * xb[j];
    }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) c[(size_t)node * g + row] = y[(size_t)node * g + row] - sum;
}

__global__ void k_vec_in(int n, int total, const double *__restrict__ r, double *__restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) c[i] = i < n ? r[i] : 0.0;
}
__global__ void k_vec_place(int n, int r0, const double *__restrict__ r, double *__restrict__ gv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gv[r0 + i] = r[i];
}
__global__ void k_vec_out(int n, const double *__restrict__ c, double *__restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = c[i];
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
void ufe_pclu_free(PcLU *pc) {
  if (!pc) return;
  ufe_nd_solver_free(pc->nd);
  cudaFree(pc->D); cudaFree(pc->LS); cudaFree(pc->US); cudaFree(pc->Pm); cudaFree(pc->Qm);
  cudaFree(pc->ipp); cudaFree(pc->colbuf); cudaFree(pc->c); cudaFree(pc->gvec);
  delete pc;
}

int ufe_pclu_setup(cudaStream_t st, const DevSystem &S, int replicate, size_t max_bytes, PcLU **out, const Comm *comm,
                   const HaloPlan *gather_all) {
  *out = nullptr;
  const bool repl = replicate && comm && comm->nranks > 1 && S.bell_val;
  const int tshift = repl ? (S.r1 - 1) / 2 : 0, cshift = repl ? 0 : (S.r1 - 1) / 2, ncols = repl ? S.N / 2 : S.m_loc / 2;
  const int n_cover = repl ? S.N : S.m_loc;
  int *d_bw = nullptr, bw = 0;
  UFE_CUDA(cudaMalloc(&d_bw, sizeof(int)));
  UFE_CUDA(cudaMemsetAsync(d_bw, 0, sizeof(int), st));
  if (S.bell_val)
    k_bell_bandwidth<<<ufe_div_up((long long)S.nslices * 32, 256), 256, 0, st>>>(S.m_loc / 2, tshift, cshift, ncols, S.nslices, S.bell_off, S.bell_col, d_bw);
  else
    k_csr_to_blocks<<<ufe_div_up(S.m_loc, 256), 256, 0, st>>>(S.m_loc, S.r1 - 1, S.ptr, S.ind, S.valS, 1, nullptr, nullptr, nullptr, d_bw);
  UFE_LAUNCH_CHECK();
  if (repl) { UFE_NCCL(ncclAllReduce(d_bw, d_bw, 1, ncclInt32, ncclMax, comm->nccl, st)); g_launch_count++; }
  UFE_CUDA(cudaMemcpyAsync(&bw, d_bw, sizeof(int), cudaMemcpyDeviceToHost, st));
  UFE_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_bw);
  PcLU *pc = new PcLU();
  pc->n_loc = n_cover;
  pc->replicated = repl ? 1 : 0; pc->own_n = S.m_loc; pc->own_r0 = S.r1 - 1; pc->comm = comm;
  pc->exact = repl || S.m_loc == S.N;
  if (repl) pc->gather = *gather_all;
  int g = ((bw > 0 ? bw : 1) + GT - 1) / GT * GT;
  if (g > n_cover) g = (n_cover + GT - 1) / GT * GT;
  if (g < GT) g = GT;
  pc->g = g;
  pc->K = (n_cover + g - 1) / g;
  if (pc->K < 1) pc->K = 1;
  int slots = 0, maxE = 0;
  for (int s = 1;; s *= 2) {
    PcLevel lv;
    lv.s = s; lv.n = pc->K / s; lv.nE = (lv.n + 1) / 2; lv.nK = lv.n / 2; lv.base = slots;
    if (lv.n < 1) break;
    slots += lv.n;
    if (lv.nE > maxE) maxE = lv.nE;
    pc->lev.push_back(lv);
    if (lv.n == 1) break;
  }
  const size_t gg = (size_t)g * g, blk = gg * pc->K * sizeof(double);
  const size_t slot_bytes = gg * (slots > 0 ? slots : 1) * sizeof(double);
  pc->bytes = 3 * blk + 2 * slot_bytes;
  if (pc->bytes > max_bytes) {
    ufe_set_error("bjacobi_lu preconditioner needs %.1f GB (bandwidth %d -> dense block %d, %d blocks) which exceeds the %.1f GB budget; "
                  "use 'bjacobi2' or 'jacobi' for this mesh", pc->bytes / 1e9, bw, g, pc->K, max_bytes / 1e9);
    delete pc;
    return UFE_ERR_INVALID;
  }
  UFE_CUDA(cudaMalloc(&pc->D, blk));
  UFE_CUDA(cudaMalloc(&pc->Pm, blk)); UFE_CUDA(cudaMalloc(&pc->Qm, blk));
  UFE_CUDA(cudaMalloc(&pc->LS, slot_bytes));
  UFE_CUDA(cudaMalloc(&pc->US, slot_bytes));
  UFE_CUDA(cudaMalloc(&pc->ipp, sizeof(double) * (size_t)maxE * NB * NB));
  UFE_CUDA(cudaMalloc(&pc->colbuf, sizeof(double) * (size_t)maxE * g * NB));
  UFE_CUDA(cudaMalloc(&pc->c, sizeof(double) * (size_t)pc->K * g * 2));
  if (repl) UFE_CUDA(cudaMalloc(&pc->gvec, sizeof(double) * (size_t)S.N));
  *out = pc;
  return UFE_OK;
}

// wide meshes: the banded blocks do not fit, the nested-dissection multifrontal solver does (ufe_nd_numeric.cu)
int ufe_pclu_setup_nd(cudaStream_t st, const DevSystem &S, const Comm *comm, int nT, const double *gcx, const double *gcy, PcLU **out) {
  *out = nullptr;
  const char *e = getenv("UFE_ND_LEAF");
  const int leaf = e && atoi(e) > 0 ? atoi(e) : 64;
  PcLU *pc = new PcLU();
  pc->n_loc = S.m_loc;
  const int rc = ufe_nd_pc_create(st, S, comm, nT, gcx, gcy, leaf, &pc->nd);
  if (rc != UFE_OK) { delete pc; return rc; }
  pc->exact = true;
  *out = pc;
  return UFE_OK;
}

void ufe_pclu_set_point_scaling(PcLU *pc) { if (pc && pc->nd) ufe_nd_pc_set_point_scaling(pc->nd); }

int ufe_pclu_factor(cudaStream_t st, const DevSystem &S, PcLU *pc) {
  pc->fresh = true;
  if (pc->nd) return ufe_nd_pc_factor(st, pc->nd, S.val);
  const int g = pc->g, K = pc->K;
  const size_t blk = (size_t)g * g * K * sizeof(double);
  UFE_CUDA(cudaMemsetAsync(pc->D, 0, blk, st));
  UFE_CUDA(cudaMemsetAsync(pc->LS, 0, blk, st));      // level 0 occupies the first K slots (slot = node)
  UFE_CUDA(cudaMemsetAsync(pc->US, 0, blk, st));
  if (S.bell_val)
    k_bell_to_blocks<<<ufe_div_up((long long)S.nslices * 32, 256), 256, 0, st>>>(S.m_loc / 2, pc->replicated ? (S.r1 - 1) / 2 : 0,
                                                                                 pc->replicated ? 0 : (S.r1 - 1) / 2,
                                                                                 pc->replicated ? S.N / 2 : S.m_loc / 2, S.nslices,
                                                                                 S.bell_off, S.bell_col, S.bell_val, g, pc->D, pc->LS, pc->US);
  else
    k_csr_to_blocks<<<ufe_div_up(S.m_loc, 256), 256, 0, st>>>(S.m_loc, S.r1 - 1, S.ptr, S.ind, S.valS, g, pc->D, pc->LS, pc->US, nullptr);
  UFE_LAUNCH_CHECK();
  if (K * g > pc->n_loc && (!pc->replicated || pc->comm->rank == 0)) {
    k_pad_identity<<<ufe_div_up(K * g - pc->n_loc, 256), 256, 0, st>>>(pc->n_loc, g, K, pc->D); UFE_LAUNCH_CHECK();
  }
  if (pc->replicated) {      // every rank contributed its own rows: sum = the global block-tridiagonal matrix
    const size_t cnt = (size_t)g * g * K;
    UFE_NCCL(ncclGroupStart());
    UFE_NCCL(ncclAllReduce(pc->D, pc->D, cnt, ncclDouble, ncclSum, pc->comm->nccl, st));
    UFE_NCCL(ncclAllReduce(pc->LS, pc->LS, cnt, ncclDouble, ncclSum, pc->comm->nccl, st));
    UFE_NCCL(ncclAllReduce(pc->US, pc->US, cnt, ncclDouble, ncclSum, pc->comm->nccl, st));
    UFE_NCCL(ncclGroupEnd());
    g_launch_count += 3;
  }
  const int nbk = g / NB, nt = g / GT;
  for (size_t l = 0; l < pc->lev.size(); l++) {
    const PcLevel &lv = pc->lev[l];
    const int s = lv.s;
    // eliminated nodes: Dinv (in place), P = Dinv L, Q = Dinv U
    for (int b = 0; b < nbk; b++) {
      if (2 * nbk * lv.nE <= 148) {          // small level: one wave of CTAs, fold the pivot inverse into the panel step
        k_gjb_panel_inv<<<dim3(nbk, 2, lv.nE), 1024, 0, st>>>(g, s, b, pc->D, pc->ipp, pc->colbuf);
        g_launch_count += 2;
      } else {
        k_gjb_pivot<<<dim3(1, 1, lv.nE), 1024, 0, st>>>(g, s, b, pc->D, pc->ipp);
        k_gjb_panel<<<dim3(nbk, 2, lv.nE), 1024, 0, st>>>(g, s, b, pc->D, pc->ipp, pc->colbuf);
        g_launch_count += 3;
      }
      k_gjb_update<<<dim3(nt, nt, lv.nE), 256, 0, st>>>(g, s, b, pc->D, pc->colbuf, pc->ipp);
    }
    const Opnd Dn{pc->D, 0, 0}, Pn{pc->Pm, 0, 0}, Qn{pc->Qm, 0, 0};
    const Opnd Le{pc->LS, 2, lv.base}, Ue{pc->US, 2, lv.base};               // this level, eliminated (position 2z)
    launch_bgemm2(st, lv.nE, g, K, s, 0, GemmOps{Dn, Le, Pn}, GemmOps{Dn, Ue, Qn}, 1, 1.0, 0.0);     // P = Dinv L, Q = Dinv U
    if (lv.nK > 0) {
      const int nbase = pc->lev[l + 1].base;
      const Opnd Lk{pc->LS, 2, lv.base + 1}, Uk{pc->US, 2, lv.base + 1};     // this level, kept (position 2z+1)
      const Opnd Pa{pc->Pm, 0, -s}, Pb{pc->Pm, 0, s}, Qa{pc->Qm, 0, -s}, Qb{pc->Qm, 0, s};
      const Opnd L2n{pc->LS, 1, nbase}, U2n{pc->US, 1, nbase};               // next level (position z)
      launch_bgemm(st, lv.nK, g, K, s, 1, Lk, Qa, Dn, -1.0, 1.0);       // D_i -= L_i Q_a
      launch_bgemm(st, lv.nK, g, K, s, 1, Uk, Pb, Dn, -1.0, 1.0);       // D_i -= U_i P_b
      k_bzero<<<dim3(32, 1, lv.nK), 256, 0, st>>>(g, K, s, 1, U2n);     // right neighbour may not exist
      launch_bgemm2(st, lv.nK, g, K, s, 1, GemmOps{Lk, Pa, L2n}, GemmOps{Uk, Qb, U2n}, 1, -1.0, 0.0);   // L_i' = -L_i P_a, U_i' = -U_i Q_b
      g_launch_count += 1;
    }
    if (cudaGetLastError() != cudaSuccess) { ufe_set_error("bjacobi_lu factorisation launch failed"); return UFE_ERR_CUDA; }
  }
  return UFE_OK;
}

// z = M^-1 r  (r, z owned-length vectors; may alias)
int ufe_pclu_apply(cudaStream_t st, PcLU *pc, const double *r, double *z) {
  if (pc->nd) return ufe_nd_pc_apply(st, pc->nd, r, z);
  const int g = pc->g, K = pc->K, total = K * g;
  double *c = pc->c, *y = pc->c + total;
  if (pc->replicated) {      // all-gather the distributed residual, solve globally, keep the own slice
    k_vec_place<<<ufe_div_up(pc->own_n, 256), 256, 0, st>>>(pc->own_n, pc->own_r0, r, pc->gvec);
    UFE_LAUNCH_CHECK();
    UFE_TRY(ufe_halo_exchange(st, *pc->comm, pc->gather, pc->gvec, 0, 1, 2));
    r = pc->gvec;
  }
  k_vec_in<<<ufe_div_up(total, 256), 256, 0, st>>>(pc->n_loc, total, r, c);
  UFE_LAUNCH_CHECK();
  for (const PcLevel &lv : pc->lev) {
    k_bcr_y<<<dim3(g / 8, 1, lv.nE), 256, 0, st>>>(g, lv.s, pc->D, c, y);
    if (lv.nK > 0) k_bcr_down<<<dim3(g / 8, 1, lv.nK), 256, 0, st>>>(g, K, lv.s, lv.base, pc->LS, pc->US, y, c);
    g_launch_count += lv.nK > 0 ? 2 : 1;
  }
  for (int l = (int)pc->lev.size() - 1; l >= 0; l--) {
    const PcLevel &lv = pc->lev[l];
    k_bcr_up<<<dim3(g / 8, 1, lv.nE), 256, 0, st>>>(g, K, lv.s, pc->Pm, pc->Qm, y, c);
    g_launch_count++;
  }
  if (cudaGetLastError() != cudaSuccess) { ufe_set_error("bjacobi_lu apply launch failed"); return UFE_ERR_CUDA; }
  if (pc->replicated) k_vec_out<<<ufe_div_up(pc->own_n, 256), 256, 0, st>>>(pc->own_n, c + pc->own_r0, z);
  else k_vec_out<<<ufe_div_up(pc->n_loc, 256), 256, 0, st>>>(pc->n_loc, c, z);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
