// Numeric phase of the multifrontal nested-dissection solver (device code; DESIGN.md section 3, profiles/r2_notes.md).
//
// Exact solve of the DIVA / SSA stiffness system  A x = b  (solve_linearised_SSA_DIVA.f90:159 hands exactly this
// system to PETSc, petsc_basic.f90:32-141).  The symbolic analysis (ufe_nd.cu) gives a binary elimination tree; every
// tree node owns one dense front  [sep; bnd] x [sep; bnd]  (scalar unknowns: triangle t -> 2t, 2t+1,
// mesh_translation_tables.f90:181-198).
//
// Layout in HBM.  Every front has its OWN shape: p = its pivot count rounded up to 32 (identity-padded), nb boundary
// unknowns, G = p + nb rows, leading dimension ld = G rounded up to 8; fronts of one tree level are stored back to
// back, sorted by p (descending) so that the fronts still active at pivot step b form a prefix of the level.
//
// Factorisation, deepest level first: right-looking block LU of the first p rows / columns, 32-wide pivot blocks,
//      D_b^-1 (32 x 32 inverse in registers, partial pivoting inside the block; kept in a side buffer),
//      L_ib = A_ib D_b^-1  (i > b: later pivot rows AND the boundary rows),        [k_mf_panel, k_mf_panel2: two blocks per outer step]
//      A_ij -= L_ib A_bj   (i, j > b; K = 64 per pass, 8 x 8 / 8 x 4 register tiles) [k_mf_update]
// which leaves L, U in place and the Schur complement in the boundary block.  Fronts with many pivots: the steps leave
// the boundary block F22 alone and it is updated once, F22 -= L21 U12 with K = p, on the fp64 tensor cores
// (mma.sync.m8n8k4.f64)                                                              [k_mf_update_mma]
// The parent PULLS the Schur complements of its two children through inverse index maps (child 0 then child 1: fixed
// summation order, bit-reproducible); for large problems it WRITES the parent (contributions or 0) so that the fronts
// need no memset, and the matrix entries of the level are added afterwards.          [k_mf_extend<STORE>, k_mf_assemble]
// Solve: one CTA per front and one launch per level and direction: forward  y = L^-1 w  (boundary rows collect the
// update for the parent, which pulls them), backward  x_s = U11^-1 (y - U12 x_b)  with x_b read from the parent.
//
// Several GPUs (one rank per GPU): rank r owns the sub-tree below the r-th node of level log2(P) and the nodes above it
// on its left spine (root: rank 0; level 1: ranks 0, P/2; ...).  A child owned by another rank sends its packed Schur
// complement (factorisation) and its boundary vector (forward sweep) to the parent's owner and receives the parent's
// solution vector (backward sweep): ncclSend / ncclRecv, log2(P) exchanges per sweep.  The matrix values are
// all-gathered per factorisation, the right-hand side per application; the solution is combined by one all-reduce
// (every unknown has exactly one non-zero contribution).
// The matrix is Jacobi-scaled symmetrically while it is assembled into the fronts (unit-magnitude diagonal, so the
// pivot thresholds are scale-free); ufe_nd_solver_solve applies steps of iterative refinement with the unscaled CSR.
#include "ufe_internal.cuh"
#include "ufe_nd.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <atomic>
#include <numeric>

#define MFB 32    // pivot block width
// dynamic shared memory of k_mf_update: two pivot-block buffers of [T][33] + [32][T] doubles
#define MF_UPD_SMEM(TR, TC) (2 * ((TR) * (MFB + 1) + MFB * (TC)) * sizeof(double))

struct MfStep { int n_active, max_trail; };
struct MfLevel {
  int n = 0, first = 0, steps = 0, maxG = 0, n_internal = 0, max_nb = 0;
  int pmax = 0;                  // largest pivot count of the level
  bool late_schur = false;       // many pivots: the steps leave F22 alone, one K = p pass at the end (k_mf_update modes 1, 2)
  std::vector<MfStep> step;
};
// one NCCL exchange between a child-1 front (owned by `peer_child`) and its parent (owned by `peer_parent`)
struct MfLink {
  int level;                 // level of the parent
  int child_rank, parent_rank;
  int child_front, parent_front;    // local front ids (valid on the owning rank only, else -1)
  int nb, Gp;                       // child's boundary size, parent's front size
  long long f_buf;                  // child side: packed send buffer in F; parent side: receive buffer in F (ld = nb)
  int w_buf;                        // parent side: receive buffer in W (nb doubles); child side: buffer for the parent's w (Gp doubles)
  long long child_schur = -1;       // child side: offset of the Schur block in F, its row stride, offset of the boundary vector in W
  int child_ld = 0, child_wb = -1;
  int parent_w = -1;                // parent side: offset of the parent's vector in W
};

struct MfGraph { const void *a = nullptr, *b = nullptr; int k0 = 0, k1 = 0; cudaGraphExec_t exec = nullptr; int nodes = 0; };

struct ufe_nd_solver {
  int nT = 0, N = 0, nnz = 0, n_fronts = 0;       // n_fronts: fronts owned by this rank
  int rank = 0, nranks = 1;
  ncclComm_t nccl = nullptr;
  std::vector<MfLevel> lev;
  std::vector<long long> lev_lo, lev_hi;      // range of F that holds the fronts of a level (level-major storage)
  std::vector<MfLink> links;
  // per local front (level-major order, p descending inside a level)
  int *ns = nullptr, *p = nullptr, *nb = nullptr, *G = nullptr, *ld = nullptr, *woff = nullptr, *dioff = nullptr,
      *sep_off = nullptr, *up_off = nullptr, *pwoff = nullptr;
  long long *foff = nullptr;
  // children as the extend-add / forward sweep see them: offset of the child's Schur block in F and its row stride,
  // offset of the child's boundary vector in W; -1 = no child
  long long *c_f[2] = {nullptr, nullptr};
  int *c_ld[2] = {nullptr, nullptr}, *c_w[2] = {nullptr, nullptr};
  int *pinv[2] = {nullptr, nullptr};         // per front row (indexed like W): row of child 0 / 1's boundary, or -1
  int *sepdof = nullptr, *upmap = nullptr;   // global unknown of every sep row; parent row of every bnd row
  long long *dst = nullptr;                  // per scalar CSR entry: destination in F, -1 = another rank's front
  int *ptr = nullptr, *ind = nullptr;        // 1-based scalar CSR of the WHOLE matrix (ptr offsets 1-based)
  bool own_pattern = false;                  // ptr / ind allocated here (else they alias the caller's device arrays)
  double *val = nullptr;                     // whole-matrix values (standalone solver / several ranks), else nullptr
  double *scale = nullptr, *dself = nullptr, *dpair = nullptr;
  double *F = nullptr, *W = nullptr, *Dinv = nullptr;
  double *b = nullptr, *x = nullptr, *r = nullptr;     // full-length work vectors
  size_t f_doubles = 0, w_doubles = 0, di_blocks = 0;
  std::vector<size_t> val_counts, row_counts;          // several ranks: per-rank nnz and row counts (strips)
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float factor_ms = 0.f, solve_ms = 0.f;
  double flops = 0.0;
  bool factored = false;
  int premul_pair = 1;                       // preconditioner mode: the right-hand side is multiplied by the 2x2 (1) or 1x1 (0) diagonal blocks
  int use_graphs = 1, k64 = 1, cl_max_fronts = 32, cl_min_g = 384, upd_big = 1, upd_nq = 0, schur_min_p = 256, upd_mma = 1, upd_mma_min_mode = 2, big_min_ctas = 120, lazy_zero = -1;     // lazy_zero: fronts zeroed level by level (-1: when they exceed 1 GB)      // levels with at most this many (large) fronts use the cluster sweeps
  MfGraph g_factor, g_apply[6];
};

// ------------------------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------------------------
// also keeps row i of the 2x2 (u,v) diagonal block of its triangle: dself = a_ii, dpair = a_i,i^1
__global__ void k_mf_scale(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                           double *__restrict__ scale, double *__restrict__ dself, double *__restrict__ dpair) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double d = 0.0, o = 0.0;
  for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) { const int j = ind[k] - 1; if (j == i) d += val[k]; else if (j == (i ^ 1)) o += val[k]; }
  dself[i] = d; dpair[i] = o;
  d = fabs(d);
  scale[i] = d > 0.0 ? 1.0 / sqrt(d) : 1.0;
}

// one thread per matrix row: every (row, column) has its own destination, so plain adds are race-free
// (only the destinations in [lo, hi): the fronts of one level, when the fronts are zeroed level by level)
__global__ void k_mf_assemble(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                              const double *__restrict__ scale, const long long *__restrict__ dst, double *__restrict__ F,
                              long long lo, long long hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double si = scale[i];
  for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) {
    const long long d = dst[k];
    if (d >= lo && d < hi) F[d] += si * val[k] * scale[ind[k] - 1];
  }
}

// identity on the padded pivot rows [ns, p) of every front
__global__ void k_mf_pad(int first, int nf, const int *__restrict__ ns, const int *__restrict__ p, const int *__restrict__ ld,
                         const long long *__restrict__ foff, double *__restrict__ F) {
  int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= nf) return;
  f += first;
  const int r = ns[f] + lane;
  if (r < p[f]) F[foff[f] + (size_t)r * ld[f] + r] = 1.0;
}

// parent front += Schur complements of its children (pull through the inverse maps; child 0 first, then child 1)
// STORE: the parent has NOT been zeroed -- every entry of the front is written (sum of the children's contributions, or 0),
// so a level costs one write pass instead of a memset and a read-modify-write; the matrix entries are added afterwards
template <bool STORE>
__global__ void __launch_bounds__(256)
k_mf_extend(int first, const int *__restrict__ G_, const int *__restrict__ ld_, const long long *__restrict__ foff,
            const int *__restrict__ woff, const long long *__restrict__ cf0, const long long *__restrict__ cf1,
            const int *__restrict__ cld0, const int *__restrict__ cld1, const int *__restrict__ pinv0,
            const int *__restrict__ pinv1, double *__restrict__ F) {
  const int f = first + blockIdx.z;
  const long long o0 = cf0[f], o1 = cf1[f];
  if (!STORE && o0 < 0 && o1 < 0) return;
  const int G = G_[f], ld = ld_[f];
  const int i0 = blockIdx.y * 32, j = blockIdx.x * 32 + threadIdx.x;
  if (i0 >= G || j >= G) return;
  const int *q0 = pinv0 + woff[f], *q1 = pinv1 + woff[f];
  const int b0 = o0 >= 0 ? q0[j] : -1, b1 = o1 >= 0 ? q1[j] : -1;
  if (!STORE && b0 < 0 && b1 < 0) return;
  const int l0 = cld0[f], l1 = cld1[f];
  double *A = F + foff[f];
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int i = i0 + ii;
    if (i >= G) break;
    double v = 0.0;
    bool any = false;
    if (b0 >= 0) { const int a0 = q0[i]; if (a0 >= 0) { v += F[o0 + (size_t)a0 * l0 + b0]; any = true; } }
    if (b1 >= 0) { const int a1 = q1[i]; if (a1 >= 0) { v += F[o1 + (size_t)a1 * l1 + b1]; any = true; } }
    if (STORE) A[(size_t)i * ld + j] = v;
    else if (any) A[(size_t)i * ld + j] += v;
  }
}

// Schur block of a front -> contiguous nb x nb buffer (row stride nb) for the transfer to the parent's rank
__global__ void k_mf_pack(int nb, int ld, const double *__restrict__ src, double *__restrict__ dst) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  if (j >= nb) return;
  for (int i = blockIdx.y * 32 + threadIdx.y; i < min(nb, (int)blockIdx.y * 32 + 32); i += 8) dst[(size_t)i * nb + j] = src[(size_t)i * ld + j];
}

// ------------------------------------------------------------------------------------------------------------------
// factorisation kernels
// ------------------------------------------------------------------------------------------------------------------
// 32 x 32 inverse by the 8 warps of a CTA, matrix in registers: warp w holds the columns 4 w .. 4 w + 3, lane = row.
// In-place Gauss-Jordan -- at step p column p of the work matrix is replaced by the column of the inverse that belongs to
// the pivot row, so there is no [S | I] pair.  The warp that owns the columns 4 g .. 4 g + 3 runs those four steps on its
// own: per step it picks the pivot row P, forms the multipliers m_i = a_ip / a_Pp, updates its four columns and leaves
// (m, P) in shared memory (double-buffered by the parity of g: ONE block barrier per four steps); every other warp then
// applies the four steps to its columns, broadcasting its entries of the pivot rows with shuffles.
// Threshold partial pivoting as an implicit row permutation (perm_s[k] = the row that eliminates column k), searched
// only when the natural pivot is small (the fronts are Jacobi-scaled); vanishing pivots are perturbed statically -- the
// refinement / Krylov iteration around the solver absorbs that.  Rows are scaled lazily: a pivot row keeps its
// unscaled values and its factor 1 / pivot (row scalings commute with the later row operations, whose multipliers are
// formed from the same stored values), one multiply per entry at the end.
// Result: Ip = inverse, Ip[pinv[i]][perm[q]] = a_i[q] s_i.  All MF_PANEL_THREADS threads must call this.
#define MF_PANEL_THREADS 256
struct MfInvScratch { double m[2][4][MFB]; double s[MFB]; int perm[MFB], pinv[MFB], P[2][4]; };

__device__ __forceinline__ void mf_inv32_cta(double (&a)[4], int lane, int w, double (*Ip)[MFB + 1], MfInvScratch &sc) {
  const unsigned FULL = 0xffffffffu;
  if (w == 0) { sc.perm[lane] = lane; sc.s[lane] = 1.0; }
  __syncthreads();
#pragma unroll 1
  for (int g = 0; g < MFB / 4; g++) {
    const int par = g & 1;
    if (w == g) {
      // the owner of the columns 4 g .. 4 g + 3 runs its four steps back to back on its own registers (the block barrier
      // is paid once per four steps) and leaves the multipliers and pivot rows for the other warps
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int p = 4 * g + t;
        const double c = a[t];
        int P = sc.perm[p];
        double piv = __shfl_sync(FULL, c, P);
        if (fabs(piv) < 0.05) {                       // uniform: the largest candidate among the unused rows
          const int mine = sc.perm[lane];
          double v = __shfl_sync(FULL, c, mine);      // lane k: pivot-column entry of row perm[k]
          v = lane >= p ? fabs(v) : -1.0;
          int k = lane;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double v2 = __shfl_xor_sync(FULL, v, o);
            const int k2 = __shfl_xor_sync(FULL, k, o);
            if (v2 > v || (v2 == v && k2 < k)) { v = v2; k = k2; }
          }
          const int pk = __shfl_sync(FULL, mine, k);
          __syncwarp();
          if (lane == 0) { sc.perm[k] = P; sc.perm[p] = pk; }
          __syncwarp();
          P = pk;
          piv = __shfl_sync(FULL, c, P);
        }
        if (fabs(piv) < 1e-13) piv = piv < 0.0 ? -1e-13 : 1e-13;
        const double d = 1.0 / piv;
        const bool isP = lane == P;
        const double m = isP ? 0.0 : c * d;
        sc.m[par][t][lane] = m;
        if (isP) sc.s[lane] = d;
        if (lane == 0) sc.P[par][t] = P;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const double x = __shfl_sync(FULL, a[k], P);
          a[k] = fma(-m, x, a[k]);
        }
        a[t] = isP ? 1.0 : -m;                        // the inverse's column takes the place of column p
      }
    }
    __syncthreads();                                  // one barrier for the whole block, outside the divergent branches
    if (w != g) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const double m = sc.m[par][t][lane];
        const int P = sc.P[par][t];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const double x = __shfl_sync(FULL, a[k], P);
          a[k] = fma(-m, x, a[k]);
        }
      }
    }
  }
  if (w == 0) sc.pinv[sc.perm[lane]] = lane;
  __syncthreads();
  const int row = sc.pinv[lane];
  const double s = sc.s[lane];
#pragma unroll
  for (int k = 0; k < 4; k++) Ip[row][sc.perm[4 * w + k]] = a[k] * s;
  __syncthreads();
}

// pivot step b of the fronts [first, first + gridDim.y): D_b^-1 (warp 0 of every CTA of a front inverts the 32 x 32 block
// itself while the other warps already fetch their first rows; the block is not written here, the inverse goes to the
// side buffer) and the L panel rows of this CTA:  A_ib <- A_ib D_b^-1, two rows in flight per warp.
__global__ void __launch_bounds__(MF_PANEL_THREADS)
k_mf_panel(int first, int b, int rows_per_cta, const int *__restrict__ G_, const int *__restrict__ ld_,
           const long long *__restrict__ foff, const int *__restrict__ dioff, double *__restrict__ F, double *__restrict__ Dinv) {
  constexpr int NW = MF_PANEL_THREADS / 32;
  const int f = first + blockIdx.y;
  const int G = G_[f], ld = ld_[f];
  const int r0 = (b + 1) * MFB;
  const int rbeg = r0 + blockIdx.x * rows_per_cta;
  if (blockIdx.x > 0 && rbeg >= G) return;
  __shared__ double Ip[MFB][MFB + 1], Xw[NW][2][MFB];
  __shared__ MfInvScratch isc;
  const int j = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *A = F + foff[f];
  const int rend = min(G, rbeg + rows_per_cta);
  int r = rbeg + w;
  double x0 = 0.0, x1 = 0.0;
  if (r < rend) x0 = A[(size_t)r * ld + b * MFB + j];
  if (r + NW < rend) x1 = A[(size_t)(r + NW) * ld + b * MFB + j];
  {
    double a[4];
    const double2 *row = reinterpret_cast<const double2 *>(A + (size_t)(b * MFB + j) * ld + b * MFB + 4 * w);
    const double2 v0 = row[0], v1 = row[1];
    a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
    mf_inv32_cta(a, j, w, Ip, isc);
  }
  if (blockIdx.x == 0)
    for (int q = threadIdx.x; q < MFB * MFB; q += MF_PANEL_THREADS) Dinv[((size_t)dioff[f] + b) * (MFB * MFB) + q] = Ip[q >> 5][q & 31];
  for (bool firstpass = true; r < rend; r += 2 * NW, firstpass = false) {
    double *T0 = A + (size_t)r * ld + b * MFB, *T1 = T0 + (size_t)NW * ld;
    const bool two = r + NW < rend;
    if (!firstpass) { x0 = T0[j]; x1 = two ? T1[j] : 0.0; }
    Xw[w][0][j] = x0; Xw[w][1][j] = x1;
    __syncwarp();
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
    for (int k = 0; k < MFB; k++) { const double ip = Ip[k][j]; s0 += Xw[w][0][k] * ip; s1 += Xw[w][1][k] * ip; }
    T0[j] = s0;
    if (two) T1[j] = s1;
    __syncwarp();
  }
}

// second pivot block b1 = b0 + 1 of a 64-wide outer step (the trailing matrix has NOT been updated with block b0 yet):
//   column part (blockIdx.x < nx_col):  D1 = A_11 - L_10 U_01,  D1^-1 -> side buffer,
//                                       L_i1 = (A_i1 - L_i0 U_01) D1^-1   for the rows i > b1 of this CTA,
//   row part (blockIdx.x >= nx_col):    U_1j = A_1j - L_10 U_0j            for the columns j > b1 of this CTA.
// The two parts read and write disjoint blocks, so they share one launch.  Warps work independently after the set-up.
__global__ void __launch_bounds__(MF_PANEL_THREADS)
k_mf_panel2(int first, int b1, int nx_col, int chunk, const int *__restrict__ G_, const int *__restrict__ ld_,
            const long long *__restrict__ foff, const int *__restrict__ dioff, double *__restrict__ F, double *__restrict__ Dinv) {
  constexpr int NW = MF_PANEL_THREADS / 32;
  const int f = first + blockIdx.y;
  const int G = G_[f], ld = ld_[f];
  const int b0 = b1 - 1, r1 = (b1 + 1) * MFB;
  const bool colpart = (int)blockIdx.x < nx_col;
  const int cbeg = r1 + (colpart ? (int)blockIdx.x : (int)blockIdx.x - nx_col) * chunk;
  if (cbeg >= G && !(colpart && blockIdx.x == 0)) return;
  __shared__ double W1[MFB][MFB + 1], Ip[MFB][MFB + 1], X[MFB][MFB + 1], Y[MFB][MFB + 1], Xr[NW][2][MFB];
  __shared__ MfInvScratch isc;
  const int j = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *A = F + foff[f];
  const int cend = min(G, cbeg + chunk);
  for (int i = w; i < MFB; i += NW) X[i][j] = A[(size_t)(b1 * MFB + i) * ld + b0 * MFB + j];            // L_10
  if (colpart) {
    for (int i = w; i < MFB; i += NW) Y[i][j] = A[(size_t)(b0 * MFB + i) * ld + b1 * MFB + j];          // U_01
    __syncthreads();
    for (int i = w; i < MFB; i += NW) {
      double d = A[(size_t)(b1 * MFB + i) * ld + b1 * MFB + j];
#pragma unroll 8
      for (int k = 0; k < MFB; k++) d -= X[i][k] * Y[k][j];
      W1[i][j] = d;
    }
    __syncthreads();
    {
      double a[4];
#pragma unroll
      for (int q = 0; q < 4; q++) a[q] = W1[j][4 * w + q];
      mf_inv32_cta(a, j, w, Ip, isc);
    }
    if (blockIdx.x == 0)
      for (int q = threadIdx.x; q < MFB * MFB; q += MF_PANEL_THREADS) Dinv[((size_t)dioff[f] + b1) * (MFB * MFB) + q] = Ip[q >> 5][q & 31];
    for (int r = cbeg + w; r < cend; r += 2 * NW) {        // two rows in flight per warp
      double *T0 = A + (size_t)r * ld, *T1 = T0 + (size_t)NW * ld;
      const bool two = r + NW < cend;
      const double l0 = T0[b0 * MFB + j], l1 = two ? T1[b0 * MFB + j] : 0.0;
      double v0 = T0[b1 * MFB + j], v1 = two ? T1[b1 * MFB + j] : 0.0;
      Xr[w][0][j] = l0; Xr[w][1][j] = l1;
      __syncwarp();
#pragma unroll 8
      for (int k = 0; k < MFB; k++) { const double y = Y[k][j]; v0 -= Xr[w][0][k] * y; v1 -= Xr[w][1][k] * y; }
      __syncwarp();
      Xr[w][0][j] = v0; Xr[w][1][j] = v1;
      __syncwarp();
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < MFB; k++) { const double ip = Ip[k][j]; s0 += Xr[w][0][k] * ip; s1 += Xr[w][1][k] * ip; }
      T0[b1 * MFB + j] = s0;
      if (two) T1[b1 * MFB + j] = s1;
      __syncwarp();
    }
  } else {
    __syncthreads();
    // work item = (32-column chunk, group of 8 rows of block b1); lane = column
    const int nch = (cend - cbeg + MFB - 1) / MFB;
    for (int it = w; it < nch * 4; it += NW) {
      const int col = cbeg + (it >> 2) * MFB + j, ig = (it & 3) * 8;
      if (col >= cend) continue;
      double acc[8];
#pragma unroll
      for (int q = 0; q < 8; q++) acc[q] = 0.0;
      const double *U0 = A + (size_t)(b0 * MFB) * ld + col;
#pragma unroll 8
      for (int k = 0; k < MFB; k++) {
        const double u = U0[(size_t)k * ld];
#pragma unroll
        for (int q = 0; q < 8; q++) acc[q] += X[ig + q][k] * u;
      }
      double *U1 = A + (size_t)(b1 * MFB + ig) * ld + col;
#pragma unroll
      for (int q = 0; q < 8; q++) U1[(size_t)q * ld] -= acc[q];
    }
  }
}

// trailing update after the pivot blocks b .. b + nkb - 1 (nkb = 1 or 2):  A_ij -= sum_k L_ik A_kj  for i, j >= 32 (b + nkb).
// Tile (8 TG) x (8 TG), TG x TG threads, 8 x 8 register micro-tile made of 2 x 2 sub-blocks (rows 2 ty + 2 TG p + {0,1},
// columns 2 tx + 2 TG q + {0,1}): the L values are broadcast loads, the U values and all accesses to A are 16-byte
// accesses of adjacent columns.  With nkb = 2 the tile of A is read and written once per 64 pivots.
// Both operand tiles of both pivot blocks are fetched with cp.async up front (one buffer per pivot block, all loads in
// flight at once, no staging registers), the lines of the A tile are prefetched into L2 meanwhile, and the read-modify-
// write of the A tile is issued in batches of eight 16-byte loads -- with one CTA of 8 warps per SM (208 registers per
// thread) there is nothing else to hide the memory latency behind.
__device__ __forceinline__ void mf_cp_async8(double *dst, const double *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void mf_cp_async16(double *dst, const double *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n));
}

// mode 0: the update of a step as described above.
// Fronts with many pivots (mode 1 + mode 2): the Schur complement F22 (rows and columns >= p) is NOT touched by the
// steps -- mode 1 updates only the L-shaped rest of the trailing matrix, i.e. the pivot square and the boundary panels,
// entries of F22 inside a straddling tile are masked -- and is updated ONCE at the end, mode 2:  F22 -= L21 U12  with
// K = p, the pivot blocks streamed through the same two buffers in pairs.  One pass over F22 instead of p / 64, and the
// prologue / epilogue of a tile are amortised over K = p instead of 64.
template <int TG, int NQ, bool LATE>      // LATE = false: mode 0 only (the compiler drops the masks and the K loop)
__global__ void __launch_bounds__(TG * TG, (TG == 16 ? 1 : 2) * (NQ == 2 ? 2 : 1))
k_mf_update(int first, int b, int nkb, int mode_rt, const int *__restrict__ p_, const int *__restrict__ G_, const int *__restrict__ ld_,
            const long long *__restrict__ foff, double *__restrict__ F) {
  // tile TR x TC, TG x TG threads, 8 x (2 NQ) per thread (NQ = 4: 8 x 8, 208 registers; NQ = 2: 8 x 4, twice the warps per SM)
  constexpr int TR = 8 * TG, TC = 2 * NQ * TG, NT = TG * TG, BUFA = TR * (MFB + 1) + MFB * TC;    // doubles per pivot-block buffer (even)
  const int mode = LATE ? mode_rt : 0;
  const int f = first + blockIdx.z;
  const int G = G_[f], ld = ld_[f];
  const int p = mode ? p_[f] : 0;
  const int kb_begin = mode == 2 ? 0 : b, kb_end = mode == 2 ? p / MFB : b + nkb;
  const int r0 = mode == 2 ? p : (b + nkb) * MFB;
  const int i0 = r0 + blockIdx.y * TR, j0 = r0 + blockIdx.x * TC;
  if (i0 >= G || j0 >= G) return;
  if (mode == 1 && i0 >= p && j0 >= p) return;       // a tile of F22: left to the final pass
  double *A = F + foff[f];
  extern __shared__ __align__(16) double mf_sm[];
  const int t = threadIdx.x;
  const int tx = t % TG, ty = t / TG;
  // both operand tiles of a pair of pivot blocks are fetched with cp.async up front (one buffer per pivot block)
  auto fetch = [&](int kb0, int nk) {
    for (int kk = 0; kk < nk; kk++) {
      const int kb = kb0 + kk;
      double *sLb = mf_sm + kk * BUFA;                 // [TR][33]
      double *sUb = sLb + TR * (MFB + 1);              // [32][TC], 16-byte aligned: TR * 33 is even
#pragma unroll 4
      for (int q = t; q < TR * MFB; q += NT) {
        const int row = q >> 5, k = q & 31;
        const bool ok = i0 + row < G;
        mf_cp_async8(sLb + row * (MFB + 1) + k, ok ? A + (size_t)(i0 + row) * ld + kb * MFB + k : A, ok);
      }
#pragma unroll 4
      for (int q = t; q < MFB * (TC / 2); q += NT) {
        const int k = q / (TC / 2), c = 2 * (q % (TC / 2));
        const bool ok = j0 + c < G;
        mf_cp_async16(sUb + k * TC + c, ok ? A + (size_t)(kb * MFB + k) * ld + j0 + c : A, ok);
      }
      asm volatile("cp.async.commit_group;");
    }
  };
  fetch(kb_begin, min(2, kb_end - kb_begin));
  // L2 prefetch of this thread's share of the A tile (128-byte lines; TR rows x TC / 16 lines)
  for (int q = t; q < TR * (TC / 16); q += NT) {
    const int row = q / (TC / 16), c = 16 * (q % (TC / 16));
    if (i0 + row < G && j0 + c < G) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)(i0 + row) * ld + j0 + c));
  }
  double acc[8][2 * NQ];
#pragma unroll
  for (int pp = 0; pp < 8; pp++)
#pragma unroll
    for (int q = 0; q < 2 * NQ; q++) acc[pp][q] = 0.0;
  for (int kb0 = kb_begin; kb0 < kb_end; kb0 += 2) {
    const int nk = min(2, kb_end - kb0);
    if (kb0 > kb_begin) { __syncthreads(); fetch(kb0, nk); }     // the buffers are free again
    for (int kk = 0; kk < nk; kk++) {
      if (kk == 0 && nk == 2) asm volatile("cp.async.wait_group 1;"); else asm volatile("cp.async.wait_group 0;");
      __syncthreads();
      const double *sLb = mf_sm + kk * BUFA;
      const double *sUb = sLb + TR * (MFB + 1);
#pragma unroll 2
      for (int k = 0; k < MFB; k++) {
        double a[8], u[2 * NQ];
#pragma unroll
        for (int pp = 0; pp < 4; pp++) { a[2 * pp] = sLb[(2 * ty + 2 * TG * pp) * (MFB + 1) + k]; a[2 * pp + 1] = sLb[(2 * ty + 2 * TG * pp + 1) * (MFB + 1) + k]; }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const double2 v = *reinterpret_cast<const double2 *>(sUb + k * TC + 2 * tx + 2 * TG * q);
          u[2 * q] = v.x; u[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int pp = 0; pp < 8; pp++)
#pragma unroll
          for (int q = 0; q < 2 * NQ; q++) acc[pp][q] += a[pp] * u[q];
      }
    }
  }
#pragma unroll
  for (int pp = 0; pp < 4; pp++) {                   // two rows (2 NQ x 16 bytes) per batch
    double2 v[2][NQ];
    bool on[2][NQ];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = i0 + 2 * ty + 2 * TG * pp + h;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int j = j0 + 2 * tx + 2 * TG * q;
        on[h][q] = i < G && j < G && !(mode == 1 && i >= p && j >= p);
        v[h][q] = on[h][q] ? *reinterpret_cast<const double2 *>(A + (size_t)i * ld + j) : make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = i0 + 2 * ty + 2 * TG * pp + h;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int j = j0 + 2 * tx + 2 * TG * q;
        if (on[h][q]) {
          double2 w = v[h][q];
          w.x -= acc[2 * pp + h][2 * q]; w.y -= acc[2 * pp + h][2 * q + 1];
          *reinterpret_cast<double2 *>(A + (size_t)i * ld + j) = w;
        }
      }
    }
  }
}

// The same update on the fp64 TENSOR cores for large trailing matrices: mma.sync.m8n8k4.f64 (DMMA; tcgen05 has no fp64
// path, so this is the tensor-core instruction for this precision -- measured 37.0 TFLOP/s on the B200 against 33.7 for
// DFMA, tools/micro/).  One DMMA is 256 FMA per warp instruction: 8 x fewer issue slots and 6 x fewer shared-memory
// operand loads per flop than the 8 x 4 SIMT micro-tile, whose limiter is the shared-memory pipe (ncu: "Shared is the
// highest-utilized pipeline").  128 x 128 tile, 8 warps as 2 x 4, warp tile 64 x 32 = 8 x 4 m8n8 tiles (128 accumulator
// registers); operand strides 36 / 132 doubles make the fragment loads of a half-warp hit 16 distinct bank pairs.
#define MMA_LS 36
#define MMA_US 132
#define MF_MMA_SMEM (2 * (128 * MMA_LS + MFB * MMA_US) * sizeof(double))
__device__ __forceinline__ void mf_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
k_mf_update_mma(int first, int b, int nkb, int mode, const int *__restrict__ p_, const int *__restrict__ G_, const int *__restrict__ ld_,
                const long long *__restrict__ foff, double *__restrict__ F) {
  constexpr int T = 128, NT = 256, BUFA = T * MMA_LS + MFB * MMA_US;
  const int f = first + blockIdx.z;
  const int G = G_[f], ld = ld_[f];
  const int p = mode ? p_[f] : 0;
  const int kb_begin = mode == 2 ? 0 : b, kb_end = mode == 2 ? p / MFB : b + nkb;
  const int r0 = mode == 2 ? p : (b + nkb) * MFB;
  const int i0 = r0 + blockIdx.y * T, j0 = r0 + blockIdx.x * T;
  if (i0 >= G || j0 >= G) return;
  if (mode == 1 && i0 >= p && j0 >= p) return;       // a tile of F22: left to the final pass
  double *A = F + foff[f];
  extern __shared__ __align__(16) double mf_sm[];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int wrow = 64 * (w >> 2), wcol = 32 * (w & 3), g = lane >> 2, tg = lane & 3;
  auto fetch = [&](int kb0, int nk) {
    for (int kk = 0; kk < nk; kk++) {
      const int kb = kb0 + kk;
      double *sLb = mf_sm + kk * BUFA;                 // [128][MMA_LS]
      double *sUb = sLb + T * MMA_LS;                  // [32][MMA_US]
#pragma unroll 4
      for (int q = t; q < T * (MFB / 2); q += NT) {
        const int row = q / (MFB / 2), c = 2 * (q % (MFB / 2));
        const bool ok = i0 + row < G;
        mf_cp_async16(sLb + row * MMA_LS + c, ok ? A + (size_t)(i0 + row) * ld + kb * MFB + c : A, ok);
      }
#pragma unroll 4
      for (int q = t; q < MFB * (T / 2); q += NT) {
        const int k = q / (T / 2), c = 2 * (q % (T / 2));
        const bool ok = j0 + c < G;
        mf_cp_async16(sUb + k * MMA_US + c, ok ? A + (size_t)(kb * MFB + k) * ld + j0 + c : A, ok);
      }
      asm volatile("cp.async.commit_group;");
    }
  };
  fetch(kb_begin, min(2, kb_end - kb_begin));
  for (int q = t; q < T * (T / 16); q += NT) {        // L2 prefetch of the A tile
    const int row = q / (T / 16), c = 16 * (q % (T / 16));
    if (i0 + row < G && j0 + c < G) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (size_t)(i0 + row) * ld + j0 + c));
  }
  double acc[8][4][2];
#pragma unroll
  for (int mt = 0; mt < 8; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
  for (int kb0 = kb_begin; kb0 < kb_end; kb0 += 2) {
    const int nk = min(2, kb_end - kb0);
    if (kb0 > kb_begin) { __syncthreads(); fetch(kb0, nk); }     // the buffers are free again
    for (int kk = 0; kk < nk; kk++) {
      if (kk == 0 && nk == 2) asm volatile("cp.async.wait_group 1;"); else asm volatile("cp.async.wait_group 0;");
      __syncthreads();
      const double *sLb = mf_sm + kk * BUFA + (wrow + g) * MMA_LS + tg;
      const double *sUb = mf_sm + kk * BUFA + T * MMA_LS + tg * MMA_US + wcol + g;
#pragma unroll 2
      for (int k0 = 0; k0 < MFB; k0 += 4) {
        double a[8], u[4];
#pragma unroll
        for (int mt = 0; mt < 8; mt++) a[mt] = sLb[mt * 8 * MMA_LS + k0];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) u[nt] = sUb[k0 * MMA_US + 8 * nt];
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) mf_dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], u[nt]);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 8; mt++) {
    const int i = i0 + wrow + 8 * mt + g;
    double2 v[4];
    bool on[4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const int j = j0 + wcol + 8 * nt + 2 * tg;
      on[nt] = i < G && j < G && !(mode == 1 && i >= p && j >= p);
      v[nt] = on[nt] ? *reinterpret_cast<const double2 *>(A + (size_t)i * ld + j) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      if (!on[nt]) continue;
      const int j = j0 + wcol + 8 * nt + 2 * tg;
      *reinterpret_cast<double2 *>(A + (size_t)i * ld + j) = make_double2(v[nt].x - acc[mt][nt][0], v[nt].y - acc[mt][nt][1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// solve kernels: one CTA (1024 threads = 32 warps) per front, one launch per level and direction; the front's vector
// lives in shared memory during the sweep, the factors are streamed row-wise (row-major fronts: every dot product is
// a coalesced read of one row segment).
// ------------------------------------------------------------------------------------------------------------------
// dot products of NR rows (row r at Ar + r * rstride, rows >= nrows skipped) with sw over the columns [j0, j1): all
// NR x 4 loads of a 128-column slab are issued before the first multiply (the sweeps are latency-bound)
template <int NR>
__device__ __forceinline__ void mf_rowdots(const double *__restrict__ Ar, size_t rstride, int nrows, const double *sw, int j0, int j1,
                                           int lane, double (&out)[NR]) {
  double acc[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) acc[r] = 0.0;
  for (int j = j0 + lane; j < j1; j += 128) {
    double a[NR][4], x[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int jj = j + 32 * q;
      const bool ok = jj < j1;
      x[q] = ok ? sw[jj] : 0.0;
#pragma unroll
      for (int r = 0; r < NR; r++) a[r][q] = (ok && r < nrows) ? Ar[(size_t)r * rstride + jj] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < NR; r++) acc[r] += (a[r][0] * x[0] + a[r][1] * x[1]) + (a[r][2] * x[2] + a[r][3] * x[3]);
  }
#pragma unroll
  for (int r = 0; r < NR; r++) {
    double v = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    out[r] = v;
  }
}
__device__ __forceinline__ double mf_rowdot(const double *__restrict__ Ar, const double *sw, int j0, int j1, int lane) {
  double o[1];
  mf_rowdots<1>(Ar, 0, 1, sw, j0, j1, lane, o);
  return o[0];
}

// forward:  w = [scaled rhs of the own unknowns ; 0] + the boundary vectors of the children, then  y = L^-1 w  block by
// block (left-looking: block b needs the rows of block b only), then the boundary rows  w_b -= L21 y, which are this
// front's contribution to its parent.  dself != nullptr: the right-hand side is first multiplied by the 2x2 diagonal
// blocks of A (the solver then applies (B A)^-1, B = block-Jacobi scaling).
__global__ void __launch_bounds__(1024)
k_mf_fwd(int first, const int *__restrict__ ns_, const int *__restrict__ p_, const int *__restrict__ G_,
         const int *__restrict__ ld_, const long long *__restrict__ foff, const int *__restrict__ woff_,
         const int *__restrict__ cw0, const int *__restrict__ cw1, const int *__restrict__ pinv0, const int *__restrict__ pinv1,
         const int *__restrict__ sep_off, const int *__restrict__ sepdof, const double *__restrict__ scale,
         const double *__restrict__ rhs, const double *__restrict__ dself, const double *__restrict__ dpair,
         const double *__restrict__ F, double *__restrict__ W) {
  extern __shared__ double mf_sw[];
  const int f = first + blockIdx.x;
  const int ns = ns_[f], p = p_[f], G = G_[f], ld = ld_[f], wo = woff_[f];
  const double *A = F + foff[f];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int w0 = cw0[f], w1 = cw1[f];
  for (int i = t; i < G; i += 1024) {
    double v = 0.0;
    if (i < ns) {
      const int d = sepdof[sep_off[f] + i];
      v = scale[d] * (dself ? dself[d] * rhs[d] + (dpair ? dpair[d] * rhs[d ^ 1] : 0.0) : rhs[d]);
    }
    if (w0 >= 0) { const int a = pinv0[wo + i]; if (a >= 0) v += W[w0 + a]; }
    if (w1 >= 0) { const int a = pinv1[wo + i]; if (a >= 0) v += W[w1 + a]; }
    mf_sw[i] = v;
  }
  __syncthreads();
  const int nblk = p / MFB;
  for (int b = 1; b < nblk; b++) {
    const int row = b * MFB + warp;
    const double v = mf_rowdot(A + (size_t)row * ld, mf_sw, 0, b * MFB, lane);
    if (lane == 0) mf_sw[row] -= v;
    __syncthreads();
  }
  for (int r = p + warp; r < G; r += 128) {           // rows r, r + 32, r + 64, r + 96 together
    double v[4];
    const int nr = (G - r + 31) / 32;
    mf_rowdots<4>(A + (size_t)r * ld, (size_t)32 * ld, nr, mf_sw, 0, p, lane, v);
    if (lane < 4 && lane < nr) mf_sw[r + 32 * lane] -= v[lane];
  }
  __syncthreads();
  double *w = W + wo;
  for (int i = t; i < G; i += 1024) w[i] = mf_sw[i];
}

// backward:  x_b from the parent's vector,  x_s = U11^-1 (y - U12 x_b)  block by block (D_b^-1 from the side buffer);
// [x_s; x_b] stays in w (for the children) and x_s goes, unscaled, to the global solution vector
__global__ void __launch_bounds__(1024)
k_mf_bwd(int first, const int *__restrict__ ns_, const int *__restrict__ p_, const int *__restrict__ nb_,
         const int *__restrict__ ld_, const long long *__restrict__ foff, const int *__restrict__ woff_,
         const int *__restrict__ pwoff, const int *__restrict__ up_off, const int *__restrict__ upmap,
         const int *__restrict__ dioff, const int *__restrict__ sep_off, const int *__restrict__ sepdof,
         const double *__restrict__ scale, const double *__restrict__ F, const double *__restrict__ Dinv,
         double *__restrict__ W, double *__restrict__ x, int accumulate) {
  extern __shared__ double mf_sw[];
  __shared__ double ts[MFB];
  const int f = first + blockIdx.x;
  const int ns = ns_[f], p = p_[f], nb = nb_[f], ld = ld_[f], G = p + nb;
  const double *A = F + foff[f];
  double *w = W + woff_[f];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < p; i += 1024) mf_sw[i] = w[i];
  if (nb > 0) {
    const double *pw = W + pwoff[f];
    const int *up = upmap + up_off[f];
    for (int r = t; r < nb; r += 1024) mf_sw[p + r] = pw[up[r]];
  }
  __syncthreads();
  const double *Di = Dinv + (size_t)dioff[f] * (MFB * MFB) + warp * MFB + lane;
  for (int b = p / MFB - 1; b >= 0; b--) {
    const int row = b * MFB + warp;
    const double dv = Di[(size_t)b * (MFB * MFB)];
    const double v = mf_rowdot(A + (size_t)row * ld, mf_sw, (b + 1) * MFB, G, lane);
    if (lane == 0) ts[warp] = mf_sw[row] - v;
    __syncthreads();
    double xv = dv * ts[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xv += __shfl_xor_sync(0xffffffffu, xv, o);
    if (lane == 0) mf_sw[row] = xv;
    __syncthreads();
  }
  for (int i = t; i < G; i += 1024) w[i] = mf_sw[i];
  for (int i = t; i < ns; i += 1024) {
    const int d = sepdof[sep_off[f] + i];
    const double xs = scale[d] * mf_sw[i];
    if (accumulate) x[d] += xs; else x[d] = xs;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The same sweeps for the few LARGE fronts near the root, where one SM cannot stream a front fast enough: a thread-block
// cluster of MFCL CTAs per front.  Every CTA keeps the whole vector in its shared memory; a block step's 32 dot products
// are split by columns over the CTAs, the partial sums are pushed to every CTA's shared memory (DSMEM), one cluster
// barrier later every CTA adds them in rank order (bit-identical everywhere) and continues.  The boundary rows (forward)
// are split by rows and need no exchange.
// ------------------------------------------------------------------------------------------------------------------
#define MFCL 8
namespace cg = cooperative_groups;

// columns [j0, j1) in 32-aligned pieces: piece c of CL
__device__ __forceinline__ void mf_piece(int j0, int j1, int c, int CL, int *a, int *b) {
  const int nblk = (j1 - j0 + MFB - 1) / MFB, per = (nblk + CL - 1) / CL;
  *a = min(j1, j0 + c * per * MFB); *b = min(j1, j0 + (c + 1) * per * MFB);
}

__global__ void __cluster_dims__(MFCL, 1, 1) __launch_bounds__(1024)
k_mf_fwd_cl(int first, const int *__restrict__ ns_, const int *__restrict__ p_, const int *__restrict__ G_,
            const int *__restrict__ ld_, const long long *__restrict__ foff, const int *__restrict__ woff_,
            const int *__restrict__ cw0, const int *__restrict__ cw1, const int *__restrict__ pinv0, const int *__restrict__ pinv1,
            const int *__restrict__ sep_off, const int *__restrict__ sepdof, const double *__restrict__ scale,
            const double *__restrict__ rhs, const double *__restrict__ dself, const double *__restrict__ dpair,
            const double *__restrict__ F, double *__restrict__ W) {
  extern __shared__ double mf_sw[];
  __shared__ double part[2][MFCL][MFB];
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int f = first + blockIdx.x / MFCL;
  const int ns = ns_[f], p = p_[f], G = G_[f], ld = ld_[f], wo = woff_[f];
  const double *A = F + foff[f];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int w0 = cw0[f], w1 = cw1[f];
  for (int i = t; i < G; i += 1024) {
    double v = 0.0;
    if (i < ns) {
      const int d = sepdof[sep_off[f] + i];
      v = scale[d] * (dself ? dself[d] * rhs[d] + (dpair ? dpair[d] * rhs[d ^ 1] : 0.0) : rhs[d]);
    }
    if (w0 >= 0) { const int a = pinv0[wo + i]; if (a >= 0) v += W[w0 + a]; }
    if (w1 >= 0) { const int a = pinv1[wo + i]; if (a >= 0) v += W[w1 + a]; }
    mf_sw[i] = v;
  }
  cluster.sync();
  const int nblk = p / MFB;
  for (int b = 1; b < nblk; b++) {
    const int row = b * MFB + warp, par = b & 1;
    int ja, jb;
    mf_piece(0, b * MFB, c, MFCL, &ja, &jb);
    const double v = mf_rowdot(A + (size_t)row * ld, mf_sw, ja, jb, lane);
    if (lane < MFCL) cluster.map_shared_rank(&part[par][c][warp], lane)[0] = v;
    cluster.sync();
    if (t < MFB) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < MFCL; q++) s += part[par][q][t];
      mf_sw[b * MFB + t] -= s;
    }
    __syncthreads();
  }
  double *w = W + wo;
  for (int r = p + c * 32 + warp; r < G; r += 128 * MFCL) {       // rows r, r + 32 MFCL, ... (4 together)
    double v[4];
    const int nr = (G - r + 32 * MFCL - 1) / (32 * MFCL);
    mf_rowdots<4>(A + (size_t)r * ld, (size_t)32 * MFCL * ld, nr, mf_sw, 0, p, lane, v);
    if (lane < 4 && lane < nr) w[r + 32 * MFCL * lane] = mf_sw[r + 32 * MFCL * lane] - v[lane];
  }
  if (c == 0) for (int i = t; i < p; i += 1024) w[i] = mf_sw[i];
  cluster.sync();            // nobody leaves while its shared memory may still be written by a peer
}

__global__ void __cluster_dims__(MFCL, 1, 1) __launch_bounds__(1024)
k_mf_bwd_cl(int first, const int *__restrict__ ns_, const int *__restrict__ p_, const int *__restrict__ nb_,
            const int *__restrict__ ld_, const long long *__restrict__ foff, const int *__restrict__ woff_,
            const int *__restrict__ pwoff, const int *__restrict__ up_off, const int *__restrict__ upmap,
            const int *__restrict__ dioff, const int *__restrict__ sep_off, const int *__restrict__ sepdof,
            const double *__restrict__ scale, const double *__restrict__ F, const double *__restrict__ Dinv,
            double *__restrict__ W, double *__restrict__ x, int accumulate) {
  extern __shared__ double mf_sw[];
  __shared__ double part[2][MFCL][MFB], ts[MFB];
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int f = first + blockIdx.x / MFCL;
  const int ns = ns_[f], p = p_[f], nb = nb_[f], ld = ld_[f], G = p + nb;
  const double *A = F + foff[f];
  double *w = W + woff_[f];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < p; i += 1024) mf_sw[i] = w[i];
  if (nb > 0) {
    const double *pw = W + pwoff[f];
    const int *up = upmap + up_off[f];
    for (int r = t; r < nb; r += 1024) mf_sw[p + r] = pw[up[r]];
  }
  cluster.sync();
  const double *Di = Dinv + (size_t)dioff[f] * (MFB * MFB) + warp * MFB + lane;
  for (int b = p / MFB - 1; b >= 0; b--) {
    const int row = b * MFB + warp, par = b & 1;
    const double dv = Di[(size_t)b * (MFB * MFB)];
    int ja, jb;
    mf_piece((b + 1) * MFB, G, c, MFCL, &ja, &jb);
    const double v = mf_rowdot(A + (size_t)row * ld, mf_sw, ja, jb, lane);
    if (lane < MFCL) cluster.map_shared_rank(&part[par][c][warp], lane)[0] = v;
    cluster.sync();
    if (t < MFB) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < MFCL; q++) s += part[par][q][t];
      ts[t] = mf_sw[b * MFB + t] - s;
    }
    __syncthreads();
    double xv = dv * ts[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xv += __shfl_xor_sync(0xffffffffu, xv, o);
    if (lane == 0) mf_sw[row] = xv;
    __syncthreads();
  }
  if (c == 0) {
    for (int i = t; i < G; i += 1024) w[i] = mf_sw[i];
    for (int i = t; i < ns; i += 1024) {
      const int d = sepdof[sep_off[f] + i];
      const double xs = scale[d] * mf_sw[i];
      if (accumulate) x[d] += xs; else x[d] = xs;
    }
  }
  cluster.sync();
}

// r = b - A x, unscaled CSR (1-based), one thread per row
__global__ void k_mf_residual(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                              const double *__restrict__ b, const double *__restrict__ x, double *__restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = b[i];
  for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) s -= val[k] * x[ind[k] - 1];
  r[i] = s;
}

__global__ void k_mf_copy_slice(int n, const double *__restrict__ src, double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
template <class T> static int mf_upload(T **d, const std::vector<T> &h) {
  UFE_CUDA(cudaMalloc((void **)d, std::max<size_t>(1, h.size()) * sizeof(T)));
  if (!h.empty()) UFE_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return UFE_OK;
}

static void mf_graph_drop(MfGraph &g) { if (g.exec) cudaGraphExecDestroy(g.exec); g = MfGraph(); }

extern "C" void ufe_nd_solver_free(ufe_nd_solver *S) {
  if (!S) return;
  mf_graph_drop(S->g_factor);
  for (MfGraph &g : S->g_apply) mf_graph_drop(g);
  void *q[] = {S->ns, S->p, S->nb, S->G, S->ld, S->woff, S->dioff, S->sep_off, S->up_off, S->pwoff, S->foff, S->c_f[0], S->c_f[1],
               S->c_ld[0], S->c_ld[1], S->c_w[0], S->c_w[1], S->pinv[0], S->pinv[1], S->sepdof, S->upmap, S->dst, S->val, S->scale,
               S->dself, S->dpair, S->F, S->W, S->Dinv, S->b, S->x, S->r};
  for (void *v : q) if (v) cudaFree(v);
  if (S->own_pattern) { cudaFree(S->ptr); cudaFree(S->ind); }
  if (S->e0) cudaEventDestroy(S->e0);
  if (S->e1) cudaEventDestroy(S->e1);
  if (S->st) cudaStreamDestroy(S->st);
  delete S;
}

// Builds the device layout.  ptr / ind: scalar CSR pattern of the WHOLE matrix A, 1-based (the reference's
// type_sparse_matrix_CSR_dp convention), N = 2 nT rows, on the host; its block pattern must be the one that was analysed.
// d_ptr / d_ind: the same pattern already on the device (aliased, not copied), or nullptr.
static int mf_create(const ufe_nd_tree *T, int N, const int *ptr, const int *ind, const int *d_ptr, const int *d_ind,
                     int rank, int nranks, ncclComm_t nccl, ufe_nd_solver **out) {
  *out = nullptr;
  if (!T || !ptr || !ind || N != 2 * T->nT) { ufe_set_error("ufe_nd_solver_create: bad argument"); return UFE_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    ufe_set_error("ufe_nd_solver_create: no CUDA device (there is no CPU fallback)"); return UFE_ERR_CUDA;
  }
  const int nn = (int)T->nodes.size(), nl = T->n_levels;
  std::vector<int> owner, span;      // sub-trees per rank, left spine above them (ufe_nd_owner_map)
  UFE_TRY(ufe_nd_owner_map(T, nranks, owner, span));
  ufe_nd_solver *S = new ufe_nd_solver();
  S->nT = T->nT; S->N = N; S->nnz = ptr[N] - 1; S->rank = rank; S->nranks = nranks; S->nccl = nccl;
  if (const char *e = getenv("UFE_ND_GRAPHS")) S->use_graphs = atoi(e);
  if (const char *e = getenv("UFE_ND_K64")) S->k64 = atoi(e);
  if (const char *e = getenv("UFE_ND_CLUSTER_FRONTS")) S->cl_max_fronts = atoi(e);
  if (const char *e = getenv("UFE_ND_CLUSTER_MING")) S->cl_min_g = atoi(e);
  if (const char *e = getenv("UFE_ND_UPD_BIG")) S->upd_big = atoi(e);
  if (const char *e = getenv("UFE_ND_UPD_NQ")) S->upd_nq = atoi(e);
  if (const char *e = getenv("UFE_ND_SCHUR_MIN_P")) S->schur_min_p = atoi(e);
  if (const char *e = getenv("UFE_ND_UPD_MMA")) S->upd_mma = atoi(e);
  if (const char *e = getenv("UFE_ND_UPD_MMA_MIN_MODE")) S->upd_mma_min_mode = atoi(e);
  if (const char *e = getenv("UFE_ND_LAZY_ZERO")) S->lazy_zero = atoi(e);
  if (const char *e = getenv("UFE_ND_BIG_MIN_CTAS")) S->big_min_ctas = atoi(e);      // tests: large-tile kernels on small meshes
  if (nranks > 1) S->use_graphs = 0;
  S->lev.resize(nl);
  // local fronts, level-major, p descending inside a level
  std::vector<int> lid(nn, -1);
  std::vector<std::vector<int>> by_level(nl);
  for (int q = 0; q < nn; q++) if (owner[q] == rank) by_level[T->nodes[q].level].push_back(q);
  auto pad32 = [](int v) { return std::max(MFB, (v + MFB - 1) / MFB * MFB); };
  std::vector<int> node_of_front;
  for (int l = 0; l < nl; l++) {
    std::stable_sort(by_level[l].begin(), by_level[l].end(), [&](int a, int b) { return T->nodes[a].sep.size() > T->nodes[b].sep.size(); });
    S->lev[l].first = (int)node_of_front.size(); S->lev[l].n = (int)by_level[l].size();
    for (int q : by_level[l]) { lid[q] = (int)node_of_front.size(); node_of_front.push_back(q); }
  }
  const int nf = (int)node_of_front.size();
  S->n_fronts = nf;
  std::vector<int> ns(nf), p(nf), nb(nf), G(nf), ld(nf), woff(nf), dioff(nf), sep_off(nf), up_off(nf), pwoff(nf, 0);
  std::vector<long long> foff(nf), cf[2];
  std::vector<int> cld[2], cw[2];
  for (int s = 0; s < 2; s++) { cf[s].assign(nf, -1); cld[s].assign(nf, 0); cw[s].assign(nf, -1); }
  size_t fo = 0, wo = 0, dio = 0, so = 0, uo = 0;
  for (int f = 0; f < nf; f++) {
    const NdNode &nd = T->nodes[node_of_front[f]];
    ns[f] = 2 * (int)nd.sep.size(); p[f] = pad32(ns[f]); nb[f] = nd.parent >= 0 ? 2 * (int)nd.bnd.size() : 0;
    G[f] = p[f] + nb[f]; ld[f] = (G[f] + 7) / 8 * 8;
    foff[f] = (long long)fo; fo += (size_t)ld[f] * ld[f];
    woff[f] = (int)wo; wo += ld[f];
    dioff[f] = (int)dio; dio += p[f] / MFB;
    sep_off[f] = (int)so; so += ns[f];
    up_off[f] = (int)uo; uo += nb[f];
    const double pp = p[f], bb = nb[f];
    S->flops += 2.0 / 3.0 * pp * pp * pp + 2.0 * pp * pp * bb + 2.0 * pp * bb * bb;
  }
  // exchange buffers at the cut levels (child 1 of a node whose ranks span > 1 lives on another rank)
  for (int q = 0; q < nn; q++) {
    const NdNode &nd = T->nodes[q];
    if (span[q] <= 1 || nd.child[1] < 0) continue;
    const int c = nd.child[1];
    if (owner[q] != rank && owner[c] != rank) continue;
    MfLink L;
    L.level = nd.level; L.child_rank = owner[c]; L.parent_rank = owner[q];
    L.child_front = owner[c] == rank ? lid[c] : -1; L.parent_front = owner[q] == rank ? lid[q] : -1;
    L.nb = 2 * (int)T->nodes[c].bnd.size();
    L.Gp = pad32(2 * (int)nd.sep.size()) + (nd.parent >= 0 ? 2 * (int)nd.bnd.size() : 0);
    L.f_buf = (long long)fo; fo += ((size_t)L.nb * L.nb + 7) / 8 * 8;
    L.w_buf = (int)wo; wo += (size_t)((owner[q] == rank ? L.nb : L.Gp) + 7) / 8 * 8;
    S->links.push_back(L);
  }
  S->f_doubles = fo; S->w_doubles = wo; S->di_blocks = dio;
  if (wo > 0x7fffffffULL) { ufe_set_error("nd_lu: work vectors exceed 2^31 entries"); delete S; return UFE_ERR_INVALID; }
  // children / parent views, inverse maps, sep dofs, up maps
  std::vector<int> sepdof(so), upmap(uo), pinv0(wo, -1), pinv1(wo, -1);
  for (int f = 0; f < nf; f++) {
    const int q = node_of_front[f];
    const NdNode &nd = T->nodes[q];
    for (size_t k = 0; k < nd.sep.size(); k++) { sepdof[sep_off[f] + 2 * k] = 2 * nd.sep[k]; sepdof[sep_off[f] + 2 * k + 1] = 2 * nd.sep[k] + 1; }
    if (nd.parent >= 0) {
      const NdNode &pa = T->nodes[nd.parent];
      const int pns = (int)pa.sep.size(), pp = pad32(2 * pns);
      for (size_t k = 0; k < nd.bnd.size(); k++) {
        const int u = nd.up[k], row = u < pns ? 2 * u : pp + 2 * (u - pns);
        upmap[up_off[f] + 2 * k] = row; upmap[up_off[f] + 2 * k + 1] = row + 1;
      }
      if (owner[nd.parent] == rank) pwoff[f] = woff[lid[nd.parent]];
    }
    for (int s = 0; s < 2; s++) {
      const int c = nd.child[s];
      if (c < 0) continue;
      const NdNode &ch = T->nodes[c];
      const int pns = (int)nd.sep.size();
      std::vector<int> &pv = s == 0 ? pinv0 : pinv1;
      for (size_t k = 0; k < ch.bnd.size(); k++) {
        const int u = ch.up[k], row = u < pns ? 2 * u : p[f] + 2 * (u - pns);
        pv[woff[f] + row] = 2 * (int)k; pv[woff[f] + row + 1] = 2 * (int)k + 1;
      }
      if (owner[c] == rank) {
        const int cfr = lid[c];
        cf[s][f] = foff[cfr] + (long long)p[cfr] * ld[cfr] + p[cfr]; cld[s][f] = ld[cfr]; cw[s][f] = woff[cfr] + p[cfr];
      }
    }
  }
  for (MfLink &L : S->links) {
    if (L.parent_front >= 0) {
      cf[1][L.parent_front] = L.f_buf; cld[1][L.parent_front] = L.nb; cw[1][L.parent_front] = L.w_buf;
      L.parent_w = woff[L.parent_front];
    }
    if (L.child_front >= 0) {
      const int c = L.child_front;
      pwoff[c] = L.w_buf;
      L.child_schur = foff[c] + (long long)p[c] * ld[c] + p[c]; L.child_ld = ld[c]; L.child_wb = woff[c] + p[c];
    }
  }
  S->lev_lo.assign(nl, 0); S->lev_hi.assign(nl, 0);
  for (int l = 0; l < nl; l++) {
    const MfLevel &Lv = S->lev[l];
    if (Lv.n == 0) continue;
    const int fl = Lv.first + Lv.n - 1;
    S->lev_lo[l] = foff[Lv.first]; S->lev_hi[l] = foff[fl] + (long long)ld[fl] * ld[fl];
  }
  // per-level / per-step launch shapes
  for (int l = 0; l < nl; l++) {
    MfLevel &Lv = S->lev[l];
    for (int z = 0; z < Lv.n; z++) {
      const int f = Lv.first + z;
      Lv.steps = std::max(Lv.steps, p[f] / MFB); Lv.maxG = std::max(Lv.maxG, G[f]); Lv.max_nb = std::max(Lv.max_nb, nb[f]);
      Lv.pmax = std::max(Lv.pmax, p[f]);
      if (cf[0][f] >= 0 || cf[1][f] >= 0) Lv.n_internal++;
    }
    Lv.late_schur = S->schur_min_p > 0 && Lv.pmax >= S->schur_min_p && Lv.max_nb > 0;
    Lv.step.resize(Lv.steps);
    for (int b = 0; b < Lv.steps; b++) {
      MfStep st{0, 0};
      for (int z = 0; z < Lv.n; z++) {
        const int f = Lv.first + z;
        if (p[f] > b * MFB) { st.n_active = z + 1; st.max_trail = std::max(st.max_trail, G[f] - (b + 1) * MFB); }
      }
      Lv.step[b] = st;
    }
    if ((size_t)Lv.maxG * sizeof(double) > 200 * 1024) { ufe_set_error("ufe_nd_solver_create: a front of %d unknowns does not fit the solve kernels' shared memory", Lv.maxG); delete S; return UFE_ERR_INVALID; }
    if (Lv.n > 65535) { ufe_set_error("ufe_nd_solver_create: %d fronts in one level (limit 65535): use larger leaves", Lv.n); delete S; return UFE_ERR_INVALID; }
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if ((double)S->f_doubles * 8.0 > 0.92 * (double)free_b) {
    ufe_set_error("ufe_nd_solver_create: the fronts need %.1f GB, %.1f GB are free", S->f_doubles * 8e-9, free_b * 1e-9);
    delete S; return UFE_ERR_INVALID;
  }
  // destination of every scalar entry: its block entry's front and position there (-1: a front of another rank)
  std::vector<long long> dst(S->nnz);
  std::atomic<int> bad_i{-1}, bad_j{-1};
  ufe_nd_host::parallel_for(N, [&](int ia, int ib) {
  for (int i = ia; i < ib; i++) {
    const int bi = i >> 1;
    for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) {
      const int j = ind[k] - 1, bj = j >> 1;
      int e = -1;
      if (j >= 0 && j < N) {
        const int *lo = T->bind.data() + T->bptr[bi], *hi = T->bind.data() + T->bptr[bi + 1];
        const int *it = std::lower_bound(lo, hi, bj);
        if (it != hi && *it == bj) e = (int)(it - T->bind.data());
        else for (int t = T->bptr[bi]; t < T->bptr[bi + 1]; t++) if (T->bind[t] == bj) { e = t; break; }   // unsorted block rows
      }
      if (e < 0) { bad_i = i; bad_j = j; return; }
      const int q = T->entry_node[e];
      if (owner[q] != rank) { dst[k] = -1; continue; }
      const int f = lid[q];
      const int nsb = ns[f] / 2;
      const int br = T->entry_row[e], bc = T->entry_col[e];
      const int row = (br < nsb ? 2 * br : p[f] + 2 * (br - nsb)) + (i & 1), col = (bc < nsb ? 2 * bc : p[f] + 2 * (bc - nsb)) + (j & 1);
      dst[k] = foff[f] + (long long)row * ld[f] + col;
    }
  }
  });
  if (bad_i >= 0) { ufe_set_error("ufe_nd_solver_create: entry (%d,%d) is not in the analysed block pattern", bad_i + 1, bad_j + 1); delete S; return UFE_ERR_INVALID; }
  int rc = UFE_OK;
  auto fail = [&](int c) { ufe_nd_solver_free(S); return c; };
  if ((rc = mf_upload(&S->ns, ns)) || (rc = mf_upload(&S->p, p)) || (rc = mf_upload(&S->nb, nb)) || (rc = mf_upload(&S->G, G)) ||
      (rc = mf_upload(&S->ld, ld)) || (rc = mf_upload(&S->woff, woff)) || (rc = mf_upload(&S->dioff, dioff)) ||
      (rc = mf_upload(&S->sep_off, sep_off)) || (rc = mf_upload(&S->up_off, up_off)) || (rc = mf_upload(&S->pwoff, pwoff)) ||
      (rc = mf_upload(&S->foff, foff)) || (rc = mf_upload(&S->c_f[0], cf[0])) || (rc = mf_upload(&S->c_f[1], cf[1])) ||
      (rc = mf_upload(&S->c_ld[0], cld[0])) || (rc = mf_upload(&S->c_ld[1], cld[1])) || (rc = mf_upload(&S->c_w[0], cw[0])) ||
      (rc = mf_upload(&S->c_w[1], cw[1])) || (rc = mf_upload(&S->pinv[0], pinv0)) || (rc = mf_upload(&S->pinv[1], pinv1)) ||
      (rc = mf_upload(&S->sepdof, sepdof)) || (rc = mf_upload(&S->upmap, upmap)) || (rc = mf_upload(&S->dst, dst)))
    return fail(rc);
  if (d_ptr && d_ind) { S->ptr = const_cast<int *>(d_ptr); S->ind = const_cast<int *>(d_ind); }
  else {
    std::vector<int> hp(ptr, ptr + N + 1), hi(ind, ind + S->nnz);
    S->own_pattern = true;
    if ((rc = mf_upload(&S->ptr, hp)) || (rc = mf_upload(&S->ind, hi))) return fail(rc);
  }
  const struct { double **q; size_t n; } bufs[] = {{&S->scale, (size_t)N}, {&S->dself, (size_t)N}, {&S->dpair, (size_t)N}, {&S->F, S->f_doubles},
      {&S->W, S->w_doubles}, {&S->Dinv, S->di_blocks * MFB * MFB}, {&S->b, (size_t)N}, {&S->x, (size_t)N}, {&S->r, (size_t)N}};
  for (const auto &bf : bufs)
    if (cudaMalloc((void **)bf.q, std::max<size_t>(1, bf.n) * sizeof(double)) != cudaSuccess) {
      ufe_set_error("ufe_nd_solver_create: out of device memory (%zu doubles)", bf.n); cudaGetLastError(); return fail(UFE_ERR_CUDA);
    }
  UFE_CUDA(cudaMemset(S->W, 0, S->w_doubles * sizeof(double)));
  if (cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&S->e0) != cudaSuccess ||
      cudaEventCreate(&S->e1) != cudaSuccess) { ufe_set_error("ufe_nd_solver_create: stream / event creation failed"); return fail(UFE_ERR_CUDA); }
  static bool attr_set = false;
  if (!attr_set) {
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<16, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(128, 128))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<16, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(128, 128))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<16, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(128, 64))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<16, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(128, 64))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<8, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(64, 64))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<8, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(64, 64))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<8, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(64, 32))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_UPD_SMEM(64, 32))));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_update_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MF_MMA_SMEM));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_fwd_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    UFE_CUDA(cudaFuncSetAttribute(k_mf_bwd_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  UFE_CUDA(cudaDeviceSynchronize());
  *out = S;
  return UFE_OK;
}

extern "C" int ufe_nd_solver_create(const ufe_nd_tree *T, int32_t N, const int32_t *ptr, const int32_t *ind, ufe_nd_solver **out) {
  if (!out) { ufe_set_error("ufe_nd_solver_create: bad argument"); return UFE_ERR_INVALID; }
  ufe_nd_solver *S = nullptr;
  UFE_TRY(mf_create(T, N, ptr, ind, nullptr, nullptr, 0, 1, nullptr, &S));
  if (cudaMalloc((void **)&S->val, std::max<size_t>(1, (size_t)S->nnz) * sizeof(double)) != cudaSuccess) {
    ufe_set_error("ufe_nd_solver_create: out of device memory"); ufe_nd_solver_free(S); return UFE_ERR_CUDA;
  }
  *out = S;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// launch sequences
// ------------------------------------------------------------------------------------------------------------------
// trailing-update launch: extent = largest trailing (mode 0, 1) or boundary (mode 2) extent of the n fronts
static int mf_launch_update(ufe_nd_solver *S, cudaStream_t st, const MfLevel &L, bool big, int nq, int extent, int n, int b, int nkb, int mode) {
  // large trailing matrices: 128 x 64 tiles, 8 x 4 per thread, two CTAs (16 warps) per SM -- 6 % faster at 1 M vertices
  // than 128 x 128 / 8 x 8 with one CTA per SM; small ones: 64 x 64 tiles of 64 threads (8 x 8 per thread)
  if (big && S->upd_mma && mode >= S->upd_mma_min_mode) {      // fp64 tensor cores, 128 x 128 tiles
    k_mf_update_mma<<<dim3((extent + 127) / 128, (extent + 127) / 128, n), 256, MF_MMA_SMEM, st>>>(L.first, b, nkb, mode, S->p, S->G, S->ld, S->foff, S->F);
    UFE_LAUNCH_CHECK();
    return UFE_OK;
  }
  const int TR = big ? 128 : 64, TC = (big ? 128 : 64) / (nq == 2 ? 2 : 1);
  const dim3 grid((extent + TC - 1) / TC, (extent + TR - 1) / TR, n);
#define MF_UPD_LAUNCH(TG_, NQ_, NT_, TR_, TC_)                                                                                      \
  do {                                                                                                                             \
    if (mode) k_mf_update<TG_, NQ_, true><<<grid, NT_, MF_UPD_SMEM(TR_, TC_), st>>>(L.first, b, nkb, mode, S->p, S->G, S->ld, S->foff, S->F); \
    else k_mf_update<TG_, NQ_, false><<<grid, NT_, MF_UPD_SMEM(TR_, TC_), st>>>(L.first, b, nkb, 0, S->p, S->G, S->ld, S->foff, S->F);       \
  } while (0)
  if (big && nq == 2) MF_UPD_LAUNCH(16, 2, 256, 128, 64);
  else if (big) MF_UPD_LAUNCH(16, 4, 256, 128, 128);
  else if (nq == 2) MF_UPD_LAUNCH(8, 2, 64, 64, 32);
  else MF_UPD_LAUNCH(8, 4, 64, 64, 64);
#undef MF_UPD_LAUNCH
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

static int mf_factor_launches(ufe_nd_solver *S, cudaStream_t st, const double *dval) {
  const int N = S->N, tb = 256, nf = S->n_fronts;
  k_mf_scale<<<(N + tb - 1) / tb, tb, 0, st>>>(N, S->ptr, S->ind, dval, S->scale, S->dself, S->dpair); UFE_LAUNCH_CHECK();
  // Large problems: the fronts are zeroed level by level -- an internal level is WRITTEN by the extend-add (children's
  // contributions or 0), then the matrix entries of that level are added -- instead of one memset of everything, one
  // scatter and a read-modify-write extend-add per level (1 M vertices: 44 GB of fronts, 7 ms per avoided pass).
  const bool lazy = S->lazy_zero >= 0 ? S->lazy_zero != 0 : S->f_doubles * sizeof(double) > ((size_t)1 << 30);
  if (!lazy) {
    UFE_CUDA(cudaMemsetAsync(S->F, 0, S->f_doubles * sizeof(double), st));
    k_mf_assemble<<<(N + tb - 1) / tb, tb, 0, st>>>(N, S->ptr, S->ind, dval, S->scale, S->dst, S->F, 0LL, (long long)S->f_doubles); UFE_LAUNCH_CHECK();
    if (nf > 0) { k_mf_pad<<<(nf + 7) / 8, 256, 0, st>>>(0, nf, S->ns, S->p, S->ld, S->foff, S->F); UFE_LAUNCH_CHECK(); }
  }
  for (int l = (int)S->lev.size() - 1; l >= 0; l--) {
    const MfLevel &L = S->lev[l];
    // Schur complements of children on other ranks arrive first (the level below is complete on every rank)
    for (const MfLink &K : S->links) {
      if (K.level != l || K.nb == 0) continue;
      const size_t cnt = (size_t)K.nb * K.nb;
      if (K.child_front >= 0) {
        k_mf_pack<<<dim3((K.nb + 31) / 32, (K.nb + 31) / 32), dim3(32, 8), 0, st>>>(K.nb, K.child_ld, S->F + K.child_schur, S->F + K.f_buf);
        UFE_LAUNCH_CHECK();
        UFE_NCCL(ncclSend(S->F + K.f_buf, cnt, ncclDouble, K.parent_rank, S->nccl, st));
      } else {
        UFE_NCCL(ncclRecv(S->F + K.f_buf, cnt, ncclDouble, K.child_rank, S->nccl, st));
      }
      g_launch_count++;
    }
    if (L.n == 0) continue;
    if (lazy) {
      const long long lo = S->lev_lo[l], hi = S->lev_hi[l];
      if (L.n_internal > 0) {
        const int t = (L.maxG + 31) / 32;
        k_mf_extend<true><<<dim3(t, t, L.n), dim3(32, 8), 0, st>>>(L.first, S->G, S->ld, S->foff, S->woff, S->c_f[0], S->c_f[1], S->c_ld[0],
                                                                    S->c_ld[1], S->pinv[0], S->pinv[1], S->F);
        UFE_LAUNCH_CHECK();
      } else UFE_CUDA(cudaMemsetAsync(S->F + lo, 0, (size_t)(hi - lo) * sizeof(double), st));
      k_mf_pad<<<(L.n + 7) / 8, 256, 0, st>>>(L.first, L.n, S->ns, S->p, S->ld, S->foff, S->F); UFE_LAUNCH_CHECK();
      k_mf_assemble<<<(N + tb - 1) / tb, tb, 0, st>>>(N, S->ptr, S->ind, dval, S->scale, S->dst, S->F, lo, hi); UFE_LAUNCH_CHECK();
    } else if (L.n_internal > 0) {
      const int t = (L.maxG + 31) / 32;
      k_mf_extend<false><<<dim3(t, t, L.n), dim3(32, 8), 0, st>>>(L.first, S->G, S->ld, S->foff, S->woff, S->c_f[0], S->c_f[1], S->c_ld[0],
                                                                   S->c_ld[1], S->pinv[0], S->pinv[1], S->F);
      UFE_LAUNCH_CHECK();
    }
    for (int b = 0; b < L.steps;) {
      const MfStep &sp = L.step[b];
      // few fronts: split the L panel of a front over several CTAs (each inverts the pivot block itself)
      auto split = [&](int n_active, int extent, int *chunk) {
        int nx = 1; *chunk = 1 << 28;
        if (n_active < 96 && extent > 64) {
          nx = std::min((extent + MFB - 1) / MFB, std::max(1, 296 / n_active));
          *chunk = ((extent + nx - 1) / nx + MFB - 1) / MFB * MFB;
          nx = (extent + *chunk - 1) / *chunk;
        }
        return std::max(nx, 1);
      };
      int chunk = 0;
      const int nx = split(sp.n_active, sp.max_trail, &chunk);
      k_mf_panel<<<dim3(nx, sp.n_active), MF_PANEL_THREADS, 0, st>>>(L.first, b, chunk, S->G, S->ld, S->foff, S->dioff, S->F, S->Dinv);
      UFE_LAUNCH_CHECK();
      // second block of a 64-wide outer step: only when the same fronts are active in it (they are a prefix, too)
      int nkb = 1, n_upd = sp.n_active, trail = sp.max_trail;
      if (S->k64 && b + 1 < L.steps && L.step[b + 1].n_active == sp.n_active) {
        const MfStep &s2 = L.step[b + 1];
        int ch2 = 0;
        const int nx2 = split(s2.n_active, s2.max_trail, &ch2);
        k_mf_panel2<<<dim3(2 * nx2, s2.n_active), MF_PANEL_THREADS, 0, st>>>(L.first, b + 1, nx2, ch2, S->G, S->ld, S->foff, S->dioff, S->F, S->Dinv);
        UFE_LAUNCH_CHECK();
        nkb = 2; n_upd = s2.n_active; trail = s2.max_trail;
      }
      b += nkb;
      if (trail <= 0) continue;
      const long long big_ctas = (long long)((trail + 127) / 128) * ((trail + 127) / 128) * n_upd;
      const bool big = S->upd_big && trail >= 192 && big_ctas >= S->big_min_ctas;
      // large trailing matrices: 128 x 64 tiles, 8 x 4 per thread, two CTAs (16 warps) per SM -- 6 % faster at 1 M vertices
      // than 128 x 128 / 8 x 8 with one CTA per SM; small ones: 64 x 64 tiles of 64 threads (8 x 8 per thread)
      const int nq = S->upd_nq ? S->upd_nq : (big ? 2 : 4);
      UFE_TRY(mf_launch_update(S, st, L, big, nq, trail, n_upd, b - nkb, nkb, L.late_schur ? 1 : 0));
    }
    if (L.late_schur) {        // F22 -= L21 U12, one pass with K = p
      const long long big_ctas = (long long)((L.max_nb + 127) / 128) * ((L.max_nb + 127) / 128) * L.n;
      const bool big = S->upd_big && L.max_nb >= 192 && big_ctas >= S->big_min_ctas;
      UFE_TRY(mf_launch_update(S, st, L, big, S->upd_nq ? S->upd_nq : (big ? 2 : 4), L.max_nb, L.n, 0, 0, 2));
    }
  }
  return UFE_OK;
}

// x (+)= A^-1 r with the factors; r: full-length right-hand side, x: full-length solution (this rank writes the unknowns
// of its own fronts only).  premul: r is first multiplied by the 2x2 diagonal blocks.
static int mf_apply_launches(ufe_nd_solver *S, cudaStream_t st, const double *r, double *x, int accumulate, int premul) {
  const int nl = (int)S->lev.size();
  for (int l = nl - 1; l >= 0; l--) {            // forward sweep, deepest level first
    const MfLevel &L = S->lev[l];
    for (const MfLink &K : S->links) {           // boundary vectors of children on other ranks
      if (K.level != l || K.nb == 0) continue;
      if (K.child_front >= 0) UFE_NCCL(ncclSend(S->W + K.child_wb, (size_t)K.nb, ncclDouble, K.parent_rank, S->nccl, st));
      else UFE_NCCL(ncclRecv(S->W + K.w_buf, (size_t)K.nb, ncclDouble, K.child_rank, S->nccl, st));
      g_launch_count++;
    }
    if (L.n == 0) continue;
    if (L.n <= S->cl_max_fronts && L.maxG >= S->cl_min_g)
      k_mf_fwd_cl<<<L.n * MFCL, 1024, (size_t)L.maxG * sizeof(double), st>>>(L.first, S->ns, S->p, S->G, S->ld, S->foff, S->woff, S->c_w[0], S->c_w[1], S->pinv[0], S->pinv[1],
                                        S->sep_off, S->sepdof, S->scale, r, premul ? S->dself : nullptr, S->premul_pair ? S->dpair : nullptr, S->F, S->W);
    else
      k_mf_fwd<<<L.n, 1024, (size_t)L.maxG * sizeof(double), st>>>(L.first, S->ns, S->p, S->G, S->ld, S->foff, S->woff, S->c_w[0], S->c_w[1], S->pinv[0], S->pinv[1],
                                   S->sep_off, S->sepdof, S->scale, r, premul ? S->dself : nullptr, S->premul_pair ? S->dpair : nullptr, S->F, S->W);
    UFE_LAUNCH_CHECK();
  }
  for (int l = 0; l < nl; l++) {                 // backward sweep, root first
    const MfLevel &L = S->lev[l];
    if (L.n > 0) {
      if (L.n <= S->cl_max_fronts && L.maxG >= S->cl_min_g)
        k_mf_bwd_cl<<<L.n * MFCL, 1024, (size_t)L.maxG * sizeof(double), st>>>(L.first, S->ns, S->p, S->nb, S->ld, S->foff, S->woff, S->pwoff, S->up_off, S->upmap, S->dioff,
                                          S->sep_off, S->sepdof, S->scale, S->F, S->Dinv, S->W, x, accumulate);
      else
        k_mf_bwd<<<L.n, 1024, (size_t)L.maxG * sizeof(double), st>>>(L.first, S->ns, S->p, S->nb, S->ld, S->foff, S->woff, S->pwoff, S->up_off, S->upmap, S->dioff,
                                     S->sep_off, S->sepdof, S->scale, S->F, S->Dinv, S->W, x, accumulate);
      UFE_LAUNCH_CHECK();
    }
    for (const MfLink &K : S->links) {           // the parent's solution vector goes to the child on the other rank
      if (K.level != l || K.nb == 0) continue;
      if (K.parent_front >= 0) UFE_NCCL(ncclSend(S->W + K.parent_w, (size_t)K.Gp, ncclDouble, K.child_rank, S->nccl, st));
      else UFE_NCCL(ncclRecv(S->W + K.w_buf, (size_t)K.Gp, ncclDouble, K.parent_rank, S->nccl, st));
      g_launch_count++;
    }
  }
  return UFE_OK;
}

// Runs a launch sequence through a cached CUDA graph (single rank): the sequences are static per pattern, so they are
// captured once per distinct argument set and replayed (the small-mesh solves are bound by launch latency otherwise).
// Inside somebody else's capture the kernels are issued directly and become part of that graph.
template <class Fn>
static int mf_run(ufe_nd_solver *S, cudaStream_t st, MfGraph *slots, int nslots, const void *a, const void *b, int k0, int k1, Fn fn) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  UFE_CUDA(cudaStreamIsCapturing(st, &cs));
  if (!S->use_graphs || cs != cudaStreamCaptureStatusNone) return fn();
  MfGraph *g = nullptr;
  for (int i = 0; i < nslots; i++) if (slots[i].exec && slots[i].a == a && slots[i].b == b && slots[i].k0 == k0 && slots[i].k1 == k1) { g = &slots[i]; break; }
  if (!g) {
    for (int i = 0; i < nslots; i++) if (!slots[i].exec) { g = &slots[i]; break; }
    if (!g) { g = &slots[0]; mf_graph_drop(*g); }
    const int64_t l0 = g_launch_count;
    cudaGraph_t graph = nullptr;
    UFE_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = fn();
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != UFE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { ufe_set_error("nd_lu: graph capture failed: %s", cudaGetErrorString(ce)); return UFE_ERR_CUDA; }
    const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { g->exec = nullptr; ufe_set_error("nd_lu: graph instantiation failed: %s", cudaGetErrorString(ie)); return UFE_ERR_CUDA; }
    g->a = a; g->b = b; g->k0 = k0; g->k1 = k1; g->nodes = (int)(g_launch_count - l0);
    g_launch_count = l0;
  }
  UFE_CUDA(cudaGraphLaunch(g->exec, st));
  g_launch_count += g->nodes;
  return UFE_OK;
}

static int mf_factor_device(ufe_nd_solver *S, cudaStream_t st, const double *dval) {
  UFE_TRY(mf_run(S, st, &S->g_factor, 1, dval, nullptr, 0, 0, [&]() { return mf_factor_launches(S, st, dval); }));
  S->factored = true;
  return UFE_OK;
}
static int mf_apply_device(ufe_nd_solver *S, cudaStream_t st, const double *r, double *x, int accumulate, int premul) {
  return mf_run(S, st, S->g_apply, 6, r, x, accumulate, premul, [&]() { return mf_apply_launches(S, st, r, x, accumulate, premul); });
}

// val: host values of A in the order of the pattern given to ufe_nd_solver_create
extern "C" int ufe_nd_solver_factor(ufe_nd_solver *S, const double *val) {
  if (!S || !val || !S->val) { ufe_set_error("ufe_nd_solver_factor: bad argument"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaMemcpyAsync(S->val, val, (size_t)S->nnz * sizeof(double), cudaMemcpyHostToDevice, S->st));
  UFE_CUDA(cudaEventRecord(S->e0, S->st));
  UFE_TRY(mf_factor_device(S, S->st, S->val));
  UFE_CUDA(cudaEventRecord(S->e1, S->st));
  UFE_CUDA(cudaStreamSynchronize(S->st));
  UFE_CUDA(cudaEventElapsedTime(&S->factor_ms, S->e0, S->e1));
  return UFE_OK;
}

// x = A^-1 b followed by n_refine steps of iterative refinement  x += A^-1 (b - A x);  relres = |b - A x| / |b|
extern "C" int ufe_nd_solver_solve(ufe_nd_solver *S, const double *b, double *x, int32_t n_refine, double *relres) {
  if (!S || !b || !x || n_refine < 0) { ufe_set_error("ufe_nd_solver_solve: bad argument"); return UFE_ERR_INVALID; }
  if (!S->factored) { ufe_set_error("ufe_nd_solver_solve: call ufe_nd_solver_factor first"); return UFE_ERR_INVALID; }
  const int N = S->N, tb = 256;
  UFE_CUDA(cudaMemcpyAsync(S->b, b, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, S->st));
  UFE_CUDA(cudaEventRecord(S->e0, S->st));
  UFE_TRY(mf_apply_device(S, S->st, S->b, S->x, 0, 0));
  for (int it = 0; it < n_refine; it++) {
    k_mf_residual<<<(N + tb - 1) / tb, tb, 0, S->st>>>(N, S->ptr, S->ind, S->val, S->b, S->x, S->r); UFE_LAUNCH_CHECK();
    UFE_TRY(mf_apply_device(S, S->st, S->r, S->x, 1, 0));
  }
  UFE_CUDA(cudaEventRecord(S->e1, S->st));
  k_mf_residual<<<(N + tb - 1) / tb, tb, 0, S->st>>>(N, S->ptr, S->ind, S->val, S->b, S->x, S->r); UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaMemcpyAsync(x, S->x, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  std::vector<double> hr(N);
  UFE_CUDA(cudaMemcpyAsync(hr.data(), S->r, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  UFE_CUDA(cudaStreamSynchronize(S->st));
  UFE_CUDA(cudaEventElapsedTime(&S->solve_ms, S->e0, S->e1));
  if (relres) {
    double rr = 0.0, bb = 0.0;
    for (int i = 0; i < N; i++) { rr += hr[i] * hr[i]; bb += b[i] * b[i]; }
    *relres = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
  }
  return UFE_OK;
}

// timings of the last factor / solve call (CUDA events on the solver's stream), storage and flop count of the
// factorisation (2/3 p^3 + 2 p^2 nb + 2 p nb^2 per front, on the 32-padded pivot counts)
extern "C" int ufe_nd_solver_info(const ufe_nd_solver *S, double *factor_ms, double *solve_ms, double *front_bytes, double *factor_flops) {
  if (!S) { ufe_set_error("null solver"); return UFE_ERR_INVALID; }
  if (factor_ms) *factor_ms = S->factor_ms;
  if (solve_ms) *solve_ms = S->solve_ms;
  if (front_bytes) *front_bytes = 8.0 * (double)S->f_doubles;
  if (factor_flops) *factor_flops = S->flops;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// the same solver as the exact preconditioner of the Krylov loop (krylov_pc = UFE_PC_ND_LU; ufe_pclu.cu dispatches here).
// The Krylov loop iterates on B A x = B b with B the 2x2 block-Jacobi scaling folded in at assembly time
// (ufe_assembly.cu), so the preconditioner is  z = (B A)^-1 r = A^-1 (D r),  D = the 2x2 diagonal blocks of A.
// Several ranks: the rows of A are partitioned into contiguous strips (partition_list); pattern (once) and values
// (per factorisation) are gathered so that every rank can assemble the fronts it owns.
// ------------------------------------------------------------------------------------------------------------------
// all-gather of contiguous per-rank pieces of different sizes (rank order = strip order): out = [piece_0 | piece_1 | ...]
static int mf_gatherv(ncclComm_t nccl, cudaStream_t st, int rank, const void *mine, const std::vector<size_t> &bytes, char *out) {
  size_t off = 0;
  UFE_NCCL(ncclGroupStart());
  for (size_t q = 0; q < bytes.size(); q++) {
    UFE_NCCL(ncclBroadcast((int)q == rank ? mine : (const void *)(out + off), out + off, bytes[q], ncclChar, (int)q, nccl, st));
    off += bytes[q];
  }
  UFE_NCCL(ncclGroupEnd());
  g_launch_count++;
  return UFE_OK;
}

int ufe_nd_pc_create(cudaStream_t st, const DevSystem &S, const Comm *comm, int nT, const double *gcx, const double *gcy, int leaf,
                     ufe_nd_solver **out) {
  *out = nullptr;
  const int P = comm ? comm->nranks : 1, me = comm ? comm->rank : 0;
  if (S.N != 2 * nT) { ufe_set_error("nd_lu preconditioner: the system must be the stiffness system of the mesh (N = 2 nTri)"); return UFE_ERR_INVALID; }
  if (P == 1 && (S.m_loc != S.N || S.r1 != 1)) { ufe_set_error("nd_lu preconditioner: one rank must hold all rows"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaStreamSynchronize(st));
  std::vector<int> ptr(S.N + 1), ind;
  std::vector<size_t> vcnt(P, 0), rcnt(P, 0);
  int *g_ptr = nullptr, *g_ind = nullptr;
  if (P == 1) {
    ind.resize(S.nnz);
    UFE_CUDA(cudaMemcpy(ptr.data(), S.ptr, ptr.size() * sizeof(int), cudaMemcpyDeviceToHost));
    UFE_CUDA(cudaMemcpy(ind.data(), S.ind, ind.size() * sizeof(int), cudaMemcpyDeviceToHost));
  } else {
    // sizes first, then row lengths and column indices
    long long mine[2] = {S.m_loc, S.nnz}, *d_sz = nullptr;
    std::vector<long long> all(2 * P);
    UFE_CUDA(cudaMalloc(&d_sz, sizeof(long long) * 2 * (P + 1)));
    UFE_CUDA(cudaMemcpy(d_sz + 2 * P, mine, sizeof mine, cudaMemcpyHostToDevice));
    UFE_NCCL(ncclAllGather(d_sz + 2 * P, d_sz, 2, ncclInt64, comm->nccl, st));
    UFE_CUDA(cudaStreamSynchronize(st));
    UFE_CUDA(cudaMemcpy(all.data(), d_sz, sizeof(long long) * 2 * P, cudaMemcpyDeviceToHost));
    cudaFree(d_sz);
    size_t rows = 0, nnz = 0;
    std::vector<size_t> pb(P), ib(P);
    for (int q = 0; q < P; q++) { rcnt[q] = (size_t)all[2 * q]; vcnt[q] = (size_t)all[2 * q + 1]; pb[q] = (rcnt[q] + 1) * sizeof(int); ib[q] = vcnt[q] * sizeof(int); rows += rcnt[q]; nnz += vcnt[q]; }
    if ((int)rows != S.N) { ufe_set_error("nd_lu preconditioner: the strips of the ranks do not add up to the whole system"); return UFE_ERR_INVALID; }
    int *d_lp = nullptr;
    UFE_CUDA(cudaMalloc(&d_lp, sizeof(int) * (rows + P)));
    UFE_CUDA(cudaMalloc(&g_ind, sizeof(int) * std::max<size_t>(1, nnz)));
    UFE_TRY(mf_gatherv(comm->nccl, st, me, S.ptr, pb, reinterpret_cast<char *>(d_lp)));
    UFE_TRY(mf_gatherv(comm->nccl, st, me, S.ind, ib, reinterpret_cast<char *>(g_ind)));
    UFE_CUDA(cudaStreamSynchronize(st));
    std::vector<int> lp(rows + P);
    UFE_CUDA(cudaMemcpy(lp.data(), d_lp, lp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_lp);
    ind.resize(nnz);
    UFE_CUDA(cudaMemcpy(ind.data(), g_ind, nnz * sizeof(int), cudaMemcpyDeviceToHost));
    size_t r = 0, o = 0, base = 0;
    for (int q = 0; q < P; q++) {                       // local 1-based offsets -> global 1-based offsets
      for (size_t i = 0; i < rcnt[q]; i++) ptr[r + i] = (int)(base + lp[o + i]);
      r += rcnt[q]; o += rcnt[q] + 1; base += vcnt[q];
    }
    ptr[S.N] = (int)(base + 1);
    UFE_CUDA(cudaMalloc(&g_ptr, sizeof(int) * (S.N + 1)));
    UFE_CUDA(cudaMemcpy(g_ptr, ptr.data(), sizeof(int) * (S.N + 1), cudaMemcpyHostToDevice));
  }
  // block pattern over triangles (0-based, sorted)
  std::vector<int> bptr(nT + 1, 0), bind, row;
  bind.reserve(ind.size() / 3);
  for (int t = 0; t < nT; t++) {
    row.clear();
    for (int k = ptr[2 * t] - 1; k < ptr[2 * t + 2] - 1; k++) row.push_back((ind[k] - 1) >> 1);
    std::sort(row.begin(), row.end());
    row.erase(std::unique(row.begin(), row.end()), row.end());
    bind.insert(bind.end(), row.begin(), row.end());
    bptr[t + 1] = (int)bind.size();
  }
  ufe_nd_tree *T = nullptr;
  int rc = ufe_nd_analyse(nT, gcx, gcy, bptr.data(), bind.data(), leaf, &T);
  if (rc == UFE_OK) rc = mf_create(T, S.N, ptr.data(), ind.data(), P == 1 ? S.ptr : g_ptr, P == 1 ? S.ind : g_ind, me, P, comm ? comm->nccl : nullptr, out);
  ufe_nd_tree_free(T);
  if (rc != UFE_OK) { cudaFree(g_ptr); cudaFree(g_ind); return rc; }
  ufe_nd_solver *M = *out;
  if (P > 1) {
    M->own_pattern = true;                 // g_ptr / g_ind belong to the solver now
    M->val_counts = vcnt; M->row_counts = rcnt;
    if (cudaMalloc((void **)&M->val, std::max<size_t>(1, (size_t)M->nnz) * sizeof(double)) != cudaSuccess) {
      ufe_set_error("nd_lu preconditioner: out of device memory"); ufe_nd_solver_free(M); *out = nullptr; return UFE_ERR_CUDA;
    }
  }
  return UFE_OK;
}

// dval: the rows of this rank (S.val)
int ufe_nd_pc_factor(cudaStream_t st, ufe_nd_solver *S, const double *dval) {
  if (S->nranks == 1) return mf_factor_device(S, st, dval);
  std::vector<size_t> bytes(S->nranks);
  for (int q = 0; q < S->nranks; q++) bytes[q] = S->val_counts[q] * sizeof(double);
  UFE_TRY(mf_gatherv(S->nccl, st, S->rank, dval, bytes, reinterpret_cast<char *>(S->val)));
  return mf_factor_device(S, st, S->val);
}

// r, z: the rows of this rank
int ufe_nd_pc_apply(cudaStream_t st, ufe_nd_solver *S, const double *r, double *z) {
  if (S->nranks == 1) return mf_apply_device(S, st, r, z, 0, 1);
  std::vector<size_t> bytes(S->nranks);
  size_t r0 = 0;
  for (int q = 0; q < S->nranks; q++) { bytes[q] = S->row_counts[q] * sizeof(double); if (q < S->rank) r0 += S->row_counts[q]; }
  UFE_TRY(mf_gatherv(S->nccl, st, S->rank, r, bytes, reinterpret_cast<char *>(S->b)));
  UFE_CUDA(cudaMemsetAsync(S->x, 0, (size_t)S->N * sizeof(double), st));
  UFE_TRY(mf_apply_device(S, st, S->b, S->x, 0, 1));
  UFE_NCCL(ncclAllReduce(S->x, S->x, (size_t)S->N, ncclDouble, ncclSum, S->nccl, st));
  g_launch_count++;
  const int n = (int)S->row_counts[S->rank];
  if (n > 0) { k_mf_copy_slice<<<(n + 255) / 256, 256, 0, st>>>(n, S->x + r0, z); UFE_LAUNCH_CHECK(); }
  return UFE_OK;
}

void ufe_nd_pc_set_point_scaling(ufe_nd_solver *S) { S->premul_pair = 0; }

// storage and flop count of this rank's part of the factorisation
void ufe_nd_pc_info(const ufe_nd_solver *S, double *front_bytes, double *flops, int *n_fronts) {
  if (front_bytes) *front_bytes = 8.0 * (double)S->f_doubles;
  if (flops) *flops = S->flops;
  if (n_fronts) *n_fronts = S->n_fronts;
}
