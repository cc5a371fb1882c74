// Numeric phase of the multifrontal nested-dissection solver (device code).
//
// Exact solve of the DIVA / SSA stiffness system  A x = b  (solve_linearised_SSA_DIVA.f90:159 hands exactly this
// system to PETSc) for meshes whose x-sorted bandwidth is too wide for the banded block cyclic reduction of
// ufe_pclu.cu.  The symbolic analysis (ufe_nd.cu) gives a binary elimination tree; every tree node owns one dense
// front  [sep; bnd] x [sep; bnd]  (scalar unknowns: triangle t -> 2t, 2t+1, mesh_translation_tables.f90:181-198).
//
// Layout in HBM.  All fronts of one tree level are stored as ONE batch of row-major g_l x g_l matrices, g_l and the
// pivot extent p_l being the level maxima rounded to 64 / 32:
//      rows [0, p_l)   : the front's own unknowns (sep), padded with identity rows,
//      rows [p_l, g_l) : the boundary unknowns (bnd), padded with zero rows,
// so a level is a handful of batched launches (blockIdx.z = front) with no per-front shapes in the kernels.
//
// Factorisation, deepest level first.  A blocked in-place Gauss-Jordan sweep over the pivot blocks of [0, p_l) only
// (32-wide panels, partial pivoting inside the 32 x 32 pivot block) leaves, with X = F11^-1,
//      F11 <- X,    F12 <- X F12,    F21 <- -F21 X,    F22 <- F22 - F21 X F12   (the Schur complement),
// and F22 is then added into the parent's front through the child -> parent index map (child slot 0, then slot 1:
// the sums have a fixed order, so the factors are bit-reproducible).
// Solve.  Every front has a work vector w (length g_l):  up the tree  y = F[:, 0:p] w_s;  z = y_s is kept,
// w_b += y_b goes to the parent's w; down the tree  x_b comes from the parent,  x_s = z - F12' x_b.
// The matrix is Jacobi-scaled symmetrically while it is assembled into the fronts (unit-magnitude diagonal, so the
// pivot thresholds are scale-free); ufe_nd_solver_solve applies steps of iterative refinement with the unscaled CSR.
#include "ufe_internal.cuh"
#include "ufe_nd.cuh"

#include <algorithm>

#define NDB 32    // pivot panel width
#define NDT 64    // trailing-update tile

struct NdLevelDev {
  int n = 0, g = 0, p = 0, first = 0;   // fronts, padded front size, padded pivot extent, index of the level's first front
  int max_nb = 0;                       // largest boundary (scalars) in the level
  size_t f_off = 0, w_off = 0;          // offsets (doubles) of the level's batch in F and in W / Z
};

struct ufe_nd_solver {
  int nT = 0, N = 0, nnz = 0, n_fronts = 0;
  std::vector<NdLevelDev> lev;
  int *ns = nullptr, *nb = nullptr, *parent = nullptr, *slot = nullptr;   // per front (level-major order)
  int *sep_off = nullptr, *up_off = nullptr;                              // per front: offsets into sepdof / upmap
  int *sepdof = nullptr, *upmap = nullptr;   // global unknown of every sep row; parent row of every bnd row
  long long *dst = nullptr;                  // per scalar CSR entry: destination in F
  int *ptr = nullptr, *ind = nullptr;        // 0-based scalar CSR of A
  double *val = nullptr, *scale = nullptr, *dself = nullptr, *dpair = nullptr;
  double *F = nullptr, *W = nullptr, *Z = nullptr, *ipp = nullptr, *colbuf = nullptr;
  double *b = nullptr, *x = nullptr, *r = nullptr, *dx = nullptr;
  size_t f_doubles = 0, w_doubles = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float factor_ms = 0.f, solve_ms = 0.f;
  double flops = 0.0;
  bool factored = false;
};

// ------------------------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------------------------
// also keeps row i of the 2x2 (u,v) diagonal block of its triangle: dself = a_ii, dpair = a_i,i^1
__global__ void k_nd_scale(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                           double *__restrict__ scale, double *__restrict__ dself, double *__restrict__ dpair) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double d = 0.0, o = 0.0;
  for (int k = ptr[i]; k < ptr[i + 1]; k++) { if (ind[k] == i) d += val[k]; else if (ind[k] == (i ^ 1)) o += val[k]; }
  dself[i] = d; dpair[i] = o;
  d = fabs(d);
  scale[i] = d > 0.0 ? 1.0 / sqrt(d) : 1.0;
}

// one thread per matrix row: every (row, column) has its own destination, so plain adds are race-free
__global__ void k_nd_assemble(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                              const double *__restrict__ scale, const long long *__restrict__ dst, double *__restrict__ F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double si = scale[i];
  for (int k = ptr[i]; k < ptr[i + 1]; k++) F[dst[k]] += si * val[k] * scale[ind[k]];
}

// identity on the padded pivot rows [ns, p) of every front of a level
__global__ void k_nd_pad_identity(int g, int p, int first, const int *__restrict__ ns, double *__restrict__ Fl) {
  const int z = blockIdx.y, r = ns[first + z] + blockIdx.x * blockDim.x + threadIdx.x;
  if (r < p) Fl[(size_t)z * g * g + (size_t)r * g + r] = 1.0;
}

// ------------------------------------------------------------------------------------------------------------------
// partial Gauss-Jordan sweep of a level's fronts
// ------------------------------------------------------------------------------------------------------------------
// 32 x 32 inverse in shared memory by 1024 threads (i, j): Gauss-Jordan on [S | I2], partial pivoting as a row
// permutation, searched only when the natural pivot is small (the fronts are Jacobi-scaled); vanishing pivots are
// perturbed statically -- the refinement / Krylov iteration around the solver absorbs that.
__device__ __forceinline__ void nd_invert32(double (*S)[NDB + 1], double (*I2)[NDB + 1], double (*Ip)[NDB + 1], int *perm,
                                            int i, int j) {
  I2[i][j] = (i == j) ? 1.0 : 0.0;
  if (i == 0) perm[j] = j;
  __syncthreads();
  for (int p = 0; p < NDB; p++) {
    if (fabs(S[perm[p]][p]) < 0.05) {
      __syncthreads();
      if (i == 0) {
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
    const int P = perm[p];
    double piv = S[P][p];
    if (fabs(piv) < 1e-13) piv = piv < 0.0 ? -1e-13 : 1e-13;
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

// panel step for pivot block b; every CTA inverts A_PP itself (A_PP is not written here).
//   y = 0, tile (P, J = x):  A_PJ <- Ipp A_PJ        (x == b: publish Ipp for the update kernel)
//   y = 1, tile (I = x, P):  colbuf <- A_IP (old),  A_IP <- -A_IP Ipp
__global__ void __launch_bounds__(1024)
k_nd_panel(int g, int b, double *__restrict__ Fl, double *__restrict__ ipp, double *__restrict__ colbuf) {
  const int z = blockIdx.z;
  __shared__ double X[NDB][NDB + 1], Ip[NDB][NDB + 1], W1[NDB][NDB + 1], W2[NDB][NDB + 1];
  __shared__ int perm[NDB];
  const int j = threadIdx.x & 31, i = threadIdx.x >> 5, q = blockIdx.x;
  if (blockIdx.y == 1 && q == b) return;
  double *A = Fl + (size_t)z * g * g;
  W1[i][j] = A[(size_t)(b * NDB + i) * g + b * NDB + j];
  __syncthreads();
  nd_invert32(W1, W2, Ip, perm, i, j);
  if (blockIdx.y == 0) {
    if (q == b) { ipp[(size_t)z * NDB * NDB + i * NDB + j] = Ip[i][j]; return; }
    double *T = A + (size_t)(b * NDB) * g + q * NDB;
    X[i][j] = T[(size_t)i * g + j];
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NDB; k++) sum += Ip[i][k] * X[k][j];
    T[(size_t)i * g + j] = sum;
  } else {
    double *T = A + (size_t)(q * NDB) * g + b * NDB;
    X[i][j] = T[(size_t)i * g + j];
    colbuf[((size_t)z * g + q * NDB + i) * NDB + j] = X[i][j];
    __syncthreads();
    double sum = 0.0;
#pragma unroll 8
    for (int k = 0; k < NDB; k++) sum += X[i][k] * Ip[k][j];
    T[(size_t)i * g + j] = -sum;
  }
}

// trailing update  A_IJ <- A_IJ - A_IP(old) A_PJ(new)  for I, J != P; 64 x 64 tile per CTA, 4 x 4 micro-tile
__global__ void __launch_bounds__(256)
k_nd_update(int g, int b, double *__restrict__ Fl, const double *__restrict__ colbuf, const double *__restrict__ ipp) {
  const int z = blockIdx.z;
  __shared__ double Cb[NDT][NDB + 1], R[NDB][NDT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * NDT, j0 = blockIdx.x * NDT;
  double *A = Fl + (size_t)z * g * g;
  for (int q = threadIdx.x; q < NDT * NDB; q += 256) {
    const int ii = q / NDB, kk = q % NDB;
    Cb[ii][kk] = ((i0 + ii) / NDB == b) ? 0.0 : colbuf[((size_t)z * g + i0 + ii) * NDB + kk];
    const int k2 = q / NDT, jj = q % NDT;
    R[k2][jj] = A[(size_t)(b * NDB + k2) * g + j0 + jj];
  }
  __syncthreads();
  if ((int)blockIdx.y == (b * NDB) / NDT && (int)blockIdx.x == (b * NDB) / NDT)
    for (int q = threadIdx.x; q < NDB * NDB; q += 256)
      A[(size_t)(b * NDB + q / NDB) * g + b * NDB + q % NDB] = ipp[(size_t)z * NDB * NDB + q];
  double acc[4][4] = {};
#pragma unroll
  for (int kk = 0; kk < NDB; kk++) {
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
    if (i / NDB == b) continue;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int j = j0 + tx + 16 * q;
      if (j / NDB == b) continue;
      A[(size_t)i * g + j] -= acc[p][q];
    }
  }
}

// extend-add of the Schur complements of the fronts in child slot `want` into their parents' fronts
__global__ void k_nd_extend(int g, int p, int first, const double *__restrict__ Fl, int gp, double *__restrict__ Fp,
                            const int *__restrict__ nb, const int *__restrict__ parent, const int *__restrict__ slot,
                            const int *__restrict__ up_off, const int *__restrict__ upmap, int want) {
  const int z = blockIdx.z, f = first + z;
  if (slot[f] != want || parent[f] < 0) return;
  const int n = nb[f];
  const int r = blockIdx.y * 16 + threadIdx.y, c = blockIdx.x * 16 + threadIdx.x;
  if (r >= n || c >= n) return;
  const int *up = upmap + up_off[f];
  Fp[(size_t)parent[f] * gp * gp + (size_t)up[r] * gp + up[c]] += Fl[(size_t)z * g * g + (size_t)(p + r) * g + p + c];
}

// ------------------------------------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------------------------------------
// w_s = scaled right-hand side of the front's own unknowns, everything else zero.  dself != nullptr: the right-hand
// side is first multiplied by the 2x2 diagonal blocks of A (the solver then applies (B A)^-1, B = block-Jacobi scaling)
__global__ void k_nd_rhs(int g, int first, const int *__restrict__ ns, const int *__restrict__ sep_off,
                         const int *__restrict__ sepdof, const double *__restrict__ scale, const double *__restrict__ b,
                         const double *__restrict__ dself, const double *__restrict__ dpair, double *__restrict__ Wl) {
  const int z = blockIdx.y, f = first + z, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= g) return;
  double v = 0.0;
  if (r < ns[f]) {
    const int d = sepdof[sep_off[f] + r];
    v = scale[d] * (dself ? dself[d] * b[d] + dpair[d] * b[d ^ 1] : b[d]);
  }
  Wl[(size_t)z * g + r] = v;
}

// y = F[:, 0:ns] w_s (one warp per row):  own rows -> z,  boundary rows -> w_b += y
__global__ void __launch_bounds__(256)
k_nd_fwd(int g, int p, int first, const int *__restrict__ ns, const int *__restrict__ nb, const double *__restrict__ Fl,
         double *__restrict__ Wl, double *__restrict__ Zl) {
  const int z = blockIdx.z, f = first + z;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n = ns[f];
  if (row >= p + nb[f] || (row >= n && row < p)) return;
  const double *Fr = Fl + (size_t)z * g * g + (size_t)row * g, *w = Wl + (size_t)z * g;
  double sum = 0.0;
  for (int j = lane; j < n; j += 32) sum += Fr[j] * w[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    if (row < p) Zl[(size_t)z * g + row] = sum;
    else Wl[(size_t)z * g + row] += sum;
  }
}

// child -> parent: parent.w[up[r]] += child.w_b[r] (slot `want`), or parent -> child: child.w_b[r] = parent.w[up[r]]
__global__ void k_nd_vec_updown(int g, int p, int first, int gp, double *__restrict__ Wl, double *__restrict__ Wp,
                                const int *__restrict__ nb, const int *__restrict__ parent, const int *__restrict__ slot,
                                const int *__restrict__ up_off, const int *__restrict__ upmap, int want, int down) {
  const int z = blockIdx.y, f = first + z, r = blockIdx.x * blockDim.x + threadIdx.x;
  if (parent[f] < 0 || r >= nb[f]) return;
  if (!down && slot[f] != want) return;
  const size_t ip = (size_t)parent[f] * gp + upmap[up_off[f] + r], ic = (size_t)z * g + p + r;
  if (down) Wl[ic] = Wp[ip];
  else Wp[ip] += Wl[ic];
}

// x_s = z - F12' x_b (one warp per own row); written to w_s (for the children) and, unscaled, to the global solution
__global__ void __launch_bounds__(256)
k_nd_bwd(int g, int p, int first, const int *__restrict__ ns, const int *__restrict__ nb, const int *__restrict__ sep_off,
         const int *__restrict__ sepdof, const double *__restrict__ scale, const double *__restrict__ Fl,
         double *__restrict__ Wl, const double *__restrict__ Zl, double *__restrict__ x, int accumulate) {
  const int z = blockIdx.z, f = first + z;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= ns[f]) return;
  const int n = nb[f];
  const double *Fr = Fl + (size_t)z * g * g + (size_t)row * g + p, *w = Wl + (size_t)z * g + p;
  double sum = 0.0;
  for (int j = lane; j < n; j += 32) sum += Fr[j] * w[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    const double xs = Zl[(size_t)z * g + row] - sum;
    const int d = sepdof[sep_off[f] + row];
    Wl[(size_t)z * g + row] = xs;
    if (accumulate) x[d] += scale[d] * xs; else x[d] = scale[d] * xs;
  }
}

// r = b - A x, unscaled CSR, one thread per row
__global__ void k_nd_residual(int N, const int *__restrict__ ptr, const int *__restrict__ ind, const double *__restrict__ val,
                              const double *__restrict__ b, const double *__restrict__ x, double *__restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = b[i];
  for (int k = ptr[i]; k < ptr[i + 1]; k++) s -= val[k] * x[ind[k]];
  r[i] = s;
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
template <class T> static int nd_upload(T **d, const std::vector<T> &h) {
  UFE_CUDA(cudaMalloc((void **)d, std::max<size_t>(1, h.size()) * sizeof(T)));
  if (!h.empty()) UFE_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return UFE_OK;
}

extern "C" void ufe_nd_solver_free(ufe_nd_solver *S) {
  if (!S) return;
  void *p[] = {S->ns, S->nb, S->parent, S->slot, S->sep_off, S->up_off, S->sepdof, S->upmap, S->dst, S->ptr, S->ind, S->val,
               S->scale, S->dself, S->dpair, S->F, S->W, S->Z, S->ipp, S->colbuf, S->b, S->x, S->r, S->dx};
  for (void *q : p) if (q) cudaFree(q);
  if (S->e0) cudaEventDestroy(S->e0);
  if (S->e1) cudaEventDestroy(S->e1);
  if (S->st) cudaStreamDestroy(S->st);
  delete S;
}

// ptr / ind: scalar CSR pattern of A, 0-based, N = 2 nT rows; its block pattern must be the one that was analysed.
extern "C" int ufe_nd_solver_create(const ufe_nd_tree *T, int32_t N, const int32_t *ptr, const int32_t *ind, ufe_nd_solver **out) {
  if (!T || !ptr || !ind || !out || N != 2 * T->nT) { ufe_set_error("ufe_nd_solver_create: bad argument"); return UFE_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    ufe_set_error("ufe_nd_solver_create: no CUDA device (there is no CPU fallback)"); return UFE_ERR_CUDA;
  }
  const int nn = (int)T->nodes.size(), nl = T->n_levels;
  ufe_nd_solver *S = new ufe_nd_solver();
  S->nT = T->nT; S->N = N; S->nnz = ptr[N]; S->n_fronts = nn;
  S->lev.resize(nl);
  // level-major front numbering
  std::vector<int> zof(nn), fof(nn);
  for (int q = 0; q < nn; q++) {
    const NdNode &nd = T->nodes[q];
    NdLevelDev &L = S->lev[nd.level];
    zof[q] = L.n++;
    L.p = std::max(L.p, 2 * (int)nd.sep.size());
    L.max_nb = std::max(L.max_nb, 2 * (int)nd.bnd.size());
  }
  int first = 0;
  for (NdLevelDev &L : S->lev) {
    L.p = std::max(NDB, (L.p + NDB - 1) / NDB * NDB);
    L.g = (L.p + L.max_nb + NDT - 1) / NDT * NDT;
    L.first = first; first += L.n;
    L.f_off = S->f_doubles; S->f_doubles += (size_t)L.n * L.g * L.g;
    L.w_off = S->w_doubles; S->w_doubles += (size_t)L.n * L.g;
    S->flops += 2.0 * (double)L.n * (double)L.g * (double)L.g * (double)L.p;
  }
  for (const NdLevelDev &L : S->lev)
    if (L.n > 65535) { ufe_set_error("ufe_nd_solver_create: %d fronts in one level (limit 65535): use larger leaves", L.n); delete S; return UFE_ERR_INVALID; }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if ((double)S->f_doubles * 8.0 > 0.9 * (double)free_b) {
    ufe_set_error("ufe_nd_solver_create: the padded fronts need %.1f GB, %.1f GB are free", S->f_doubles * 8e-9, free_b * 1e-9);
    delete S; return UFE_ERR_INVALID;
  }
  for (int q = 0; q < nn; q++) fof[q] = S->lev[T->nodes[q].level].first + zof[q];
  std::vector<int> ns(nn), nb(nn), parent(nn), slot(nn), sep_off(nn), up_off(nn), sepdof, upmap;
  {
    std::vector<int> so(nn + 1, 0), uo(nn + 1, 0);
    for (int q = 0; q < nn; q++) {
      const NdNode &nd = T->nodes[q];
      const int f = fof[q];
      ns[f] = 2 * (int)nd.sep.size(); nb[f] = nd.parent >= 0 ? 2 * (int)nd.bnd.size() : 0;
      parent[f] = nd.parent >= 0 ? zof[nd.parent] : -1;
      slot[f] = nd.parent >= 0 && T->nodes[nd.parent].child[1] == q ? 1 : 0;
    }
    for (int f = 0; f < nn; f++) { so[f + 1] = so[f] + ns[f]; uo[f + 1] = uo[f] + nb[f]; }
    sepdof.resize(so[nn]); upmap.resize(uo[nn]);
    for (int q = 0; q < nn; q++) {
      const NdNode &nd = T->nodes[q];
      const int f = fof[q];
      sep_off[f] = so[f]; up_off[f] = uo[f];
      for (size_t k = 0; k < nd.sep.size(); k++) { sepdof[so[f] + 2 * k] = 2 * nd.sep[k]; sepdof[so[f] + 2 * k + 1] = 2 * nd.sep[k] + 1; }
      if (nd.parent < 0) continue;
      const NdNode &pa = T->nodes[nd.parent];
      const int pns = (int)pa.sep.size(), pp = S->lev[pa.level].p;
      for (size_t k = 0; k < nd.bnd.size(); k++) {
        const int u = nd.up[k], row = u < pns ? 2 * u : pp + 2 * (u - pns);
        upmap[uo[f] + 2 * k] = row; upmap[uo[f] + 2 * k + 1] = row + 1;
      }
    }
  }
  // destination of every scalar entry: its block entry's front and block position
  std::vector<long long> dst(S->nnz);
  for (int i = 0; i < N; i++) {
    const int bi = i >> 1;
    for (int k = ptr[i]; k < ptr[i + 1]; k++) {
      const int j = ind[k], bj = j >> 1;
      int e = -1;
      if (j >= 0 && j < N) for (int t = T->bptr[bi]; t < T->bptr[bi + 1]; t++) if (T->bind[t] == bj) { e = t; break; }
      if (e < 0) { ufe_set_error("ufe_nd_solver_create: entry (%d,%d) is not in the analysed block pattern", i, j); delete S; return UFE_ERR_INVALID; }
      const int q = T->entry_node[e];
      const NdNode &nd = T->nodes[q];
      const NdLevelDev &L = S->lev[nd.level];
      const int nsb = (int)nd.sep.size();
      const int br = T->entry_row[e], bc = T->entry_col[e];
      const int row = (br < nsb ? 2 * br : L.p + 2 * (br - nsb)) + (i & 1), col = (bc < nsb ? 2 * bc : L.p + 2 * (bc - nsb)) + (j & 1);
      dst[k] = (long long)(L.f_off + (size_t)zof[q] * L.g * L.g + (size_t)row * L.g + col);
    }
  }
  int rc = UFE_OK;
  auto fail = [&](int c) { ufe_nd_solver_free(S); return c; };
  if ((rc = nd_upload(&S->ns, ns)) || (rc = nd_upload(&S->nb, nb)) || (rc = nd_upload(&S->parent, parent)) ||
      (rc = nd_upload(&S->slot, slot)) || (rc = nd_upload(&S->sep_off, sep_off)) || (rc = nd_upload(&S->up_off, up_off)) ||
      (rc = nd_upload(&S->sepdof, sepdof)) || (rc = nd_upload(&S->upmap, upmap)) || (rc = nd_upload(&S->dst, dst)))
    return fail(rc);
  {
    std::vector<int> hp(ptr, ptr + N + 1), hi(ind, ind + S->nnz);
    if ((rc = nd_upload(&S->ptr, hp)) || (rc = nd_upload(&S->ind, hi))) return fail(rc);
  }
  size_t ipp_n = 0, col_n = 0;
  for (const NdLevelDev &L : S->lev) { ipp_n = std::max(ipp_n, (size_t)L.n * NDB * NDB); col_n = std::max(col_n, (size_t)L.n * L.g * NDB); }
  const struct { double **p; size_t n; } bufs[] = {{&S->val, (size_t)S->nnz}, {&S->scale, (size_t)N}, {&S->dself, (size_t)N}, {&S->dpair, (size_t)N}, {&S->F, S->f_doubles},
      {&S->W, S->w_doubles}, {&S->Z, S->w_doubles}, {&S->ipp, ipp_n}, {&S->colbuf, col_n}, {&S->b, (size_t)N}, {&S->x, (size_t)N},
      {&S->r, (size_t)N}, {&S->dx, (size_t)N}};
  for (const auto &bf : bufs)
    if (cudaMalloc((void **)bf.p, std::max<size_t>(1, bf.n) * sizeof(double)) != cudaSuccess) {
      ufe_set_error("ufe_nd_solver_create: out of device memory (%zu doubles)", bf.n); cudaGetLastError(); return fail(UFE_ERR_CUDA);
    }
  if (cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&S->e0) != cudaSuccess ||
      cudaEventCreate(&S->e1) != cudaSuccess) { ufe_set_error("ufe_nd_solver_create: stream / event creation failed"); return fail(UFE_ERR_CUDA); }
  UFE_CUDA(cudaDeviceSynchronize());
  *out = S;
  return UFE_OK;
}

// factorisation from values already on the device (same order as the pattern given to create)
static int nd_factor_device(ufe_nd_solver *S, cudaStream_t st, const double *dval) {
  const int N = S->N, tb = 256;
  k_nd_scale<<<(N + tb - 1) / tb, tb, 0, st>>>(N, S->ptr, S->ind, dval, S->scale, S->dself, S->dpair); UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaMemsetAsync(S->F, 0, S->f_doubles * sizeof(double), st));
  k_nd_assemble<<<(N + tb - 1) / tb, tb, 0, st>>>(N, S->ptr, S->ind, dval, S->scale, S->dst, S->F); UFE_LAUNCH_CHECK();
  for (int l = (int)S->lev.size() - 1; l >= 0; l--) {
    const NdLevelDev &L = S->lev[l];
    double *Fl = S->F + L.f_off;
    k_nd_pad_identity<<<dim3((L.p + 127) / 128, L.n), 128, 0, st>>>(L.g, L.p, L.first, S->ns, Fl); UFE_LAUNCH_CHECK();
    for (int b = 0; b < L.p / NDB; b++) {
      k_nd_panel<<<dim3(L.g / NDB, 2, L.n), 1024, 0, st>>>(L.g, b, Fl, S->ipp, S->colbuf); UFE_LAUNCH_CHECK();
      k_nd_update<<<dim3(L.g / NDT, L.g / NDT, L.n), 256, 0, st>>>(L.g, b, Fl, S->colbuf, S->ipp); UFE_LAUNCH_CHECK();
    }
    if (l > 0 && L.max_nb > 0) {
      const NdLevelDev &P = S->lev[l - 1];
      const int t = (L.max_nb + 15) / 16;
      for (int want = 0; want < 2; want++) {
        k_nd_extend<<<dim3(t, t, L.n), dim3(16, 16), 0, st>>>(L.g, L.p, L.first, Fl, P.g, S->F + P.f_off, S->nb, S->parent, S->slot,
                                                              S->up_off, S->upmap, want);
        UFE_LAUNCH_CHECK();
      }
    }
  }
  S->factored = true;
  return UFE_OK;
}

// x (+)= A^-1 r with the factors; r, x device vectors of length N
static int nd_apply_device(ufe_nd_solver *S, cudaStream_t st, const double *r, double *x, int accumulate, int premul) {
  const int nl = (int)S->lev.size();
  for (int l = 0; l < nl; l++) {
    const NdLevelDev &L = S->lev[l];
    k_nd_rhs<<<dim3((L.g + 255) / 256, L.n), 256, 0, st>>>(L.g, L.first, S->ns, S->sep_off, S->sepdof, S->scale, r, premul ? S->dself : nullptr, S->dpair, S->W + L.w_off);
    UFE_LAUNCH_CHECK();
  }
  for (int l = nl - 1; l >= 0; l--) {
    const NdLevelDev &L = S->lev[l];
    k_nd_fwd<<<dim3(L.g / 8, 1, L.n), 256, 0, st>>>(L.g, L.p, L.first, S->ns, S->nb, S->F + L.f_off, S->W + L.w_off, S->Z + L.w_off);
    UFE_LAUNCH_CHECK();
    if (l > 0 && L.max_nb > 0) {
      const NdLevelDev &P = S->lev[l - 1];
      for (int want = 0; want < 2; want++) {
        k_nd_vec_updown<<<dim3((L.max_nb + 127) / 128, L.n), 128, 0, st>>>(L.g, L.p, L.first, P.g, S->W + L.w_off, S->W + P.w_off, S->nb,
                                                                           S->parent, S->slot, S->up_off, S->upmap, want, 0);
        UFE_LAUNCH_CHECK();
      }
    }
  }
  for (int l = 0; l < nl; l++) {
    const NdLevelDev &L = S->lev[l];
    if (l > 0 && L.max_nb > 0) {
      const NdLevelDev &P = S->lev[l - 1];
      k_nd_vec_updown<<<dim3((L.max_nb + 127) / 128, L.n), 128, 0, st>>>(L.g, L.p, L.first, P.g, S->W + L.w_off, S->W + P.w_off, S->nb,
                                                                         S->parent, S->slot, S->up_off, S->upmap, 0, 1);
      UFE_LAUNCH_CHECK();
    }
    k_nd_bwd<<<dim3(L.p / 8, 1, L.n), 256, 0, st>>>(L.g, L.p, L.first, S->ns, S->nb, S->sep_off, S->sepdof, S->scale, S->F + L.f_off,
                                                    S->W + L.w_off, S->Z + L.w_off, x, accumulate);
    UFE_LAUNCH_CHECK();
  }
  return UFE_OK;
}

// val: host values of A in the order of the pattern given to ufe_nd_solver_create
extern "C" int ufe_nd_solver_factor(ufe_nd_solver *S, const double *val) {
  if (!S || !val) { ufe_set_error("ufe_nd_solver_factor: bad argument"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaMemcpyAsync(S->val, val, (size_t)S->nnz * sizeof(double), cudaMemcpyHostToDevice, S->st));
  UFE_CUDA(cudaEventRecord(S->e0, S->st));
  UFE_TRY(nd_factor_device(S, S->st, S->val));
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
  UFE_TRY(nd_apply_device(S, S->st, S->b, S->x, 0, 0));
  for (int it = 0; it < n_refine; it++) {
    k_nd_residual<<<(N + tb - 1) / tb, tb, 0, S->st>>>(N, S->ptr, S->ind, S->val, S->b, S->x, S->r); UFE_LAUNCH_CHECK();
    UFE_TRY(nd_apply_device(S, S->st, S->r, S->x, 1, 0));
  }
  UFE_CUDA(cudaEventRecord(S->e1, S->st));
  k_nd_residual<<<(N + tb - 1) / tb, tb, 0, S->st>>>(N, S->ptr, S->ind, S->val, S->b, S->x, S->r); UFE_LAUNCH_CHECK();
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

// timings of the last factor / solve call (CUDA events on the solver's stream), storage and flop count of the sweep
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
// ------------------------------------------------------------------------------------------------------------------
int ufe_nd_pc_create(cudaStream_t st, const DevSystem &S, int nT, const double *gcx, const double *gcy, int leaf, ufe_nd_solver **out) {
  *out = nullptr;
  if (S.m_loc != S.N || S.r1 != 1 || S.N != 2 * nT) {
    ufe_set_error("nd_lu preconditioner: the rows of the system must not be partitioned (one GPU)"); return UFE_ERR_INVALID;
  }
  std::vector<int> ptr(S.N + 1), ind(S.nnz);
  UFE_CUDA(cudaStreamSynchronize(st));
  UFE_CUDA(cudaMemcpy(ptr.data(), S.ptr, ptr.size() * sizeof(int), cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(ind.data(), S.ind, ind.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for (int &v : ptr) v -= 1;
  for (int &v : ind) v -= 1;
  // block pattern over triangles
  std::vector<int> bptr(nT + 1, 0), bind, row;
  for (int t = 0; t < nT; t++) {
    row.clear();
    for (int k = ptr[2 * t]; k < ptr[2 * t + 2]; k++) row.push_back(ind[k] >> 1);
    std::sort(row.begin(), row.end());
    row.erase(std::unique(row.begin(), row.end()), row.end());
    bind.insert(bind.end(), row.begin(), row.end());
    bptr[t + 1] = (int)bind.size();
  }
  ufe_nd_tree *T = nullptr;
  UFE_TRY(ufe_nd_analyse(nT, gcx, gcy, bptr.data(), bind.data(), leaf, &T));
  const int rc = ufe_nd_solver_create(T, S.N, ptr.data(), ind.data(), out);
  ufe_nd_tree_free(T);
  return rc;
}

int ufe_nd_pc_factor(cudaStream_t st, ufe_nd_solver *S, const double *dval) { return nd_factor_device(S, st, dval); }

int ufe_nd_pc_apply(cudaStream_t st, ufe_nd_solver *S, const double *r, double *z) { return nd_apply_device(S, st, r, z, 0, 1); }
