// Operator construction on the GPU: one thread per matrix row.
// Replaces calc_matrix_operators_mesh_a_b / _b_a / _b_b_2nd_order
// (src/UPSY/mesh/discretisation/mesh_disc_calc_matrix_operators_2D.f90:198,337,612),
// extend_group_single_iteration_a/_b (src/UPSY/mesh/mesh_utilities.f90:1856,1896),
// calc_shape_functions_2D_stag_1st_order / _reg_2nd_order
// (src/UPSY/basic/math_utilities/shape_functions.f90:366,218) and the closed-form 3x3 /
// 5x5 inverses (matrix_algebra.f90:133,185).
//
// The reference grows the neighbourhood with a global `map` array per rank and a serial
// loop over rows; here each row keeps a private stack (membership = linear search in a
// stack of at most UFE_STACK_MAX entries), which visits neighbours in the same order and
// therefore yields the same column order.  Two passes: count -> exclusive scan -> fill.
// Compiled with -fmad=false so that +,-,*,/ round exactly as the CPU evaluation order.
#include "ufe_internal.cuh"
#include <cub/device/device_scan.cuh>
#include <float.h>

struct MeshView {
  int nV, nTri, nC_mem;
  const double *V, *TriGC;
  const int *Tri, *TriC, *C, *nC, *iTri, *niTri;
};

__constant__ signed char c_perm5[120][5];
__constant__ signed char c_sign5[120];
__constant__ signed char c_perm4[24][4];
__constant__ signed char c_sign4[24];

// Leibniz tables: permutations sorted by (sigma(n), ..., sigma(1)) ascending, i.e. the
// term order of the expanded determinant / cofactors in matrix_algebra.f90:185-457.
template <int N>
static void make_perm_table(signed char (*perm)[N], signed char *sign) {
  int idx[N];
  int count = 0;
  // odometer over (sigma(N) slowest ... sigma(1) fastest)
  int total = 1;
  for (int i = 0; i < N; i++) total *= N;
  for (int code = 0; code < total; code++) {
    int c = code;
    for (int pos = 0; pos < N; pos++) { idx[pos] = c % N; c /= N; }   // idx[0] = sigma(1) fastest
    bool ok = true;
    for (int a = 0; a < N && ok; a++)
      for (int b = a + 1; b < N; b++) if (idx[a] == idx[b]) { ok = false; break; }
    if (!ok) continue;
    int inv = 0;
    for (int a = 0; a < N; a++) for (int b = a + 1; b < N; b++) if (idx[a] > idx[b]) inv++;
    for (int a = 0; a < N; a++) perm[count][a] = (signed char)idx[a];
    sign[count] = (inv & 1) ? -1 : 1;
    count++;
  }
}

int ufe_operators_init_tables() {
  static bool done = false;
  static signed char p5[120][5], s5[120], p4[24][4], s4[24];
  if (!done) { make_perm_table<5>(p5, s5); make_perm_table<4>(p4, s4); done = true; }
  UFE_CUDA(cudaMemcpyToSymbol(c_perm5, p5, sizeof p5));
  UFE_CUDA(cudaMemcpyToSymbol(c_sign5, s5, sizeof s5));
  UFE_CUDA(cudaMemcpyToSymbol(c_perm4, p4, sizeof p4));
  UFE_CUDA(cudaMemcpyToSymbol(c_sign4, s4, sizeof s4));
  return UFE_OK;
}

// gfortran NORM2 of a 2-vector (scaled sum of squares)
__device__ __forceinline__ double norm2_2(double a, double b) {
  double scale = 1.0, ssq = 0.0;
  if (a != 0.0) {
    const double ax = fabs(a);
    if (scale < ax) { const double t = scale / ax; ssq = 1.0 + ssq * t * t; scale = ax; }
    else { const double t = ax / scale; ssq += t * t; }
  }
  if (b != 0.0) {
    const double ax = fabs(b);
    if (scale < ax) { const double t = scale / ax; ssq = 1.0 + ssq * t * t; scale = ax; }
    else { const double t = ax / scale; ssq += t * t; }
  }
  return scale * sqrt(ssq);
}

__device__ double det5(const double (*A)[5]) {
  double acc = 0.0;
  for (int t = 0; t < 120; t++) {
    const double prod = A[0][c_perm5[t][0]] * A[1][c_perm5[t][1]] * A[2][c_perm5[t][2]] *
                        A[3][c_perm5[t][3]] * A[4][c_perm5[t][4]];
    if (t == 0) acc = c_sign5[t] > 0 ? prod : -prod;
    else acc = c_sign5[t] > 0 ? acc + prod : acc - prod;
  }
  return acc;
}

__device__ bool inv5(const double (*A)[5], double (*M)[5]) {
  const double det = det5(A);
  if (fabs(det) < DBL_MIN) return false;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) {
      int rows[4], cols[4], nr = 0, nc = 0;
      for (int r = 0; r < 5; r++) if (r != i) rows[nr++] = r;
      for (int c = 0; c < 5; c++) if (c != j) cols[nc++] = c;
      const int sij = ((i + j) & 1) ? -1 : 1;
      double acc = 0.0;
      for (int t = 0; t < 24; t++) {
        const double prod = A[rows[0]][cols[c_perm4[t][0]]] * A[rows[1]][cols[c_perm4[t][1]]] *
                            A[rows[2]][cols[c_perm4[t][2]]] * A[rows[3]][cols[c_perm4[t][3]]];
        const int s = sij * c_sign4[t];
        if (t == 0) acc = s > 0 ? prod : -prod;
        else acc = s > 0 ? acc + prod : acc - prod;
      }
      M[j][i] = acc / det;      // transpose(COFACTOR) / det
    }
  return true;
}

__device__ bool inv3(const double (*A)[3], double (*M)[3]) {
  double m[3][3];
  m[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  m[0][1] = A[1][0] * A[2][2] - A[1][2] * A[2][0];
  m[0][2] = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  m[1][0] = A[0][1] * A[2][2] - A[0][2] * A[2][1];
  m[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0];
  m[1][2] = A[0][0] * A[2][1] - A[0][1] * A[2][0];
  m[2][0] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
  m[2][1] = A[0][0] * A[1][2] - A[0][2] * A[1][0];
  m[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const double det = A[0][0] * m[0][0] - A[0][1] * m[0][1] + A[0][2] * m[0][2];
  if (fabs(det) <= DBL_MIN) return false;     // shape_functions.f90:421 uses "<="
  m[0][1] = -m[0][1]; m[1][0] = -m[1][0]; m[1][2] = -m[1][2]; m[2][1] = -m[2][1];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[i][j] = m[j][i] / det;
  return true;
}

__device__ __forceinline__ bool in_stack(const int *stack, int n, int id) {
  for (int i = 0; i < n; i++) if (stack[i] == id) return true;
  return false;
}

// one BFS sweep over vertices (a-grid) / triangles (b-grid); returns false on overflow
__device__ bool extend_a(const MeshView &m, int *stack, int &n) {
  const int n0 = n;
  for (int i = 0; i < n0; i++) {
    const int vi = stack[i];
    const int nc = m.nC[vi - 1];
    for (int ci = 0; ci < nc; ci++) {
      const int vj = m.C[(size_t)ci * m.nV + vi - 1];
      if (!in_stack(stack, n, vj)) { if (n >= UFE_STACK_MAX) return false; stack[n++] = vj; }
    }
  }
  return true;
}
__device__ bool extend_b(const MeshView &m, int *stack, int &n) {
  const int n0 = n;
  for (int i = 0; i < n0; i++) {
    const int ti = stack[i];
    for (int k = 0; k < 3; k++) {
      const int tj = m.TriC[(size_t)k * m.nTri + ti - 1];
      if (tj == 0) continue;
      if (!in_stack(stack, n, tj)) { if (n >= UFE_STACK_MAX) return false; stack[n++] = tj; }
    }
  }
  return true;
}

// staggered 1st order: returns false if singular
__device__ bool shape_stag_1st(double x, double y, int n_c, const double *xc, const double *yc, double *Nf,
                               double *Nfx, double *Nfy) {
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, M[3][3];
  for (int i = 0; i < n_c; i++) {
    const double dx = xc[i] - x, dy = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(dx, dy), 1.5);
    const double w2 = w * w;
    A[0][0] += w2 * 1.0 * 1.0; A[0][1] += w2 * 1.0 * dx; A[0][2] += w2 * 1.0 * dy;
    A[1][0] += w2 * dx * 1.0;  A[1][1] += w2 * dx * dx;  A[1][2] += w2 * dx * dy;
    A[2][0] += w2 * dy * 1.0;  A[2][1] += w2 * dy * dx;  A[2][2] += w2 * dy * dy;
  }
  if (!inv3(A, M)) return false;
  for (int i = 0; i < n_c; i++) {
    const double dx = xc[i] - x, dy = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(dx, dy), 1.5);
    const double w2 = w * w;
    Nf[i] = w2 * ((M[0][0] * 1.0) + (M[0][1] * dx) + (M[0][2] * dy));
    Nfx[i] = w2 * ((M[1][0] * 1.0) + (M[1][1] * dx) + (M[1][2] * dy));
    Nfy[i] = w2 * ((M[2][0] * 1.0) + (M[2][1] * dx) + (M[2][2] * dy));
  }
  return true;
}

// regular 2nd order: N[k][i], Ni[k]; returns false if singular
__device__ bool shape_reg_2nd(double x, double y, int n_c, const double *xc, const double *yc, double *Ni,
                              double (*N)[UFE_STACK_MAX]) {
  double A[5][5], M[5][5];
  for (int a = 0; a < 5; a++) for (int b = 0; b < 5; b++) A[a][b] = 0.0;
  for (int i = 0; i < n_c; i++) {
    const double X = xc[i] - x, Y = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(X, Y), 1.5);
    const double w2 = w * w, X2 = X * X, Y2 = Y * Y;
    double r[5];
    r[0] = w2 * X; r[1] = w2 * Y; r[2] = w2 * 1.0 / 2.0 * X2; r[3] = w2 * X * Y; r[4] = w2 * 1.0 / 2.0 * Y2;
    for (int k = 0; k < 5; k++) {
      A[k][0] += r[k] * X;
      A[k][1] += r[k] * Y;
      A[k][2] += r[k] * 1.0 / 2.0 * X2;
      A[k][3] += r[k] * X * Y;
      A[k][4] += r[k] * 1.0 / 2.0 * Y2;
    }
  }
  if (!inv5(A, M)) return false;
  double s[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n_c; i++) {
    const double X = xc[i] - x, Y = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(X, Y), 1.5);
    const double w2 = w * w, X2 = X * X, Y2 = Y * Y;
    for (int k = 0; k < 5; k++)
      N[k][i] = w2 * ((M[k][0] * X) + (M[k][1] * Y) + (M[k][2] * 1.0 / 2.0 * X2) + (M[k][3] * X * Y) +
                      (M[k][4] * 1.0 / 2.0 * Y2));
  }
  for (int k = 0; k < 5; k++) { for (int i = 0; i < n_c; i++) s[k] += N[k][i]; Ni[k] = -s[k]; }
  return true;
}

// regular 1st order (shape_functions.f90:140-216, 2x2 inverse matrix_algebra.f90:11-19, 110-131): Nfx[i], Nfy[i] of the
// neighbours and the home point's Nfx_i, Nfy_i; returns false if singular
__device__ bool shape_reg_1st(double x, double y, int n_c, const double *xc, const double *yc, double *Ni, double *Nfx,
                              double *Nfy) {
  double A00 = 0.0, A01 = 0.0, A10 = 0.0, A11 = 0.0;
  for (int i = 0; i < n_c; i++) {
    const double dx = xc[i] - x, dy = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(dx, dy), 1.5);
    const double w2 = w * w;
    A00 += w2 * dx * dx; A01 += w2 * dx * dy; A10 += w2 * dy * dx; A11 += w2 * dy * dy;
  }
  const double det = A00 * A11 - A01 * A10;
  if (fabs(det) < DBL_MIN) return false;
  const double M00 = A11 / det, M01 = -A01 / det, M10 = -A10 / det, M11 = A00 / det;
  double sx = 0.0, sy = 0.0;
  for (int i = 0; i < n_c; i++) {
    const double dx = xc[i] - x, dy = yc[i] - y;
    const double w = 1.0 / pow(norm2_2(dx, dy), 1.5);
    const double w2 = w * w;
    Nfx[i] = w2 * ((M00 * dx) + (M01 * dy));
    Nfy[i] = w2 * ((M10 * dx) + (M11 * dy));
    sx += Nfx[i]; sy += Nfy[i];
  }
  Ni[0] = -sx; Ni[1] = -sy;
  return true;
}

// FAMILY 0: a_b (rows = triangles, cols = vertices), 1: b_a (rows = vertices, cols = triangles),
// 2: b_b 2nd order, 3: a_a (ddx, ddy; regular 1st order).  PASS 0: count, PASS 1: fill.
template <int FAMILY, int PASS>
__global__ void __launch_bounds__(128)
k_build_operator(MeshView m, int row1, int m_loc, int *__restrict__ counts, const int *__restrict__ ptr,
                 int *__restrict__ ind, double *v0, double *v1, double *v2, double *v3, double *v4,
                 int *err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_loc) return;
  const int row = row1 + r;      // 1-based global row id
  int stack[UFE_STACK_MAX];
  double xc[UFE_STACK_MAX], yc[UFE_STACK_MAX];
  int n = 0;
  double x, y;
  if (FAMILY == 0) {
    x = m.TriGC[row - 1]; y = m.TriGC[(size_t)m.nTri + row - 1];
    for (int k = 0; k < 3; k++) stack[n++] = m.Tri[(size_t)k * m.nTri + row - 1];
    while (n < 3) if (!extend_a(m, stack, n)) { atomicExch(err, 1); return; }
  } else if (FAMILY == 1) {
    x = m.V[row - 1]; y = m.V[(size_t)m.nV + row - 1];
    const int nt = m.niTri[row - 1];
    for (int k = 0; k < nt; k++) stack[n++] = m.iTri[(size_t)k * m.nV + row - 1];
    while (n < 3) if (!extend_b(m, stack, n)) { atomicExch(err, 1); return; }
  } else if (FAMILY == 3) {
    x = m.V[row - 1]; y = m.V[(size_t)m.nV + row - 1];
    stack[n++] = row;
    while (n - 1 < 2) if (!extend_a(m, stack, n)) { atomicExch(err, 1); return; }
  } else {
    x = m.TriGC[row - 1]; y = m.TriGC[(size_t)m.nTri + row - 1];
    stack[n++] = row;
    while (n - 1 < 5) if (!extend_b(m, stack, n)) { atomicExch(err, 1); return; }
  }
  if (FAMILY == 3) {
    double Nfx[UFE_STACK_MAX], Nfy[UFE_STACK_MAX], Ni[2];
    int nc;
    for (;;) {
      nc = 0;
      for (int i = 0; i < n; i++) {
        const int id = stack[i];
        if (id == row) continue;
        xc[nc] = m.V[id - 1]; yc[nc] = m.V[(size_t)m.nV + id - 1]; nc++;
      }
      if (shape_reg_1st(x, y, nc, xc, yc, Ni, Nfx, Nfy)) break;
      if (!extend_a(m, stack, n)) { atomicExch(err, 1); return; }
    }
    if (PASS == 0) { counts[r] = nc + 1; return; }
    int k = ptr[r] - 1;
    ind[k] = row; v0[k] = Ni[0]; v1[k] = Ni[1];
    k++;
    int c = 0;
    for (int i = 0; i < n; i++) {
      const int id = stack[i];
      if (id == row) continue;
      ind[k] = id; v0[k] = Nfx[c]; v1[k] = Nfy[c];
      k++; c++;
    }
  } else if (FAMILY < 2) {
    double Nf[UFE_STACK_MAX], Nfx[UFE_STACK_MAX], Nfy[UFE_STACK_MAX];
    for (;;) {
      for (int i = 0; i < n; i++) {
        const int id = stack[i];
        if (FAMILY == 0) { xc[i] = m.V[id - 1]; yc[i] = m.V[(size_t)m.nV + id - 1]; }
        else { xc[i] = m.TriGC[id - 1]; yc[i] = m.TriGC[(size_t)m.nTri + id - 1]; }
      }
      if (shape_stag_1st(x, y, n, xc, yc, Nf, Nfx, Nfy)) break;
      const bool ok = (FAMILY == 0) ? extend_a(m, stack, n) : extend_b(m, stack, n);
      if (!ok) { atomicExch(err, 1); return; }
    }
    if (PASS == 0) { counts[r] = n; return; }
    const int k0 = ptr[r] - 1;
    for (int i = 0; i < n; i++) { ind[k0 + i] = stack[i]; v0[k0 + i] = Nf[i]; v1[k0 + i] = Nfx[i]; v2[k0 + i] = Nfy[i]; }
  } else {
    double N[5][UFE_STACK_MAX], Ni[5];
    int nc;
    for (;;) {
      nc = 0;
      for (int i = 0; i < n; i++) {
        const int id = stack[i];
        if (id == row) continue;
        xc[nc] = m.TriGC[id - 1]; yc[nc] = m.TriGC[(size_t)m.nTri + id - 1]; nc++;
      }
      if (shape_reg_2nd(x, y, nc, xc, yc, Ni, N)) break;
      if (!extend_b(m, stack, n)) { atomicExch(err, 1); return; }
    }
    if (PASS == 0) { counts[r] = nc + 1; return; }
    int k = ptr[r] - 1;
    double *out[5] = {v0, v1, v2, v3, v4};
    ind[k] = row;
    for (int q = 0; q < 5; q++) out[q][k] = Ni[q];
    k++;
    int c = 0;
    for (int i = 0; i < n; i++) {
      const int id = stack[i];
      if (id == row) continue;
      ind[k] = id;
      for (int q = 0; q < 5; q++) out[q][k] = N[q][c];
      k++; c++;
    }
  }
}

__global__ void k_ptr_from_scan(int m_loc, const int *__restrict__ excl, const int *__restrict__ counts, int *ptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m_loc) ptr[r] = excl[r] + 1;
  if (r == m_loc - 1) ptr[m_loc] = excl[r] + counts[r] + 1;
}

// exclusive scan of counts -> 1-based ptr (m_loc+1); returns nnz through *nnz_out
int ufe_counts_to_ptr(cudaStream_t st, int m_loc, int *counts, int *ptr, int *nnz_out) {
  if (m_loc <= 0) { *nnz_out = 0; int one = 1; UFE_CUDA(cudaMemcpyAsync(ptr, &one, sizeof(int), cudaMemcpyHostToDevice, st)); UFE_CUDA(cudaStreamSynchronize(st)); return UFE_OK; }
  int *excl = nullptr; void *tmp = nullptr; size_t tmp_bytes = 0;
  UFE_CUDA(cudaMalloc(&excl, sizeof(int) * m_loc));
  UFE_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, excl, m_loc, st));
  UFE_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  UFE_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, excl, m_loc, st));
  g_launch_count++;
  k_ptr_from_scan<<<ufe_div_up(m_loc, 256), 256, 0, st>>>(m_loc, excl, counts, ptr);
  UFE_LAUNCH_CHECK();
  int last = 0;
  UFE_CUDA(cudaMemcpyAsync(&last, ptr + m_loc, sizeof(int), cudaMemcpyDeviceToHost, st));
  UFE_CUDA(cudaStreamSynchronize(st));
  *nnz_out = last - 1;
  cudaFree(excl); cudaFree(tmp);
  return UFE_OK;
}

template <int FAMILY>
static int build_family(cudaStream_t st, const MeshView &mv, int row1, int m_loc, int m, int n, DevFamily &F) {
  F.m_loc = m_loc; F.m = m; F.n = n; F.i1 = row1; F.nval = (FAMILY == 2) ? 5 : (FAMILY == 3 ? 2 : 3);
  int *counts = nullptr, *err = nullptr;
  UFE_CUDA(cudaMalloc(&counts, sizeof(int) * (m_loc > 0 ? m_loc : 1)));
  UFE_CUDA(cudaMalloc(&err, sizeof(int)));
  UFE_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  UFE_CUDA(cudaMalloc(&F.ptr, sizeof(int) * (m_loc + 1)));
  const int blocks = ufe_div_up(m_loc, 128);
  if (m_loc > 0) {
    k_build_operator<FAMILY, 0><<<blocks, 128, 0, st>>>(mv, row1, m_loc, counts, nullptr, nullptr, nullptr, nullptr,
                                                        nullptr, nullptr, nullptr, err);
    UFE_LAUNCH_CHECK();
  }
  UFE_TRY(ufe_counts_to_ptr(st, m_loc, counts, F.ptr, &F.nnz));
  int herr = 0;
  UFE_CUDA(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) { ufe_set_error("expanded local neighbourhood too far! (operator family %d)", FAMILY); return UFE_ERR_OPERATOR; }
  const size_t nz = F.nnz > 0 ? F.nnz : 1;
  UFE_CUDA(cudaMalloc(&F.ind, sizeof(int) * nz));
  for (int q = 0; q < F.nval; q++) UFE_CUDA(cudaMalloc(&F.val[q], sizeof(double) * nz));
  if (m_loc > 0) {
    k_build_operator<FAMILY, 1><<<blocks, 128, 0, st>>>(mv, row1, m_loc, counts, F.ptr, F.ind, F.val[0], F.val[1],
                                                        F.val[2], F.val[3], F.val[4], err);
    UFE_LAUNCH_CHECK();
  }
  UFE_CUDA(cudaStreamSynchronize(st));
  cudaFree(counts); cudaFree(err);
  return UFE_OK;
}

// M_ddx_a_a / M_ddy_a_a (calc_matrix_operators_mesh_a_a, mesh_disc_calc_matrix_operators_2D.f90:60-196), all rows
int ufe_build_operators_a_a(cudaStream_t st, const DevMesh &dm, DevFamily &F) {
  UFE_TRY(ufe_operators_init_tables());
  MeshView mv{dm.nV, dm.nTri, dm.nC_mem, dm.V, dm.TriGC, dm.Tri, dm.TriC, dm.C, dm.nC, dm.iTri, dm.niTri};
  return build_family<3>(st, mv, 1, dm.nV, dm.nV, dm.nV, F);
}

int ufe_build_operators(cudaStream_t st, const DevMesh &dm, int vi1, int vi2, int ti1, int ti2, bool need[3],
                        DevFamily fam[3]) {
  UFE_TRY(ufe_operators_init_tables());
  MeshView mv{dm.nV, dm.nTri, dm.nC_mem, dm.V, dm.TriGC, dm.Tri, dm.TriC, dm.C, dm.nC, dm.iTri, dm.niTri};
  if (need[0]) UFE_TRY(build_family<0>(st, mv, ti1, ti2 - ti1 + 1, dm.nTri, dm.nV, fam[0]));
  if (need[1]) UFE_TRY(build_family<1>(st, mv, vi1, vi2 - vi1 + 1, dm.nV, dm.nTri, fam[1]));
  if (need[2]) UFE_TRY(build_family<2>(st, mv, ti1, ti2 - ti1 + 1, dm.nTri, dm.nTri, fam[2]));
  return UFE_OK;
}
