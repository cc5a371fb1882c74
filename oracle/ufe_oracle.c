/*
 * ufe_oracle.c -- CPU restatement of the reference's DIVA/SSA velocity-solve path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it, and only as the checker / CPU baseline.
 *
 * Parity pinning: the reference (Fortran 2018 + MPI + PETSc + NetCDF) cannot be built in
 * this environment, so this file is a restatement.  It is pinned against
 * every known-answer vector the reference's own tests hold for this path (tests/golden,
 * tests/test_oracle_golden.py).  The PETSc KSP solve itself has no unit-level
 * known-answer test in the reference ("parity unpinned" for KSPSolve iterates; the
 * mathematically exact solution A^-1 b from a direct solve is the oracle there).
 *
 * All arrays are column-major, 1-based in content, as in the reference.
 * Each function cites the reference file:line it follows (paths relative to the
 * reference root).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int nV, nTri, nC_mem;
  const double *V;      /* (nV,2)   */
  const int *Tri;       /* (nTri,3) */
  const int *TriC;      /* (nTri,3) */
  const int *C;         /* (nV,nC_mem) */
  const int *nC;        /* (nV) */
  const int *iTri;      /* (nV,nC_mem) */
  const int *niTri;     /* (nV) */
  const int *VBI;       /* (nV) */
  const int *TriBI;     /* (nTri) */
  const double *TriGC;  /* (nTri,2) */
  double xmin, xmax, ymin, ymax;
} ora_mesh;

#define V_(m, vi, d) ((m)->V[(size_t)((d)-1) * (m)->nV + ((vi)-1)])
#define TRI_(m, ti, n) ((m)->Tri[(size_t)((n)-1) * (m)->nTri + ((ti)-1)])
#define TRIC_(m, ti, n) ((m)->TriC[(size_t)((n)-1) * (m)->nTri + ((ti)-1)])
#define C_(m, vi, ci) ((m)->C[(size_t)((ci)-1) * (m)->nV + ((vi)-1)])
#define ITRI_(m, vi, ci) ((m)->iTri[(size_t)((ci)-1) * (m)->nV + ((vi)-1)])
#define GC_(m, ti, d) ((m)->TriGC[(size_t)((d)-1) * (m)->nTri + ((ti)-1)])

/* ------------------------------------------------------------------------------------
 * partition_list -- src/UPSY/basic/mpi_parallelisation/mpi_distributed_memory.f90:42-68
 * ---------------------------------------------------------------------------------- */
void ora_partition_list(int ntot, int i, int n, int *i1, int *i2) {
  if (ntot > n * 2) {
    int remainder = ntot % n, slice = ntot / n;
    *i1 = slice * i + (i < remainder ? i : remainder) + 1;
    *i2 = slice * (i + 1) + ((i + 1) < remainder ? (i + 1) : remainder);
  } else if (i == 0) {
    *i1 = 1; *i2 = ntot;
  } else {
    *i1 = 1; *i2 = 0;
  }
}

/* gfortran's NORM2 intrinsic for a 2-vector (libgfortran generated/norm2_r8.c: scaled
 * sum of squares).  External to the reference tree; restated from its published
 * algorithm. */
static double norm2_2(double a, double b) {
  double scale = 1.0, ssq = 0.0;
  double v[2] = {a, b};
  for (int i = 0; i < 2; i++) {
    if (v[i] != 0.0) {
      double ax = fabs(v[i]);
      if (scale < ax) { double t = scale / ax; ssq = 1.0 + ssq * t * t; scale = ax; }
      else { double t = ax / scale; ssq += t * t; }
    }
  }
  return scale * sqrt(ssq);
}

/* ------------------------------------------------------------------------------------
 * 3x3 / 5x5 closed-form inverses -- src/UPSY/basic/math_utilities/matrix_algebra.f90
 * :110-131 (det 3x3), :133-183 (inverse 3x3), :185-457 (5x5, Leibniz expansion).
 * The 5x5 expansion is restated table-driven: permutations in ascending lexicographic
 * order of (sigma(5),sigma(4),...,sigma(1)), products taken row 1..5 left to right,
 * terms accumulated left to right -- the evaluation order of the reference expression.
 * ---------------------------------------------------------------------------------- */
static double det3(const double A[3][3]) {
  double m11 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  double m12 = A[1][0] * A[2][2] - A[1][2] * A[2][0];
  double m13 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  return A[0][0] * m11 - A[0][1] * m12 + A[0][2] * m13;
}

static int inv3(const double A[3][3], double M[3][3]) {
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
  double det = A[0][0] * m[0][0] - A[0][1] * m[0][1] + A[0][2] * m[0][2];
  if (fabs(det) < DBL_MIN) return 1;
  m[0][1] = -m[0][1]; m[1][0] = -m[1][0]; m[1][2] = -m[1][2]; m[2][1] = -m[2][1];
  /* transpose of the cofactor matrix, divided by det */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M[i][j] = m[j][i] / det;
  return 0;
}

/* permutation tables, built once */
static int perm5[120][5], sign5[120];
static int perm4[24][4], sign4[24];
static int tables_ready = 0;

static int perm_parity(const int *p, int n) {
  int inv = 0;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) if (p[i] > p[j]) inv++;
  return (inv & 1) ? -1 : 1;
}

static void build_tables(void) {
  if (tables_ready) return;
  int n5 = 0;
  /* loop order = sort key order: sigma(5) slowest ... sigma(1) fastest */
  for (int s5 = 0; s5 < 5; s5++) for (int s4 = 0; s4 < 5; s4++) for (int s3 = 0; s3 < 5; s3++)
  for (int s2 = 0; s2 < 5; s2++) for (int s1 = 0; s1 < 5; s1++) {
    int p[5] = {s1, s2, s3, s4, s5}, ok = 1;
    for (int i = 0; i < 5 && ok; i++) for (int j = i + 1; j < 5; j++) if (p[i] == p[j]) { ok = 0; break; }
    if (!ok) continue;
    memcpy(perm5[n5], p, sizeof p); sign5[n5] = perm_parity(p, 5); n5++;
  }
  int n4 = 0;
  for (int s4 = 0; s4 < 4; s4++) for (int s3 = 0; s3 < 4; s3++)
  for (int s2 = 0; s2 < 4; s2++) for (int s1 = 0; s1 < 4; s1++) {
    int p[4] = {s1, s2, s3, s4}, ok = 1;
    for (int i = 0; i < 4 && ok; i++) for (int j = i + 1; j < 4; j++) if (p[i] == p[j]) { ok = 0; break; }
    if (!ok) continue;
    memcpy(perm4[n4], p, sizeof p); sign4[n4] = perm_parity(p, 4); n4++;
  }
  tables_ready = 1;
}

static double det5(const double A[5][5]) {
  double acc = 0.0;
  for (int t = 0; t < 120; t++) {
    const int *p = perm5[t];
    double prod = A[0][p[0]] * A[1][p[1]] * A[2][p[2]] * A[3][p[3]] * A[4][p[4]];
    if (t == 0) acc = sign5[t] > 0 ? prod : -prod;
    else acc = sign5[t] > 0 ? acc + prod : acc - prod;
  }
  return acc;
}

static int inv5(const double A[5][5], double M[5][5]) {
  double det = det5(A);
  if (fabs(det) < DBL_MIN) return 1;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
    int rows[4], cols[4], nr = 0, nc = 0;
    for (int r = 0; r < 5; r++) if (r != i) rows[nr++] = r;
    for (int c = 0; c < 5; c++) if (c != j) cols[nc++] = c;
    double acc = 0.0;
    int sgn_ij = ((i + j) & 1) ? -1 : 1;
    for (int t = 0; t < 24; t++) {
      const int *p = perm4[t];
      double prod = A[rows[0]][cols[p[0]]] * A[rows[1]][cols[p[1]]] *
                    A[rows[2]][cols[p[2]]] * A[rows[3]][cols[p[3]]];
      int s = sgn_ij * sign4[t];
      if (t == 0) acc = s > 0 ? prod : -prod;
      else acc = s > 0 ? acc + prod : acc - prod;
    }
    M[j][i] = acc / det;   /* AINV = transpose(COFACTOR) / detA, matrix_algebra.f90:455 */
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * shape functions -- src/UPSY/basic/math_utilities/shape_functions.f90
 *   calc_shape_functions_2D_stag_1st_order :366-440
 *   calc_shape_functions_2D_reg_2nd_order  :218-364
 * weights w = 1/dist^q, q = 1.5 (:13).  Fortran evaluates "a * b * 1/2 * c" left to
 * right as (((a*b)*1)/2)*c, restated as such.
 * ---------------------------------------------------------------------------------- */
#define Q_EXP 1.5

static int shape_stag_1st(double x, double y, int n_c, const double *x_c, const double *y_c,
                          double *Nf, double *Nfx, double *Nfy) {
  double A[3][3] = {{0}}, M[3][3];
  double dx[n_c], dy[n_c], w[n_c];
  for (int i = 0; i < n_c; i++) { dx[i] = x_c[i] - x; dy[i] = y_c[i] - y; }
  for (int i = 0; i < n_c; i++) w[i] = 1.0 / pow(norm2_2(dx[i], dy[i]), Q_EXP);
  for (int i = 0; i < n_c; i++) {
    double w2 = w[i] * w[i];
    A[0][0] += w2 * 1.0 * 1.0;   A[0][1] += w2 * 1.0 * dx[i];   A[0][2] += w2 * 1.0 * dy[i];
    A[1][0] += w2 * dx[i] * 1.0; A[1][1] += w2 * dx[i] * dx[i]; A[1][2] += w2 * dx[i] * dy[i];
    A[2][0] += w2 * dy[i] * 1.0; A[2][1] += w2 * dy[i] * dx[i]; A[2][2] += w2 * dy[i] * dy[i];
  }
  if (fabs(det3(A)) <= DBL_MIN) return 1;       /* :421, "<=" here */
  if (inv3(A, M)) return 1;
  for (int i = 0; i < n_c; i++) {
    double w2 = w[i] * w[i];
    Nf[i]  = w2 * ((M[0][0] * 1.0) + (M[0][1] * dx[i]) + (M[0][2] * dy[i]));
    Nfx[i] = w2 * ((M[1][0] * 1.0) + (M[1][1] * dx[i]) + (M[1][2] * dy[i]));
    Nfy[i] = w2 * ((M[2][0] * 1.0) + (M[2][1] * dx[i]) + (M[2][2] * dy[i]));
  }
  return 0;
}

static int shape_reg_2nd(double x, double y, int n_c, const double *x_c, const double *y_c,
                         double Ni[5], double *Nfx, double *Nfy, double *Nfxx, double *Nfxy,
                         double *Nfyy) {
  double A[5][5] = {{0}}, M[5][5];
  double dx[n_c], dy[n_c], w[n_c];
  for (int i = 0; i < n_c; i++) { dx[i] = x_c[i] - x; dy[i] = y_c[i] - y; }
  for (int i = 0; i < n_c; i++) w[i] = 1.0 / pow(norm2_2(dx[i], dy[i]), Q_EXP);
  for (int i = 0; i < n_c; i++) {
    double w2 = w[i] * w[i], X = dx[i], Y = dy[i], X2 = X * X, Y2 = Y * Y;
    /* row factors f_r, evaluated left to right together with w2 */
    double r1 = w2 * X;            /* w^2 * dx                  */
    double r2 = w2 * Y;            /* w^2 * dy                  */
    double r3 = w2 * 1.0 / 2.0 * X2; /* w^2 * 1/2 * dx^2          */
    double r4 = w2 * X * Y;        /* w^2 * dx * dy             */
    double r5 = w2 * 1.0 / 2.0 * Y2; /* w^2 * 1/2 * dy^2          */
    double r[5] = {r1, r2, r3, r4, r5};
    for (int k = 0; k < 5; k++) {
      A[k][0] += r[k] * X;
      A[k][1] += r[k] * Y;
      A[k][2] += r[k] * 1.0 / 2.0 * X2;
      A[k][3] += r[k] * X * Y;
      A[k][4] += r[k] * 1.0 / 2.0 * Y2;
    }
  }
  if (fabs(det5(A)) < DBL_MIN) return 1;        /* :300 */
  if (inv5(A, M)) return 1;
  double s[5] = {0, 0, 0, 0, 0};
  double *out[5] = {Nfx, Nfy, Nfxx, Nfxy, Nfyy};
  for (int i = 0; i < n_c; i++) {
    double w2 = w[i] * w[i], X = dx[i], Y = dy[i], X2 = X * X, Y2 = Y * Y;
    for (int k = 0; k < 5; k++) {
      out[k][i] = w2 * ((M[k][0] * X) + (M[k][1] * Y) + (M[k][2] * 1.0 / 2.0 * X2) +
                        (M[k][3] * X * Y) + (M[k][4] * 1.0 / 2.0 * Y2));
    }
  }
  for (int k = 0; k < 5; k++) { for (int i = 0; i < n_c; i++) s[k] += out[k][i]; Ni[k] = -s[k]; }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * BFS neighbourhood growth -- src/UPSY/mesh/mesh_utilities.f90:1856-1894 (a), :1896-1935 (b)
 * ---------------------------------------------------------------------------------- */
static void extend_group_a(const ora_mesh *m, int *map, int *stack, int *stackN) {
  int n = *stackN;
  for (int i = 0; i < n; i++) {
    int vi = stack[i];
    for (int ci = 1; ci <= m->nC[vi - 1]; ci++) {
      int vj = C_(m, vi, ci);
      if (map[vj - 1] == 0) { map[vj - 1] = 1; stack[(*stackN)++] = vj; }
    }
  }
}
static void extend_group_b(const ora_mesh *m, int *map, int *stack, int *stackN) {
  int n = *stackN;
  for (int i = 0; i < n; i++) {
    int ti = stack[i];
    for (int n2 = 1; n2 <= 3; n2++) {
      int tj = TRIC_(m, ti, n2);
      if (tj == 0) continue;
      if (map[tj - 1] == 0) { map[tj - 1] = 1; stack[(*stackN)++] = tj; }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * calc_matrix_operators_mesh_a_b -- mesh_disc_calc_matrix_operators_2D.f90:198-335
 * rows row1..row2 (1-based, inclusive) of M_map_a_b / M_ddx_a_b / M_ddy_a_b.
 * ptr has (row2-row1+2) entries, local 1-based offsets (CSR_sparse_matrix_type.f90).
 * returns nnz, or -1 if cap exceeded.
 * ---------------------------------------------------------------------------------- */
int ora_calc_operators_a_b(const ora_mesh *m, int row1, int row2, int cap, int *ptr, int *ind,
                           double *vmap, double *vddx, double *vddy) {
  build_tables();
  const int n_min = 3, n_max = m->nC_mem * m->nC_mem;
  int *map = calloc(m->nV, sizeof(int)), *stack = calloc(m->nV, sizeof(int));
  int stackN = 0, nnz = 0;
  int *i_c = malloc(n_max * sizeof(int));
  double *x_c = malloc(n_max * sizeof(double)), *y_c = malloc(n_max * sizeof(double));
  double *Nf = malloc(n_max * sizeof(double)), *Nfx = malloc(n_max * sizeof(double)),
         *Nfy = malloc(n_max * sizeof(double));
  ptr[0] = 1;
  for (int row = row1; row <= row2; row++) {
    int ti = row;                       /* n2ti = identity, mesh_translation_tables.f90 */
    double x = GC_(m, ti, 1), y = GC_(m, ti, 2);
    for (int i = 0; i < stackN; i++) map[stack[i] - 1] = 0;
    stackN = 0;
    for (int n = 1; n <= 3; n++) { int vi = TRI_(m, ti, n); map[vi - 1] = 1; stack[stackN++] = vi; }
    while (stackN < n_min) extend_group_a(m, map, stack, &stackN);
    int n_c = 0, ok = 0;
    while (!ok) {
      n_c = 0;
      for (int i = 0; i < stackN; i++) {
        if (n_c == n_max) break;
        int vi = stack[i];
        i_c[n_c] = vi; x_c[n_c] = V_(m, vi, 1); y_c[n_c] = V_(m, vi, 2); n_c++;
      }
      ok = !shape_stag_1st(x, y, n_c, x_c, y_c, Nf, Nfx, Nfy);
      if (!ok) extend_group_a(m, map, stack, &stackN);
    }
    if (nnz + n_c > cap) { nnz = -1; break; }
    for (int i = 0; i < n_c; i++) {
      ind[nnz] = i_c[i]; vmap[nnz] = Nf[i]; vddx[nnz] = Nfx[i]; vddy[nnz] = Nfy[i]; nnz++;
    }
    ptr[row - row1 + 1] = nnz + 1;
  }
  free(map); free(stack); free(i_c); free(x_c); free(y_c); free(Nf); free(Nfx); free(Nfy);
  return nnz;
}

/* calc_matrix_operators_mesh_b_a -- mesh_disc_calc_matrix_operators_2D.f90:337-474 */
int ora_calc_operators_b_a(const ora_mesh *m, int row1, int row2, int cap, int *ptr, int *ind,
                           double *vmap, double *vddx, double *vddy) {
  build_tables();
  const int n_min = 3, n_max = m->nC_mem * m->nC_mem;
  int *map = calloc(m->nTri, sizeof(int)), *stack = calloc(m->nTri, sizeof(int));
  int stackN = 0, nnz = 0;
  int *i_c = malloc(n_max * sizeof(int));
  double *x_c = malloc(n_max * sizeof(double)), *y_c = malloc(n_max * sizeof(double));
  double *Nf = malloc(n_max * sizeof(double)), *Nfx = malloc(n_max * sizeof(double)),
         *Nfy = malloc(n_max * sizeof(double));
  ptr[0] = 1;
  for (int row = row1; row <= row2; row++) {
    int vi = row;
    double x = V_(m, vi, 1), y = V_(m, vi, 2);
    for (int i = 0; i < stackN; i++) map[stack[i] - 1] = 0;
    stackN = 0;
    for (int iti = 1; iti <= m->niTri[vi - 1]; iti++) {
      int ti = ITRI_(m, vi, iti); map[ti - 1] = 1; stack[stackN++] = ti;
    }
    while (stackN < n_min) extend_group_b(m, map, stack, &stackN);
    int n_c = 0, ok = 0;
    while (!ok) {
      n_c = 0;
      for (int i = 0; i < stackN; i++) {
        if (n_c == n_max) break;
        int ti = stack[i];
        i_c[n_c] = ti; x_c[n_c] = GC_(m, ti, 1); y_c[n_c] = GC_(m, ti, 2); n_c++;
      }
      ok = !shape_stag_1st(x, y, n_c, x_c, y_c, Nf, Nfx, Nfy);
      if (!ok) extend_group_b(m, map, stack, &stackN);
    }
    if (nnz + n_c > cap) { nnz = -1; break; }
    for (int i = 0; i < n_c; i++) {
      ind[nnz] = i_c[i]; vmap[nnz] = Nf[i]; vddx[nnz] = Nfx[i]; vddy[nnz] = Nfy[i]; nnz++;
    }
    ptr[row - row1 + 1] = nnz + 1;
  }
  free(map); free(stack); free(i_c); free(x_c); free(y_c); free(Nf); free(Nfx); free(Nfy);
  return nnz;
}

/* calc_matrix_operators_mesh_b_b_2nd_order -- mesh_disc_calc_matrix_operators_2D.f90:612-764
 * five matrices share one pattern; vals[5] = ddx, ddy, d2dx2, d2dxdy, d2dy2. */
int ora_calc_operators_b_b_2nd(const ora_mesh *m, int row1, int row2, int cap, int *ptr, int *ind,
                               double *vddx, double *vddy, double *vxx, double *vxy, double *vyy) {
  build_tables();
  const int n_min = 5, n_max = m->nC_mem * m->nC_mem;
  int *map = calloc(m->nTri, sizeof(int)), *stack = calloc(m->nTri, sizeof(int));
  int stackN = 0, nnz = 0;
  int *i_c = malloc(n_max * sizeof(int));
  double *x_c = malloc(n_max * sizeof(double)), *y_c = malloc(n_max * sizeof(double));
  double *N[5];
  for (int k = 0; k < 5; k++) N[k] = malloc(n_max * sizeof(double));
  double *out[5] = {vddx, vddy, vxx, vxy, vyy};
  ptr[0] = 1;
  for (int row = row1; row <= row2; row++) {
    int ti = row;
    double x = GC_(m, ti, 1), y = GC_(m, ti, 2), Ni[5];
    for (int i = 0; i < stackN; i++) map[stack[i] - 1] = 0;
    map[ti - 1] = 1; stackN = 1; stack[0] = ti;
    while (stackN - 1 < n_min) extend_group_b(m, map, stack, &stackN);
    int n_c = 0, ok = 0;
    while (!ok) {
      n_c = 0;
      for (int i = 0; i < stackN; i++) {
        if (n_c == n_max) break;
        int tj = stack[i];
        if (tj == ti) continue;
        i_c[n_c] = tj; x_c[n_c] = GC_(m, tj, 1); y_c[n_c] = GC_(m, tj, 2); n_c++;
      }
      ok = !shape_reg_2nd(x, y, n_c, x_c, y_c, Ni, N[0], N[1], N[2], N[3], N[4]);
      if (!ok) extend_group_b(m, map, stack, &stackN);
    }
    if (nnz + n_c + 1 > cap) { nnz = -1; break; }
    ind[nnz] = row;                      /* diagonal first, :734-738 */
    for (int k = 0; k < 5; k++) out[k][nnz] = Ni[k];
    nnz++;
    for (int i = 0; i < n_c; i++) {
      ind[nnz] = i_c[i];
      for (int k = 0; k < 5; k++) out[k][nnz] = N[k][i];
      nnz++;
    }
    ptr[row - row1 + 1] = nnz + 1;
  }
  free(map); free(stack); free(i_c); free(x_c); free(y_c);
  for (int k = 0; k < 5; k++) free(N[k]);
  return nnz;
}

/* ------------------------------------------------------------------------------------
 * SpMV -- src/UPSY/basic/CSR_matrix_algebra/CSR_matrix_vector_multiplication.f90:266-278,
 * :317-329: y(i) = sum_k val(k) * x(ind(k)), scalar row loop, x indexed globally.
 * ptr local 1-based offsets for rows i1..i1+m_loc-1; x_tot covers all columns.
 * ---------------------------------------------------------------------------------- */
void ora_spmv(int m_loc, const int *ptr, const int *ind, const double *val, const double *x_tot,
              double *y) {
  for (int i = 0; i < m_loc; i++) {
    double s = 0.0;
    for (int k = ptr[i]; k <= ptr[i + 1] - 1; k++) s += val[k - 1] * x_tot[ind[k - 1] - 1];
    y[i] = s;
  }
}
/* threaded variant used only by the CPU-baseline timing leg */
void ora_spmv_mt(int m_loc, const int *ptr, const int *ind, const double *val, const double *x_tot,
                 double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m_loc; i++) {
    double s = 0.0;
    for (int k = ptr[i]; k <= ptr[i + 1] - 1; k++) s += val[k - 1] * x_tot[ind[k - 1] - 1];
    y[i] = s;
  }
}

/* ------------------------------------------------------------------------------------
 * find_containing_vertex -- src/UPSY/mesh/mesh_utilities.f90:1368-1412
 * ---------------------------------------------------------------------------------- */
static int find_containing_vertex(const ora_mesh *m, double px, double py, int vi) {
  int vi_prev = vi;
  for (;;) {
    double d = norm2_2(V_(m, vi, 1) - px, V_(m, vi, 2) - py);
    double dcmin = d + 10.0; int vcmin = 0;
    for (int ci = 1; ci <= m->nC[vi - 1]; ci++) {
      int vc = C_(m, vi, ci);
      if (vc == vi_prev) continue;
      double dc = norm2_2(V_(m, vc, 1) - px, V_(m, vc, 2) - py);
      if (dc < dcmin) { dcmin = dc; vcmin = vc; }
    }
    if (dcmin < d) { vi_prev = vi; vi = vcmin; } else return vi;
  }
}

/* find_ti_copy_ISMIP_HOM_periodic (mesh_utilities.f90:2623-2679) when kind==1,
 * find_ti_copy_SSA_icestream_infinite (:2681-2730) when kind==2.
 * ti_copy / wti_copy have nC_mem entries. */
void ora_find_ti_copy(const ora_mesh *m, int kind, double L, int ti, int *ti_copy, double *wti_copy) {
  double gx = GC_(m, ti, 1), gy = GC_(m, ti, 2), px, py;
  if (kind == 1) {
    px = (gx > 0.0) ? gx - L / 2.0 : gx + L / 2.0;
    py = (gy > 0.0) ? gy - L / 2.0 : gy + L / 2.0;
  } else {
    py = gy;
    px = (gx < 0.0) ? m->xmin + (m->xmax - m->xmin) * 1.0 / 3.0
                    : m->xmin + (m->xmax - m->xmin) * 2.0 / 3.0;
  }
  int vi = find_containing_vertex(m, px, py, 5);
  for (int n = 0; n < m->nC_mem; n++) { ti_copy[n] = 0; wti_copy[n] = 0.0; }
  double sum = 0.0;
  int nt = m->niTri[vi - 1];
  for (int iti = 1; iti <= nt; iti++) {
    int tj = ITRI_(m, vi, iti);
    double dist = norm2_2(px - GC_(m, tj, 1), py - GC_(m, tj, 2));
    ti_copy[iti - 1] = tj;
    wti_copy[iti - 1] = 1.0 / (dist * dist);
  }
  for (int i = 0; i < nt; i++) sum += wti_copy[i];
  for (int i = 0; i < nt; i++) wti_copy[i] = wti_copy[i] / sum;
}

/* ------------------------------------------------------------------------------------
 * Stiffness-matrix assembly -- src/UFEMISM/ice_dynamics/conservation_of_momentum/SSA_DIVA/
 * solve_linearised_SSA_DIVA.f90:23-153 (driver loop :89-151), row builders
 * _row_free :180-329, _sans_ :331-479, _row_BC :481-641.
 *
 * bc_u[4], bc_v[4]: BC codes for north, east, south, west:
 *   1 'infinite', 2 'zero', 3 'periodic_ISMIP-HOM', 4 'infinite_SSA_icestream'.
 * rows ti1..ti2 (triangles) -> matrix rows 2*ti1-1 .. 2*ti2.
 * M2 arrays: full matrices (all nTri rows), ptr 1-based.
 * u_b_prev, v_b_prev: (nTri).  returns nnz or -1.
 * ---------------------------------------------------------------------------------- */
int ora_assemble_stiffness(const ora_mesh *m, int ti1, int ti2,
                           const int *m2ptr, const int *m2ind, const double *ddx, const double *ddy,
                           const double *d2dx2, const double *d2dxdy, const double *d2dy2,
                           const double *N_b, const double *dN_dx_b, const double *dN_dy_b,
                           const double *beta_b, const double *tau_dx_b, const double *tau_dy_b,
                           const double *u_b_prev, const double *v_b_prev,
                           const int *bc_mask, const double *bc_u_val, const double *bc_v_val,
                           const int *bc_u, const int *bc_v, int crossterms,
                           double visc_it_relax, double ISMIP_HOM_L,
                           int cap, int *ptr, int *ind, double *val, double *bb) {
  int nnz = 0;
  int *ti_copy = malloc(m->nC_mem * sizeof(int));
  double *wti_copy = malloc(m->nC_mem * sizeof(double));
  ptr[0] = 1;
#define ADD(col, v) do { if (nnz >= cap) { nnz = -1; goto done; } ind[nnz] = (col); val[nnz] = (v); nnz++; } while (0)
  for (int row = 2 * ti1 - 1; row <= 2 * ti2; row++) {
    int ti = (row + 1) / 2, uv = (row % 2 == 1) ? 1 : 2;   /* n2tiuv */
    int r = row - (2 * ti1 - 1);
    /* fields are indexed by global triangle (full-length arrays) */
    if (bc_mask[ti - 1] == 1) {
      ADD(row, 1.0);
      bb[r] = (uv == 1) ? bc_u_val[ti - 1] : bc_v_val[ti - 1];
    } else if (m->TriBI[ti - 1] > 0) {
      int side;
      switch (m->TriBI[ti - 1]) {
        case 1: case 2: side = 0; break;
        case 3: case 4: side = 1; break;
        case 5: case 6: side = 2; break;
        default: side = 3; break;
      }
      int choice = (uv == 1) ? bc_u[side] : bc_v[side];
      const double *w_prev = (uv == 1) ? u_b_prev : v_b_prev;
      if (choice == 1) {
        int nn = 0;
        for (int n = 1; n <= 3; n++) {
          int tj = TRIC_(m, ti, n);
          if (tj == 0) continue;
          nn++;
          ADD(2 * (tj - 1) + uv, 1.0);
        }
        ADD(row, -1.0 * (double)nn);
        bb[r] = 0.0;
      } else if (choice == 2) {
        ADD(row, 1.0);
        bb[r] = 0.0;
      } else {
        ora_find_ti_copy(m, choice == 3 ? 1 : 2, ISMIP_HOM_L, ti, ti_copy, wti_copy);
        ADD(row, 1.0);
        double fixed = 0.0;
        for (int n = 0; n < m->nC_mem; n++) {
          int tj = ti_copy[n];
          if (tj == 0) continue;
          fixed = fixed + wti_copy[n] * w_prev[tj - 1];
        }
        fixed = (visc_it_relax * fixed) + ((1.0 - visc_it_relax) * w_prev[ti - 1]);
        bb[r] = fixed;
      }
    } else {
      double N = N_b[ti - 1], Nx = dN_dx_b[ti - 1], Ny = dN_dy_b[ti - 1];
      double beta = beta_b[ti - 1], tdx = tau_dx_b[ti - 1], tdy = tau_dy_b[ti - 1];
      for (int k = m2ptr[ti - 1]; k <= m2ptr[ti] - 1; k++) {
        int tj = m2ind[k - 1];
        double dx_ = ddx[k - 1], dy_ = ddy[k - 1], xx = d2dx2[k - 1], xy = d2dxdy[k - 1], yy = d2dy2[k - 1];
        double Au, Av;
        if (crossterms) {
          if (uv == 1) {
            Au = 4.0 * N * xx + 4.0 * Nx * dx_ + N * yy + Ny * dy_;
            if (tj == ti) Au = Au - beta;
            Av = 3.0 * N * xy + 2.0 * Nx * dy_ + Ny * dx_;
          } else {
            Av = 4.0 * N * yy + 4.0 * Ny * dy_ + N * xx + Nx * dx_;
            if (tj == ti) Av = Av - beta;
            Au = 3.0 * N * xy + 2.0 * Ny * dx_ + Nx * dy_;
          }
        } else {
          if (uv == 1) {
            Au = 4.0 * xx + yy;
            if (tj == ti) Au = Au - beta / N;
            Av = 3.0 * xy;
          } else {
            Av = 4.0 * yy + xx;
            if (tj == ti) Av = Av - beta / N;
            Au = 3.0 * xy;
          }
        }
        ADD(2 * (tj - 1) + 1, Au);
        ADD(2 * (tj - 1) + 2, Av);
      }
      if (crossterms) bb[r] = (uv == 1) ? -tdx : -tdy;
      else bb[r] = (uv == 1) ? -tdx / N : -tdy / N;
    }
    ptr[r + 1] = nnz + 1;
  }
done:
#undef ADD
  free(ti_copy); free(wti_copy);
  return nnz;
}

/* ------------------------------------------------------------------------------------
 * PETSc KSP defaults, restated (PETSc is an external, un-vendored dependency of the
 * reference: PETSc 3.22, call sequence src/UPSY/basic/petsc_basic.f90:66-141):
 *   KSPGMRES, restart 30, classical Gram-Schmidt without refinement, left
 *   preconditioning, PCBJACOBI with one block per rank and ILU(0) on each block
 *   (natural ordering, columns sorted), zero initial guess, default convergence test
 *   on the preconditioned residual norm  ||B r|| <= max(rtol*||B b||, abstol),
 *   divergence at ||B r|| > dtol*||B b|| (dtol 1e5), maxits 10000.
 * "ranks" are contiguous row blocks from partition_list; each OpenMP thread plays one
 * rank.  Matrix: square n x n, ptr 1-based (global), ind 1-based global, unsorted.
 * Returns iteration count (KSPGetIterationNumber); *reason: 2 rtol, 3 atol, -3 its, -4 dtol.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int r0, r1;            /* 0-based row range [r0,r1) */
  int *lptr, *lind;      /* local block CSR (0-based, sorted columns, local col ids) */
  double *lval;
  int *diag;             /* position of the diagonal in each local row */
} ilu_block;

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static void ilu0_setup(ilu_block *B, const int *ptr, const int *ind, const double *val) {
  int nloc = B->r1 - B->r0;
  B->lptr = malloc((nloc + 1) * sizeof(int));
  B->diag = malloc(nloc * sizeof(int));
  int cnt = 0;
  B->lptr[0] = 0;
  for (int i = B->r0; i < B->r1; i++) {
    for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) {
      int c = ind[k] - 1;
      if (c >= B->r0 && c < B->r1) cnt++;
    }
    B->lptr[i - B->r0 + 1] = cnt;
  }
  B->lind = malloc((cnt > 0 ? cnt : 1) * sizeof(int));
  B->lval = malloc((cnt > 0 ? cnt : 1) * sizeof(double));
  int maxrow = 0;
  for (int i = 0; i < nloc; i++) { int l = B->lptr[i + 1] - B->lptr[i]; if (l > maxrow) maxrow = l; }
  int *tmpi = malloc((maxrow + 1) * 2 * sizeof(int));
  for (int i = B->r0; i < B->r1; i++) {
    int li = i - B->r0, n = 0;
    for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) {
      int c = ind[k] - 1;
      if (c >= B->r0 && c < B->r1) { tmpi[2 * n] = c - B->r0; tmpi[2 * n + 1] = k; n++; }
    }
    qsort(tmpi, n, 2 * sizeof(int), cmp_int);
    int base = B->lptr[li];
    B->diag[li] = -1;
    for (int j = 0; j < n; j++) {
      B->lind[base + j] = tmpi[2 * j];
      B->lval[base + j] = val[tmpi[2 * j + 1]];
      if (tmpi[2 * j] == li) B->diag[li] = base + j;
    }
  }
  free(tmpi);
  /* IKJ ILU(0) */
  int *pos = malloc(nloc * sizeof(int));
  for (int i = 0; i < nloc; i++) pos[i] = -1;
  for (int i = 0; i < nloc; i++) {
    for (int k = B->lptr[i]; k < B->lptr[i + 1]; k++) pos[B->lind[k]] = k;
    for (int k = B->lptr[i]; k < B->lptr[i + 1]; k++) {
      int j = B->lind[k];
      if (j >= i) break;
      double piv = B->lval[B->diag[j]];
      double l = B->lval[k] / piv;
      B->lval[k] = l;
      for (int kk = B->diag[j] + 1; kk < B->lptr[j + 1]; kk++) {
        int p = pos[B->lind[kk]];
        if (p >= 0) B->lval[p] -= l * B->lval[kk];
      }
    }
    for (int k = B->lptr[i]; k < B->lptr[i + 1]; k++) pos[B->lind[k]] = -1;
  }
  free(pos);
}

static void ilu0_apply(const ilu_block *B, const double *r, double *z) {
  int nloc = B->r1 - B->r0;
  const double *rl = r + B->r0;
  double *zl = z + B->r0;
  for (int i = 0; i < nloc; i++) {
    double s = rl[i];
    for (int k = B->lptr[i]; k < B->diag[i]; k++) s -= B->lval[k] * zl[B->lind[k]];
    zl[i] = s;
  }
  for (int i = nloc - 1; i >= 0; i--) {
    double s = zl[i];
    for (int k = B->diag[i] + 1; k < B->lptr[i + 1]; k++) s -= B->lval[k] * zl[B->lind[k]];
    zl[i] = s / B->lval[B->diag[i]];
  }
}

static void ilu0_free(ilu_block *B) { free(B->lptr); free(B->lind); free(B->lval); free(B->diag); }

int ora_ksp_gmres_bjacobi_ilu0(int n, const int *ptr, const int *ind, const double *val,
                               const double *b, double *x, double rtol, double abstol,
                               int nranks, int maxits, int *reason, double *rnorm_out) {
  const int restart = 30;
  const double dtol = 1.0e5;
  if (nranks < 1) nranks = 1;
  /* PETSc splits the n matrix rows like partition_list over the 2*nTri unknowns */
  ilu_block *blk = calloc(nranks, sizeof(ilu_block));
  for (int p = 0; p < nranks; p++) {
    int i1, i2;
    ora_partition_list(n, p, nranks, &i1, &i2);
    blk[p].r0 = i1 - 1; blk[p].r1 = i2;
  }
#pragma omp parallel for schedule(static, 1) num_threads(nranks)
  for (int p = 0; p < nranks; p++) if (blk[p].r1 > blk[p].r0) ilu0_setup(&blk[p], ptr, ind, val);

  double *Vb = malloc((size_t)(restart + 1) * n * sizeof(double));
  double *w = malloc((size_t)n * sizeof(double)), *t = malloc((size_t)n * sizeof(double));
  double H[31][30], cs[30], sn[30], g[31], hcol[31];
  int its = 0; *reason = 0;
  memset(x, 0, (size_t)n * sizeof(double));     /* zero initial guess */

#define PC_APPLY(in, out) do { _Pragma("omp parallel for schedule(static,1) num_threads(nranks)") \
    for (int p_ = 0; p_ < nranks; p_++) if (blk[p_].r1 > blk[p_].r0) ilu0_apply(&blk[p_], (in), (out)); } while (0)
#define MATMULT(in, out) do { _Pragma("omp parallel for schedule(static) num_threads(nranks)") \
    for (int i_ = 0; i_ < n; i_++) { double s_ = 0.0; \
      for (int k_ = ptr[i_] - 1; k_ < ptr[i_ + 1] - 1; k_++) s_ += val[k_] * (in)[ind[k_] - 1]; (out)[i_] = s_; } } while (0)

  /* ||B b|| : with a zero initial guess the first preconditioned residual is B b */
  PC_APPLY(b, w);
  double bnorm = 0.0;
#pragma omp parallel for reduction(+ : bnorm) num_threads(nranks)
  for (int i = 0; i < n; i++) bnorm += w[i] * w[i];
  bnorm = sqrt(bnorm);
  double ttol = fmax(rtol * bnorm, abstol);
  double rnorm = bnorm;
  if (rnorm <= ttol) { *reason = (rnorm <= abstol) ? 3 : 2; goto finish; }

  while (its < maxits && *reason == 0) {
    /* r = B (b - A x) */
    if (its == 0) { memcpy(Vb, w, (size_t)n * sizeof(double)); }
    else {
      MATMULT(x, t);
#pragma omp parallel for num_threads(nranks)
      for (int i = 0; i < n; i++) t[i] = b[i] - t[i];
      PC_APPLY(t, Vb);
      double s = 0.0;
#pragma omp parallel for reduction(+ : s) num_threads(nranks)
      for (int i = 0; i < n; i++) s += Vb[i] * Vb[i];
      rnorm = sqrt(s);
    }
    double beta = rnorm;
#pragma omp parallel for num_threads(nranks)
    for (int i = 0; i < n; i++) Vb[i] /= beta;
    memset(g, 0, sizeof g); g[0] = beta;
    int j = 0;
    for (; j < restart && its < maxits; j++) {
      double *vj = Vb + (size_t)j * n, *vn = Vb + (size_t)(j + 1) * n;
      MATMULT(vj, t);
      PC_APPLY(t, vn);
      /* classical Gram-Schmidt: all dots against the un-updated vector, then one update */
      for (int i = 0; i <= j; i++) {
        const double *vi = Vb + (size_t)i * n; double s = 0.0;
#pragma omp parallel for reduction(+ : s) num_threads(nranks)
        for (int q = 0; q < n; q++) s += vn[q] * vi[q];
        hcol[i] = s;
      }
      for (int i = 0; i <= j; i++) {
        const double *vi = Vb + (size_t)i * n; double h = hcol[i];
#pragma omp parallel for num_threads(nranks)
        for (int q = 0; q < n; q++) vn[q] -= h * vi[q];
      }
      double s = 0.0;
#pragma omp parallel for reduction(+ : s) num_threads(nranks)
      for (int q = 0; q < n; q++) s += vn[q] * vn[q];
      double hn = sqrt(s);
      hcol[j + 1] = hn;
      if (hn != 0.0) {
#pragma omp parallel for num_threads(nranks)
        for (int q = 0; q < n; q++) vn[q] /= hn;
      }
      /* apply previous Givens rotations, form the new one */
      for (int i = 0; i < j; i++) {
        double a = hcol[i], c = hcol[i + 1];
        hcol[i] = cs[i] * a + sn[i] * c;
        hcol[i + 1] = -sn[i] * a + cs[i] * c;
      }
      double a = hcol[j], c = hcol[j + 1], rr = hypot(a, c);
      if (rr == 0.0) { cs[j] = 1.0; sn[j] = 0.0; } else { cs[j] = a / rr; sn[j] = c / rr; }
      hcol[j] = rr; hcol[j + 1] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      for (int i = 0; i <= j; i++) H[i][j] = hcol[i];
      rnorm = fabs(g[j + 1]);
      its++;
      if (rnorm <= ttol) { *reason = (rnorm <= abstol) ? 3 : 2; j++; break; }
      if (rnorm >= dtol * bnorm) { *reason = -4; j++; break; }
      if (hn == 0.0) { j++; break; }   /* happy breakdown */
    }
    /* solve the upper-triangular system and update x */
    double yv[30];
    for (int i = j - 1; i >= 0; i--) {
      double s = g[i];
      for (int k = i + 1; k < j; k++) s -= H[i][k] * yv[k];
      yv[i] = s / H[i][i];
    }
    for (int i = 0; i < j; i++) {
      const double *vi = Vb + (size_t)i * n; double yi = yv[i];
#pragma omp parallel for num_threads(nranks)
      for (int q = 0; q < n; q++) x[q] += yi * vi[q];
    }
    if (its >= maxits && *reason == 0) *reason = -3;
  }
finish:
  if (rnorm_out) *rnorm_out = rnorm;
  for (int p = 0; p < nranks; p++) if (blk[p].r1 > blk[p].r0) ilu0_free(&blk[p]);
  free(blk); free(Vb); free(w); free(t);
  return its;
#undef PC_APPLY
#undef MATMULT
}

/* Jacobi iteration on CSR -- src/UPSY/basic/CSR_matrix_algebra/CSR_matrix_solving.f90
 * :117-226 (second known-answer aid): x_i <- (b_i - sum_{j!=i} a_ij x_j) / a_ii,
 * at most nit sweeps, stop when max|dx| < tol.  Returns sweeps done. */
int ora_jacobi(int n, const int *ptr, const int *ind, const double *val, const double *b, double *x,
               int nit, double tol) {
  double *xn = malloc((size_t)n * sizeof(double));
  int it = 0;
  for (; it < nit; it++) {
    double maxd = 0.0;
    for (int i = 0; i < n; i++) {
      double lhs = 0.0, cij = 0.0;
      for (int k = ptr[i] - 1; k < ptr[i + 1] - 1; k++) {
        int j = ind[k] - 1;
        if (j == i) cij = val[k]; else lhs += val[k] * x[j];
      }
      xn[i] = (b[i] - lhs) / cij;
      double d = fabs(xn[i] - x[i]); if (d > maxd) maxd = d;
    }
    memcpy(x, xn, (size_t)n * sizeof(double));
    if (maxd < tol) { it++; break; }
  }
  free(xn);
  return it;
}
