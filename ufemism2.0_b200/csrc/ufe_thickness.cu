// Ice-thickness rates of change on the device.
//
//   calc_dHi_dt_explicit        src/UFEMISM/ice_dynamics/conservation_of_mass/conservation_of_mass_explicit.f90:23-138
//   apply_ice_thickness_BC_explicit                                                     :140-282
//   calc_dHi_dt_semiimplicit    .../conservation_of_mass_semiimplicit.f90:24-173 (+ BC rows :175-313)
//   calc_ice_flux_divergence_matrix_upwind, calc_flux_limited_timestep, calc_n_interior_neighbours
//                               .../conservation_of_mass_utilities.f90:21-131, 155-203, 205-233
//   map_velocities_from_b_to_c_2D   src/UFEMISM/ice_dynamics/utilities/map_velocities_to_c_grid.f90:17-69
//
// Layout: M_divQ / AA share one fixed pattern, row vi = [vi, C(vi,1..nC(vi))] (the order
// add_entry_CSR_dist is called in), built once per mesh.  One thread per vertex; the (nV,nC_mem)
// column-major mesh arrays make every per-connection load of a warp one coalesced line.
// HBM-bound; algorithmic bytes per vertex of k_thk_divq (nC ~ 6): VE 4 nC + Cw,D_x,D_y,D 32 nC +
// ETri 8 nC + (u,v) of two triangles 32 nC + C 4 nC + fraction_margin, Hi gathers 16 nC +
// val out 8 (nC+1) + 8 vertex fields in/out 64  ~= 104 nC + 72 ~= 700 B.
#include "ufe_internal.cuh"

struct ThkMesh {
  int nV, nTri, nC_mem, nE;
  const int *C, *nC, *VBI, *VE, *ETri, *ptr;
  const double *A, *Cw, *Dx, *Dy, *D;
};

struct ThicknessState {
  int nV = 0, nTri = 0, nE = 0, nnz = 0, nb = 0;
  int *VE = nullptr, *ETri = nullptr, *border = nullptr;
  double *A = nullptr, *Cw = nullptr, *Dx = nullptr, *Dy = nullptr, *D = nullptr;
  // fields (nV)
  double *Hi = nullptr, *Hb = nullptr, *SL = nullptr, *SMB = nullptr, *BMB = nullptr, *LMB = nullptr, *fm = nullptr,
         *target = nullptr, *bcHi = nullptr, *divQ = nullptr, *dHi_dt = nullptr, *AMB = nullptr, *Hi_tp = nullptr,
         *Hs0 = nullptr, *Hs1 = nullptr, *u_b = nullptr, *v_b = nullptr;
  int *noice = nullptr, *bcmask = nullptr;
  unsigned long long *dtlim = nullptr;   // bit pattern of the (positive) minimum of dt_lim
  double *dt_dev = nullptr;              // the constrained time step (device scalar)
  double *Mval = nullptr;                // M_divQ values
  DevSystem S;                           // AA (val), B*AA (valS), bb, bS, x
  KrylovWork kw;
  bool kw_alloc = false, have_system = false, have_M = false;
  // calc_vertical_velocities
  DevFamily aa;                          // M_ddx_a_a, M_ddy_a_a (built on first use)
  bool aa_built = false;
  double *vv_in[5] = {};                 // Hib, dHb_dt, dHi_dt, BMB, Hi (nV)
  double *vv_dz[3] = {};                 // dzeta_dx_ak, dzeta_dy_ak, dzeta_dz_ak (nV,nz)
  double *w_3D = nullptr;                // (nV,nz)
  int *vv_mask[2] = {};                  // mask_grounded_ice, mask_floating_ice
  int *found_negative = nullptr;         // calc_dHi_dt: Hi_tplusdt < -0.1 m somewhere
  cudaEvent_t ev[7] = {};                // start | inputs up | M_divQ + divQ | explicit scheme + system | Krylov | finish | outputs down
  float ms[6] = {0, 0, 0, 0, 0, 0};      // device time of those six intervals for the most recent call
};

// accessors implemented in ufe_diva.cu
struct ThkHandleView {
  DevMesh *dm; cudaStream_t st; int nranks, device;
  double *u_vav_b, *v_vav_b, *u_3D_b, *v_3D_b, *u_3D, *v_3D;
  bool sec_current;
  ThicknessState **slot;
};
int ufe_handle_thickness_view(ufe_handle *h, ThkHandleView *v);
int ufe_handle_gather_outputs(ufe_handle *h);
int ufe_build_operators_a_a(cudaStream_t st, const DevMesh &dm, DevFamily &F);

template <typename T>
static int talloc(T **p, size_t n) {
  *p = nullptr;
  UFE_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  UFE_CUDA(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
  UFE_CUDA(cudaStreamSynchronize(0));   // non-blocking handle stream: order the null-stream memset before any use
  return UFE_OK;
}

void ufe_thickness_free(ThicknessState *t) {
  if (!t) return;
  int *il[] = {t->VE, t->ETri, t->border, t->noice, t->bcmask, t->S.ptr, t->S.ind};
  for (int *p : il) cudaFree(p);
  double *dl[] = {t->A, t->Cw, t->Dx, t->Dy, t->D, t->Hi, t->Hb, t->SL, t->SMB, t->BMB, t->LMB, t->fm, t->target, t->bcHi,
                  t->divQ, t->dHi_dt, t->AMB, t->Hi_tp, t->Hs0, t->Hs1, t->u_b, t->v_b, t->dt_dev, t->Mval, t->S.val,
                  t->S.valS, t->S.bb, t->S.bS, t->S.x};
  for (double *p : dl) cudaFree(p);
  cudaFree(t->dtlim);
  for (double *p : t->vv_in) cudaFree(p);
  for (double *p : t->vv_dz) cudaFree(p);
  for (int *p : t->vv_mask) cudaFree(p);
  cudaFree(t->found_negative);
  cudaFree(t->w_3D); cudaFree(t->aa.ptr); cudaFree(t->aa.ind); cudaFree(t->aa.val[0]); cudaFree(t->aa.val[1]);
  for (cudaEvent_t e : t->ev) if (e) cudaEventDestroy(e);
  if (t->kw_alloc) ufe_krylov_free(t->kw);
  delete t;
}

// ------------------------------------------------------------------------------------
// device helpers (ice_geometry_basics.f90:28-41, 58-82; parameters.f90)
// ------------------------------------------------------------------------------------
#define THK_ICE_DENSITY 910.0
#define THK_SEAWATER_DENSITY 1028.0

__device__ __forceinline__ double ice_surface_elevation(double Hi, double Hb, double SL) {
  return Hi + fmax(SL - THK_ICE_DENSITY / THK_SEAWATER_DENSITY * Hi, Hb);
}
__device__ __forceinline__ double Hi_from_Hb_Hs_and_SL(double Hb, double Hs, double SL) {
  const double Hi_float = fmax(0.0, (SL - Hb) * (THK_SEAWATER_DENSITY / THK_ICE_DENSITY));
  const double Hs_float = Hb + Hi_float;
  if (Hs > Hs_float) return Hs - Hb;
  return fmin(Hi_float, (Hs - SL) / (1.0 - (THK_ICE_DENSITY / THK_SEAWATER_DENSITY)));
}

// pattern: ptr (1-based offsets), ind (1-based columns); row vi = [vi, C(vi,:)]
__global__ void k_thk_pattern(int nV, int nC_mem, const int *__restrict__ C, const int *__restrict__ nC,
                              const int *__restrict__ ptr, int *__restrict__ ind) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  int k = ptr[vi] - 1;
  ind[k++] = vi + 1;
  const int n = nC[vi];
  for (int ci = 0; ci < n; ci++) ind[k++] = C[(size_t)ci * nV + vi];
}

__global__ void k_thk_init(unsigned long long *dtlim, double dt_ice_max) {
  *dtlim = (unsigned long long)__double_as_longlong(dt_ice_max);
}

// M_divQ row, divQ, the explicit dH/dt and the flux-limited time step (min over vertices)
__global__ void __launch_bounds__(256)
k_thk_divq(ThkMesh M, const double *__restrict__ u_b, const double *__restrict__ v_b, const double *__restrict__ fm,
           const double *__restrict__ Hi, const double *__restrict__ SMB, const double *__restrict__ BMB,
           const double *__restrict__ LMB, const double *__restrict__ target, double Hi_min,
           double *__restrict__ val, double *__restrict__ divQ, double *__restrict__ dHi_dt, double *__restrict__ AMB,
           unsigned long long *dtlim) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  double lim = 1e300;
  if (vi < M.nV) {
    const int n = M.nC[vi];
    const int k0 = M.ptr[vi] - 1;
    const double A_i = M.A[vi];
    const bool out_ok = fm[vi] >= 1.0;
    double c0 = 0.0;
    for (int ci = 0; ci < n; ci++) {
      const size_t o = (size_t)ci * M.nV + vi;
      const int ei = M.VE[o] - 1;
      const int vj = M.C[o] - 1;
      const int til = M.ETri[ei], tir = M.ETri[M.nE + ei];
      double uc, vc;                                   // map_velocities_from_b_to_c_2D
      if (til == 0) { uc = u_b[tir - 1]; vc = v_b[tir - 1]; }
      else if (tir == 0) { uc = u_b[til - 1]; vc = v_b[til - 1]; }
      else { uc = (u_b[til - 1] + u_b[tir - 1]) / 2.0; vc = (v_b[til - 1] + v_b[tir - 1]) / 2.0; }
      const double L_c = M.Cw[o], D = M.D[o];
      const double u_perp = uc * M.Dx[o] / D + vc * M.Dy[o] / D;
      if (out_ok) c0 = c0 + L_c * fmax(0.0, u_perp) / A_i;
      double c = 0.0;
      if (fm[vj] >= 1.0) c = L_c * fmin(0.0, u_perp) / A_i;
      val[k0 + 1 + ci] = c;
    }
    val[k0] = c0;
    // multiply_CSR_matrix_with_vector_1D: entries in storage order
    double y = c0 * Hi[vi];
    for (int ci = 0; ci < n; ci++) y += val[k0 + 1 + ci] * Hi[M.C[(size_t)ci * M.nV + vi] - 1];
    divQ[vi] = y;
    const double d = -y + fm[vi] * (SMB[vi] + BMB[vi] - target[vi]) + LMB[vi];
    dHi_dt[vi] = d;
    AMB[vi] = d;
    // calc_flux_limited_timestep, as written: Hi / max(dHi_dt, 1e-9) for thinning ice
    if (Hi[vi] > Hi_min && d < 0.0) lim = Hi[vi] / fmax(d, 1e-9);
  }
  // block minimum, then one atomicMin on the bit pattern (positive doubles order like integers)
  for (int o = 16; o > 0; o >>= 1) lim = fmin(lim, __shfl_down_sync(0xffffffffu, lim, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = lim;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) lim = fmin(lim, sm[w]);
    if (lim < 1e300) atomicMin(dtlim, (unsigned long long)__double_as_longlong(fmax(lim, 0.0)));
  }
}

// mode 0: dt = min(dt_in, max(dt_ice_min, dt_lim)); Hi_tp = max(0, Hi + dHi_dt dt); Hs0 from Hi_tp.
// mode 1: Hs0 from the given Hi_tp only (second application of the border BCs).
__global__ void k_thk_hs(int nV, int mode, double dt_in, double dt_ice_min, const unsigned long long *dtlim,
                         double *dt_dev, const double *__restrict__ Hi, const double *__restrict__ dHi_dt,
                         const double *__restrict__ Hb, const double *__restrict__ SL, double *__restrict__ Hi_tp,
                         double *__restrict__ Hs0) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  double h;
  if (mode == 0) {
    const double dt_max = fmax(dt_ice_min, __longlong_as_double((long long)*dtlim));
    const double dt = fmin(dt_in, dt_max);
    if (vi == 0) *dt_dev = dt;
    h = fmax(0.0, Hi[vi] + dHi_dt[vi] * dt);
    Hi_tp[vi] = h;
  } else {
    h = Hi_tp[vi];
  }
  Hs0[vi] = ice_surface_elevation(h, Hb[vi], SL[vi]);
}

struct ThkBC { int bc[4]; };   // north, east, south, west
__device__ __forceinline__ int bc_of(const ThkBC &B, int vbi) { return B.bc[((vbi - 1) >> 1) & 3]; }   // 1,2 N; 3,4 E; 5,6 S; 7,8 W

__device__ __forceinline__ int n_interior_neighbours(int vi, int nV, const int *C, int n, const int *VBI, const int *noice) {
  int c = 0;
  for (int ci = 0; ci < n; ci++) {
    const int vj = C[(size_t)ci * nV + vi] - 1;
    if (VBI[vj] == 0 && !noice[vj]) c++;
  }
  return c;
}

// first pass over the border vertices: mean surface elevation of the interior neighbours
__global__ void k_thk_bc1(int nb, const int *__restrict__ border, int nV, const int *__restrict__ C,
                          const int *__restrict__ nC, const int *__restrict__ VBI, const int *__restrict__ noice,
                          ThkBC B, const double *__restrict__ Hb, const double *__restrict__ SL,
                          const double *__restrict__ Hs0, double *__restrict__ Hs1, double *__restrict__ Hi_tp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int vi = border[b];
  double hs = Hs0[vi];
  const int bc = bc_of(B, VBI[vi]);
  if (bc == UFE_BC_H_ZERO) {
    Hi_tp[vi] = 0.0;
  } else {
    const int n = nC[vi];
    const int nint = n_interior_neighbours(vi, nV, C, n, VBI, noice);
    if (nint > 0) {
      double s = 0.0;
      for (int ci = 0; ci < n; ci++) {
        const int vj = C[(size_t)ci * nV + vi] - 1;
        if (VBI[vj] == 0 && !noice[vj]) s += Hs0[vj];
      }
      hs = fmax(Hb[vi], s / (double)nint);
      Hi_tp[vi] = Hi_from_Hb_Hs_and_SL(Hb[vi], hs, SL[vi]);
    }
  }
  Hs1[vi] = hs;
}

// second pass: border vertices without interior neighbours take the mean over all neighbours of the
// surface elevations as they stand after the first pass (Hs1 on the border, Hs0 inside)
__global__ void k_thk_bc2(int nb, const int *__restrict__ border, int nV, const int *__restrict__ C,
                          const int *__restrict__ nC, const int *__restrict__ VBI, const int *__restrict__ noice,
                          ThkBC B, const double *__restrict__ Hb, const double *__restrict__ SL,
                          const double *__restrict__ Hs0, const double *__restrict__ Hs1, double *__restrict__ Hi_tp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int vi = border[b];
  const int bc = bc_of(B, VBI[vi]);
  if (bc == UFE_BC_H_ZERO) { Hi_tp[vi] = 0.0; return; }
  const int n = nC[vi];
  if (n_interior_neighbours(vi, nV, C, n, VBI, noice) != 0) return;
  double s = 0.0;
  for (int ci = 0; ci < n; ci++) {
    const int vj = C[(size_t)ci * nV + vi] - 1;
    s += (VBI[vj] > 0) ? Hs1[vj] : Hs0[vj];
  }
  const double hs = fmax(Hb[vi], s / (double)n);
  Hi_tp[vi] = Hi_from_Hb_Hs_and_SL(Hb[vi], hs, SL[vi]);
}

// explicit scheme, after the border BCs: prescribed thickness, no-ice mask, dH/dt, artificial mass balance
__global__ void k_thk_finish_explicit(int nV, const double *dt_dev, const int *__restrict__ bcmask,
                                      const double *__restrict__ bcHi, const int *__restrict__ noice,
                                      const double *__restrict__ Hi, double *__restrict__ Hi_tp,
                                      double *__restrict__ dHi_dt, double *__restrict__ AMB) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  double h = Hi_tp[vi];
  if (bcmask && bcmask[vi] == 1) h = fmax(0.0, bcHi[vi]);
  if (noice[vi]) h = 0.0;
  Hi_tp[vi] = h;
  const double d = (h - Hi[vi]) / *dt_dev;
  dHi_dt[vi] = d;
  AMB[vi] = d - AMB[vi];
}

// AA = 1 + dt f_s M_divQ with the BC rows, bb, and their point-Jacobi-scaled copies for the Krylov loop
__global__ void k_thk_system(int nV, const int *__restrict__ ptr, const int *__restrict__ VBI,
                             const int *__restrict__ bcmask, const double *__restrict__ bcHi,
                             const int *__restrict__ noice, double dt, double fs, const double *__restrict__ Mval,
                             const double *__restrict__ Hi, const double *__restrict__ divQ,
                             const double *__restrict__ fm, const double *__restrict__ SMB,
                             const double *__restrict__ BMB, const double *__restrict__ LMB,
                             const double *__restrict__ target, const double *__restrict__ Hi_ex,
                             double *__restrict__ val, double *__restrict__ valS, double *__restrict__ bb,
                             double *__restrict__ bS) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  const int k0 = ptr[vi] - 1, k1 = ptr[vi + 1] - 1;
  int kind = 0;
  double rhs = 0.0;
  if (VBI[vi] > 0) { kind = 1; rhs = Hi_ex[vi]; }                 // apply_ice_thickness_BC_matrix_domain_border
  if (bcmask && bcmask[vi] == 1) { kind = 1; rhs = bcHi[vi]; }    // ..._mask_prescribed_thickness
  if (noice[vi]) { kind = 1; rhs = 0.0; }                         // ..._mask_noice
  if (kind) {
    val[k0] = 1.0; valS[k0] = 1.0;
    for (int k = k0 + 1; k < k1; k++) { val[k] = 0.0; valS[k] = 0.0; }
    bb[vi] = rhs; bS[vi] = rhs;
    return;
  }
  double d = Mval[k0] * dt * fs + 1.0;
  const double r = Hi[vi] - (dt * (1.0 - fs) * divQ[vi]) +
                   fmax(-1.0 * Hi[vi], dt * (fm[vi] * (SMB[vi] + BMB[vi] - target[vi]) + LMB[vi]));
  val[k0] = d;
  bb[vi] = r;
  if (d == 0.0) d = 1.0;
  valS[k0] = val[k0] / d;
  bS[vi] = r / d;
  for (int k = k0 + 1; k < k1; k++) {
    const double a = Mval[k] * dt * fs;
    val[k] = a;
    valS[k] = a / d;
  }
}

__global__ void k_thk_finish_semi(int nV, double dt, const double *__restrict__ x, const double *__restrict__ Hi,
                                  double *__restrict__ Hi_tp, double *__restrict__ dHi_dt, double *__restrict__ AMB) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  const double h = x[vi];
  Hi_tp[vi] = h;
  const double a = (h - Hi[vi]) / dt;      // AMB = (Hi_tplusdt - Hi) / dt            :165
  const double d = (h - Hi[vi]) / dt;      // dHi_dt                                   :168
  dHi_dt[vi] = d;
  AMB[vi] = d - a;                         // AMB = dHi_dt - AMB                       :175
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct ThkCtx {
  ufe_handle *h; DevMesh *dm; cudaStream_t st; ThicknessState *t; double *u_res, *v_res;
  ThkHandleView hv;
};

static int thk_ctx(ufe_handle *h, ThkCtx &c, bool need_state) {
  if (!h) { ufe_set_error("null handle"); return UFE_ERR_INVALID; }
  UFE_TRY(ufe_handle_thickness_view(h, &c.hv));
  const int nranks = c.hv.nranks, device = c.hv.device;
  ThicknessState **slot = c.hv.slot;
  c.dm = c.hv.dm; c.st = c.hv.st; c.u_res = c.hv.u_vav_b; c.v_res = c.hv.v_vav_b;
  UFE_CUDA(cudaSetDevice(device));
  // Several ranks: this sub-path is replicated, not sharded -- every rank holds the full-length inputs (the C ABI
  // exchanges global arrays) and computes the whole update, bit-identically; the only collective is the gather that
  // makes the resident velocities of the last solve full-length on every rank.
  if (nranks > 1) UFE_TRY(ufe_handle_gather_outputs(h));
  c.h = h;
  c.t = *slot;
  if (need_state && !c.t) { ufe_set_error("ufe_mesh_set_edges has not been called on this handle"); return UFE_ERR_INVALID; }
  if (!c.t) { c.t = new ThicknessState(); *slot = c.t; }
  return UFE_OK;
}

extern "C" int ufe_mesh_set_edges(ufe_handle *h, const ufe_mesh_edges *e) {
  ThkCtx c;
  if (!e || !e->VE || !e->ETri || !e->A || !e->Cw || !e->D_x || !e->D_y || !e->D || e->nE <= 0) {
    ufe_set_error("ufe_mesh_set_edges: null / empty edge data"); return UFE_ERR_INVALID;
  }
  UFE_TRY(thk_ctx(h, c, false));
  ThicknessState *t = c.t;
  if (t->nV) { ufe_set_error("ufe_mesh_set_edges: edge data already set for this handle"); return UFE_ERR_INVALID; }
  const int nV = c.dm->nV, ncm = c.dm->nC_mem;
  const size_t nvc = (size_t)nV * ncm;
  // validate on the host: edge indices in range, each edge has at least one triangle
  for (size_t i = 0; i < (size_t)e->nE * 2; i++)
    if (e->ETri[i] < 0 || e->ETri[i] > c.dm->nTri) { ufe_set_error("ufe_mesh_set_edges: ETri out of range"); return UFE_ERR_INVALID; }
  for (int ei = 0; ei < e->nE; ei++)
    if (e->ETri[ei] == 0 && e->ETri[e->nE + ei] == 0) {
      ufe_set_error("something is seriously wrong with the ETri array of this mesh!");   // map_velocities_to_c_grid.f90:60
      return UFE_ERR_INVALID;
    }
  std::vector<int> hnC(nV), hVBI(nV), hptr(nV + 1), hborder;
  UFE_CUDA(cudaMemcpy(hnC.data(), c.dm->nC, sizeof(int) * nV, cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(hVBI.data(), c.dm->VBI, sizeof(int) * nV, cudaMemcpyDeviceToHost));
  hptr[0] = 1;
  for (int vi = 0; vi < nV; vi++) {
    for (int ci = 0; ci < hnC[vi]; ci++) {
      const int ei = e->VE[(size_t)ci * nV + vi];
      if (ei < 1 || ei > e->nE) { ufe_set_error("ufe_mesh_set_edges: VE out of range at vertex %d", vi + 1); return UFE_ERR_INVALID; }
    }
    if (!(e->A[vi] > 0.0)) { ufe_set_error("ufe_mesh_set_edges: non-positive Voronoi cell area at vertex %d", vi + 1); return UFE_ERR_INVALID; }
    hptr[vi + 1] = hptr[vi] + 1 + hnC[vi];
    if (hVBI[vi] > 0) hborder.push_back(vi);
  }
  t->nV = nV; t->nTri = c.dm->nTri; t->nE = e->nE; t->nnz = hptr[nV] - 1; t->nb = (int)hborder.size();
  auto up_i = [&](int **d, const int *src, size_t n) -> int {
    UFE_TRY(talloc(d, n));
    UFE_CUDA(cudaMemcpy(*d, src, sizeof(int) * n, cudaMemcpyHostToDevice));
    return UFE_OK;
  };
  auto up_d = [&](double **d, const double *src, size_t n) -> int {
    UFE_TRY(talloc(d, n));
    UFE_CUDA(cudaMemcpy(*d, src, sizeof(double) * n, cudaMemcpyHostToDevice));
    return UFE_OK;
  };
  UFE_TRY(up_i(&t->VE, e->VE, nvc)); UFE_TRY(up_i(&t->ETri, e->ETri, (size_t)e->nE * 2));
  UFE_TRY(up_i(&t->border, hborder.data(), hborder.size())); UFE_TRY(up_i(&t->S.ptr, hptr.data(), (size_t)nV + 1));
  UFE_TRY(up_d(&t->A, e->A, nV)); UFE_TRY(up_d(&t->Cw, e->Cw, nvc)); UFE_TRY(up_d(&t->Dx, e->D_x, nvc));
  UFE_TRY(up_d(&t->Dy, e->D_y, nvc)); UFE_TRY(up_d(&t->D, e->D, nvc));
  double **fl[] = {&t->Hi, &t->Hb, &t->SL, &t->SMB, &t->BMB, &t->LMB, &t->fm, &t->target, &t->bcHi, &t->divQ, &t->dHi_dt,
                   &t->AMB, &t->Hi_tp, &t->Hs0, &t->Hs1, &t->S.bb, &t->S.bS, &t->S.x};
  for (double **p : fl) UFE_TRY(talloc(p, (size_t)nV));
  UFE_TRY(talloc(&t->u_b, (size_t)t->nTri)); UFE_TRY(talloc(&t->v_b, (size_t)t->nTri));
  UFE_TRY(talloc(&t->noice, (size_t)nV)); UFE_TRY(talloc(&t->bcmask, (size_t)nV));
  UFE_TRY(talloc(&t->dtlim, 1)); UFE_TRY(talloc(&t->dt_dev, 1));
  UFE_TRY(talloc(&t->Mval, (size_t)t->nnz)); UFE_TRY(talloc(&t->S.val, (size_t)t->nnz)); UFE_TRY(talloc(&t->S.valS, (size_t)t->nnz));
  UFE_TRY(talloc(&t->S.ind, (size_t)t->nnz));
  for (cudaEvent_t &e : t->ev) UFE_CUDA(cudaEventCreate(&e));
  t->S.N = nV; t->S.m_loc = nV; t->S.r1 = 1; t->S.nnz = t->nnz; t->S.jmin = 1; t->S.jmax = nV;
  k_thk_pattern<<<ufe_div_up(nV, 256), 256, 0, c.st>>>(nV, ncm, c.dm->C, c.dm->nC, t->S.ptr, t->S.ind);
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaStreamSynchronize(c.st));
  return UFE_OK;
}

static ThkMesh thk_mesh(const ThkCtx &c) {
  const ThicknessState *t = c.t;
  return ThkMesh{t->nV, t->nTri, c.dm->nC_mem, t->nE, c.dm->C, c.dm->nC, c.dm->VBI, t->VE, t->ETri, t->S.ptr,
                 t->A, t->Cw, t->Dx, t->Dy, t->D};
}

static int thk_check(const ufe_thickness_config *cfg, const ufe_thickness_fields *f) {
  if (!cfg || !f) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  if (!f->Hi || !f->Hb || !f->SL || !f->SMB || !f->BMB || !f->LMB || !f->fraction_margin || !f->dHi_dt_target || !f->mask_noice) {
    ufe_set_error("ufe_thickness_fields: a required input field is NULL"); return UFE_ERR_INVALID;
  }
  if ((f->u_vav_b == nullptr) != (f->v_vav_b == nullptr)) { ufe_set_error("u_vav_b and v_vav_b must both be given or both be NULL"); return UFE_ERR_INVALID; }
  if ((f->BC_prescr_mask == nullptr) != (f->BC_prescr_Hi == nullptr)) {
    ufe_set_error("need to provide prescribed both Hi and mask!");   // conservation_of_mass_explicit.f90:113
    return UFE_ERR_INVALID;
  }
  for (int i = 0; i < 4; i++)
    if (cfg->BC_H[i] != UFE_BC_H_INFINITE && cfg->BC_H[i] != UFE_BC_H_ZERO) {
      ufe_set_error("unknown BC_H code %d", cfg->BC_H[i]);           // :199 crash('unknown BC_H ...')
      return UFE_ERR_INVALID;
    }
  return UFE_OK;
}

static int thk_upload(const ThkCtx &c, const ufe_thickness_fields *f, const double **u, const double **v) {
  ThicknessState *t = c.t;
  const size_t nb = sizeof(double) * (size_t)t->nV;
  UFE_CUDA(cudaEventRecord(t->ev[0], c.st));
  const double *src[] = {f->Hi, f->Hb, f->SL, f->SMB, f->BMB, f->LMB, f->fraction_margin, f->dHi_dt_target};
  double *dst[] = {t->Hi, t->Hb, t->SL, t->SMB, t->BMB, t->LMB, t->fm, t->target};
  for (int i = 0; i < 8; i++) UFE_CUDA(cudaMemcpyAsync(dst[i], src[i], nb, cudaMemcpyHostToDevice, c.st));
  UFE_CUDA(cudaMemcpyAsync(t->noice, f->mask_noice, sizeof(int) * (size_t)t->nV, cudaMemcpyHostToDevice, c.st));
  if (f->BC_prescr_mask) {
    UFE_CUDA(cudaMemcpyAsync(t->bcmask, f->BC_prescr_mask, sizeof(int) * (size_t)t->nV, cudaMemcpyHostToDevice, c.st));
    UFE_CUDA(cudaMemcpyAsync(t->bcHi, f->BC_prescr_Hi, nb, cudaMemcpyHostToDevice, c.st));
  }
  if (f->u_vav_b) {
    UFE_CUDA(cudaMemcpyAsync(t->u_b, f->u_vav_b, sizeof(double) * (size_t)t->nTri, cudaMemcpyHostToDevice, c.st));
    UFE_CUDA(cudaMemcpyAsync(t->v_b, f->v_vav_b, sizeof(double) * (size_t)t->nTri, cudaMemcpyHostToDevice, c.st));
    *u = t->u_b; *v = t->v_b;
  } else {
    *u = c.u_res; *v = c.v_res;
  }
  UFE_CUDA(cudaEventRecord(t->ev[1], c.st));
  return UFE_OK;
}

static int thk_download(const ThkCtx &c, ufe_thickness_fields *f) {
  ThicknessState *t = c.t;
  const size_t nb = sizeof(double) * (size_t)t->nV;
  if (f->AMB) UFE_CUDA(cudaMemcpyAsync(f->AMB, t->AMB, nb, cudaMemcpyDeviceToHost, c.st));
  if (f->dHi_dt) UFE_CUDA(cudaMemcpyAsync(f->dHi_dt, t->dHi_dt, nb, cudaMemcpyDeviceToHost, c.st));
  if (f->Hi_tplusdt) UFE_CUDA(cudaMemcpyAsync(f->Hi_tplusdt, t->Hi_tp, nb, cudaMemcpyDeviceToHost, c.st));
  if (f->divQ) UFE_CUDA(cudaMemcpyAsync(f->divQ, t->divQ, nb, cudaMemcpyDeviceToHost, c.st));
  UFE_CUDA(cudaEventRecord(t->ev[6], c.st));
  UFE_CUDA(cudaStreamSynchronize(c.st));
  for (int i = 0; i < 6; i++) UFE_CUDA(cudaEventElapsedTime(&t->ms[i], t->ev[i], t->ev[i + 1]));
  return UFE_OK;
}

// apply_ice_thickness_BC_explicit on t->Hi_tp (Hs0 must hold the surface elevation of Hi_tp)
static int thk_border_bcs(const ThkCtx &c, const ufe_thickness_config *cfg) {
  ThicknessState *t = c.t;
  if (t->nb == 0) return UFE_OK;
  ThkBC B;
  for (int i = 0; i < 4; i++) B.bc[i] = cfg->BC_H[i];
  k_thk_bc1<<<ufe_div_up(t->nb, 128), 128, 0, c.st>>>(t->nb, t->border, t->nV, c.dm->C, c.dm->nC, c.dm->VBI, t->noice, B,
                                                     t->Hb, t->SL, t->Hs0, t->Hs1, t->Hi_tp);
  UFE_LAUNCH_CHECK();
  k_thk_bc2<<<ufe_div_up(t->nb, 128), 128, 0, c.st>>>(t->nb, t->border, t->nV, c.dm->C, c.dm->nC, c.dm->VBI, t->noice, B,
                                                     t->Hb, t->SL, t->Hs0, t->Hs1, t->Hi_tp);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

// the explicit scheme on resident inputs; leaves dt in t->dt_dev
static int thk_explicit_resident(const ThkCtx &c, const ufe_thickness_config *cfg, const ufe_thickness_fields *f,
                                 const double *u, const double *v, double dt_in) {
  ThicknessState *t = c.t;
  const int nV = t->nV, g = ufe_div_up(nV, 256);
  k_thk_init<<<1, 1, 0, c.st>>>(t->dtlim, cfg->dt_ice_max);
  UFE_LAUNCH_CHECK();
  k_thk_divq<<<g, 256, 0, c.st>>>(thk_mesh(c), u, v, t->fm, t->Hi, t->SMB, t->BMB, t->LMB, t->target, cfg->Hi_min,
                                  t->Mval, t->divQ, t->dHi_dt, t->AMB, t->dtlim);
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaEventRecord(t->ev[2], c.st));
  t->have_M = true;
  k_thk_hs<<<g, 256, 0, c.st>>>(nV, 0, dt_in, cfg->dt_ice_min, t->dtlim, t->dt_dev, t->Hi, t->dHi_dt, t->Hb, t->SL,
                                t->Hi_tp, t->Hs0);
  UFE_LAUNCH_CHECK();
  UFE_TRY(thk_border_bcs(c, cfg));
  k_thk_finish_explicit<<<g, 256, 0, c.st>>>(nV, t->dt_dev, f->BC_prescr_mask ? t->bcmask : nullptr, t->bcHi, t->noice,
                                             t->Hi, t->Hi_tp, t->dHi_dt, t->AMB);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

extern "C" int ufe_calc_dHi_dt_explicit(ufe_handle *h, const ufe_thickness_config *cfg, ufe_thickness_fields *f, double *dt) {
  ThkCtx c;
  UFE_TRY(thk_check(cfg, f));
  if (!dt || !(*dt > 0.0)) { ufe_set_error("dt must be positive"); return UFE_ERR_INVALID; }
  UFE_TRY(thk_ctx(h, c, true));
  const double *u, *v;
  UFE_TRY(thk_upload(c, f, &u, &v));
  UFE_TRY(thk_explicit_resident(c, cfg, f, u, v, *dt));
  for (int i = 3; i <= 5; i++) UFE_CUDA(cudaEventRecord(c.t->ev[i], c.st));
  UFE_CUDA(cudaMemcpyAsync(dt, c.t->dt_dev, sizeof(double), cudaMemcpyDeviceToHost, c.st));
  return thk_download(c, f);
}

// the semi-implicit scheme on resident inputs (everything but the host copies)
static int thk_semi_resident(const ThkCtx &c, const ufe_thickness_config *cfg, const ufe_thickness_fields *f, const double *u,
                             const double *v, double dt, int32_t *n_Axb_its, int32_t *flags) {
  ThicknessState *t = c.t;
  const int nV = t->nV, g = ufe_div_up(nV, 256);
  // the explicit solution first (:111-115): M_divQ, divQ and the border values Hi_tplusdt_ex
  UFE_TRY(thk_explicit_resident(c, cfg, f, u, v, dt));
  // apply_ice_thickness_BC_matrix_domain_border applies the explicit border BCs to Hi_tplusdt_ex once more (:224)
  k_thk_hs<<<g, 256, 0, c.st>>>(nV, 1, dt, cfg->dt_ice_min, t->dtlim, t->dt_dev, t->Hi, t->dHi_dt, t->Hb, t->SL, t->Hi_tp, t->Hs0);
  UFE_LAUNCH_CHECK();
  UFE_TRY(thk_border_bcs(c, cfg));
  k_thk_system<<<g, 256, 0, c.st>>>(nV, t->S.ptr, c.dm->VBI, f->BC_prescr_mask ? t->bcmask : nullptr, t->bcHi, t->noice, dt,
                                    cfg->dHi_semiimplicit_fs, t->Mval, t->Hi, t->divQ, t->fm, t->SMB, t->BMB, t->LMB,
                                    t->target, t->Hi_tp, t->S.val, t->S.valS, t->S.bb, t->S.bS);
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaEventRecord(t->ev[3], c.st));
  t->have_system = true;
  if (!t->kw_alloc) { UFE_TRY(ufe_krylov_alloc(t->kw, nV, nV, true)); t->kw_alloc = true; }
  Comm comm;
  int its = 0, fl = 0;
  UFE_TRY(ufe_krylov_run(c.st, t->S, t->kw, comm, nullptr, cfg->krylov_method, cfg->dHi_PETSc_rtol, cfg->dHi_PETSc_abstol,
                         cfg->krylov_maxits > 0 ? cfg->krylov_maxits : 10000, 0, &its, &fl));
  if (n_Axb_its) *n_Axb_its = its;
  if (flags) *flags = fl;
  UFE_CUDA(cudaEventRecord(t->ev[4], c.st));
  k_thk_finish_semi<<<g, 256, 0, c.st>>>(nV, dt, t->S.x, t->Hi, t->Hi_tp, t->dHi_dt, t->AMB);
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaEventRecord(t->ev[5], c.st));
  return UFE_OK;
}

static int thk_check_semi(const ufe_thickness_config *cfg, double dt) {
  if (!(dt > 0.0)) { ufe_set_error("dt must be positive"); return UFE_ERR_INVALID; }
  if (cfg->krylov_method != UFE_KRYLOV_BICGSTAB && cfg->krylov_method != UFE_KRYLOV_GMRES) {
    ufe_set_error("unknown krylov_method %d", cfg->krylov_method); return UFE_ERR_INVALID;
  }
  return UFE_OK;
}

extern "C" int ufe_calc_dHi_dt_semiimplicit(ufe_handle *h, const ufe_thickness_config *cfg, ufe_thickness_fields *f,
                                            double dt, int32_t *n_Axb_its, int32_t *flags) {
  ThkCtx c;
  UFE_TRY(thk_check(cfg, f));
  UFE_TRY(thk_check_semi(cfg, dt));
  UFE_TRY(thk_ctx(h, c, true));
  const double *u, *v;
  UFE_TRY(thk_upload(c, f, &u, &v));
  UFE_TRY(thk_semi_resident(c, cfg, f, u, v, dt, n_Axb_its, flags));
  return thk_download(c, f);
}

// calc_dHi_dt (conservation_of_mass_main.f90:22-109) after the scheme: clip negative thicknesses, flag values below
// -0.1 m on ice thicker than Hi_min, AMB = AMB + (Hi_tplusdt - Hi)/dt - dHi_dt, dHi_dt = (Hi_tplusdt - Hi)/dt
__global__ void k_thk_post(int nV, const double *dt_dev, double dt_host, double Hi_min, const double *__restrict__ Hi,
                           double *__restrict__ Hi_tp, double *__restrict__ dHi_dt, double *__restrict__ AMB, int *found_negative) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  const double dt = dt_dev ? *dt_dev : dt_host;
  double h = Hi_tp[vi];
  if (h < 0.0) {
    if (h < -0.1 && Hi[vi] > Hi_min) *found_negative = 1;
    h = 0.0;
    Hi_tp[vi] = h;
  }
  const double d = (h - Hi[vi]) / dt;
  AMB[vi] = AMB[vi] + d - dHi_dt[vi];
  dHi_dt[vi] = d;
}

extern "C" int ufe_calc_dHi_dt(ufe_handle *h, const ufe_thickness_config *cfg, int32_t choice_ice_integration_method,
                               ufe_thickness_fields *f, double *dt, int32_t *n_Axb_its, int32_t *flags) {
  ThkCtx c;
  UFE_TRY(thk_check(cfg, f));
  if (!dt) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  if (choice_ice_integration_method < UFE_THK_NONE || choice_ice_integration_method > UFE_THK_SEMI_IMPLICIT) {
    ufe_set_error("unknown choice_ice_integration_method code %d!", choice_ice_integration_method);   // :62
    return UFE_ERR_INVALID;
  }
  if (choice_ice_integration_method == UFE_THK_SEMI_IMPLICIT) UFE_TRY(thk_check_semi(cfg, *dt));
  else if (!(*dt > 0.0)) { ufe_set_error("dt must be positive"); return UFE_ERR_INVALID; }
  UFE_TRY(thk_ctx(h, c, true));
  ThicknessState *t = c.t;
  const int nV = t->nV, g = ufe_div_up(nV, 256);
  if (n_Axb_its) *n_Axb_its = 0;
  if (flags) *flags = 0;
  const double *u, *v;
  UFE_TRY(thk_upload(c, f, &u, &v));
  if (choice_ice_integration_method == UFE_THK_NONE) {          // :64-68: unchanging geometry
    const size_t nb = sizeof(double) * (size_t)nV;
    UFE_CUDA(cudaMemcpyAsync(t->Hi_tp, t->Hi, nb, cudaMemcpyDeviceToDevice, c.st));
    UFE_CUDA(cudaMemsetAsync(t->dHi_dt, 0, nb, c.st));
    UFE_CUDA(cudaMemsetAsync(t->AMB, 0, nb, c.st));
    for (int i = 2; i <= 5; i++) UFE_CUDA(cudaEventRecord(t->ev[i], c.st));
    ufe_thickness_fields g2 = *f;
    g2.divQ = nullptr;                                          // divQ is not computed on this branch
    return thk_download(c, &g2);
  }
  if (!t->found_negative) UFE_TRY(talloc(&t->found_negative, 1));
  UFE_CUDA(cudaMemsetAsync(t->found_negative, 0, sizeof(int), c.st));
  const bool semi = choice_ice_integration_method == UFE_THK_SEMI_IMPLICIT;
  if (semi) {
    UFE_TRY(thk_semi_resident(c, cfg, f, u, v, *dt, n_Axb_its, flags));
  } else {
    UFE_TRY(thk_explicit_resident(c, cfg, f, u, v, *dt));
    for (int i = 3; i <= 4; i++) UFE_CUDA(cudaEventRecord(t->ev[i], c.st));
  }
  k_thk_post<<<g, 256, 0, c.st>>>(nV, semi ? nullptr : t->dt_dev, *dt, cfg->Hi_min, t->Hi, t->Hi_tp, t->dHi_dt, t->AMB,
                                  t->found_negative);
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaEventRecord(t->ev[5], c.st));
  int neg = 0;
  if (!semi) UFE_CUDA(cudaMemcpyAsync(dt, t->dt_dev, sizeof(double), cudaMemcpyDeviceToHost, c.st));
  UFE_CUDA(cudaMemcpyAsync(&neg, t->found_negative, sizeof(int), cudaMemcpyDeviceToHost, c.st));
  UFE_TRY(thk_download(c, f));
  if (neg && flags) *flags |= UFE_FLAG_NEGATIVE_HI;             // warning('encountered negative values for Hi_tplusdt ...') :91
  return UFE_OK;
}

extern "C" int ufe_get_thickness_csr(ufe_handle *h, int32_t which, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind,
                                     double *val, double *bb) {
  ThkCtx c;
  UFE_TRY(thk_ctx(h, c, true));
  ThicknessState *t = c.t;
  if (which != 0 && which != 1) { ufe_set_error("which must be 0 (M_divQ) or 1 (AA)"); return UFE_ERR_INVALID; }
  if ((which == 0 && !t->have_M) || (which == 1 && !t->have_system)) { ufe_set_error("no matrix has been built yet"); return UFE_ERR_INVALID; }
  if (m_loc) *m_loc = t->nV;
  if (nnz) *nnz = t->nnz;
  if (!ind) return UFE_OK;
  UFE_CUDA(cudaStreamSynchronize(c.st));
  if (ptr) UFE_CUDA(cudaMemcpy(ptr, t->S.ptr, sizeof(int) * ((size_t)t->nV + 1), cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(ind, t->S.ind, sizeof(int) * (size_t)t->nnz, cudaMemcpyDeviceToHost));
  if (val) UFE_CUDA(cudaMemcpy(val, which == 0 ? t->Mval : t->S.val, sizeof(double) * (size_t)t->nnz, cudaMemcpyDeviceToHost));
  if (bb && which == 1) UFE_CUDA(cudaMemcpy(bb, t->S.bb, sizeof(double) * (size_t)t->nV, cudaMemcpyDeviceToHost));
  return UFE_OK;
}

extern "C" int ufe_get_thickness_timing(ufe_handle *h, double ms[6], double *divq_algorithmic_bytes) {
  ThkCtx c;
  UFE_TRY(thk_ctx(h, c, true));
  if (!ms) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  for (int i = 0; i < 6; i++) ms[i] = c.t->ms[i];
  // k_thk_divq: per connection VE 4 + C 4 + ETri 8 + (u,v) of two triangles 32 + Cw,D_x,D_y,D 32 + fraction_margin, Hi
  // gathers 16 + val out 8 = 104 B; per vertex ptr 4 + nC 4 + A 8 + 6 fields in 48 + diagonal 8 + 3 fields out 24 = 96 B
  if (divq_algorithmic_bytes) *divq_algorithmic_bytes = 104.0 * (c.t->nnz - c.t->nV) + 96.0 * c.t->nV;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// calc_vertical_velocities (vertical_velocities.f90:18-210): w from conservation of mass in each 3-D
// Voronoi cell.  One thread per vertex; the loop over connections is the outer loop and the loop over
// layers the inner one, which keeps the reference's summation order (connections 1..nC) for every
// layer while each edge's geometry is read once.  3-D fields are (n, nz) column-major.
// ------------------------------------------------------------------------------------
template <int NZ>
__global__ void __launch_bounds__(128)
k_vertical_velocities(ThkMesh M, const double *__restrict__ V, const double *__restrict__ zeta, const int *__restrict__ aptr,
                      const int *__restrict__ aind, const double *__restrict__ addx, const double *__restrict__ addy,
                      const double *__restrict__ Hib, const double *__restrict__ dHb_dt, const double *__restrict__ dHi_dt,
                      const double *__restrict__ BMB, const double *__restrict__ Hi, const int *__restrict__ mask_gr,
                      const int *__restrict__ mask_fl, const double *__restrict__ zx, const double *__restrict__ zy,
                      const double *__restrict__ zz, const double *__restrict__ u3b, const double *__restrict__ v3b,
                      const double *__restrict__ u3, const double *__restrict__ v3, double *__restrict__ w) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= M.nV) return;
  const size_t nV = M.nV, nT = M.nTri;
  const bool gr = mask_gr[vi] != 0, fl = mask_fl[vi] != 0;
  if (!(gr || fl)) {
#pragma unroll
    for (int k = 0; k < NZ; k++) w[k * nV + vi] = 0.0;
    return;
  }
  double dHib_dt;
  if (gr) dHib_dt = dHb_dt[vi];
  else dHib_dt = -dHi_dt[vi] * THK_ICE_DENSITY / THK_SEAWATER_DENSITY;
  // slopes of the ice base: rows of M_ddx_a_a / M_ddy_a_a
  double sx = 0.0, sy = 0.0;
  for (int k = aptr[vi] - 1; k < aptr[vi + 1] - 1; k++) {
    const double hb = Hib[aind[k] - 1];
    sx += addx[k] * hb; sy += addy[k] * hb;
  }
  const double wb = (u3[(NZ - 1) * nV + vi] * sx) + (v3[(NZ - 1) * nV + vi] * sy) + dHib_dt + fmin(0.0, BMB[vi]);
  if (Hi[vi] < 10.0) {
#pragma unroll
    for (int k = 0; k < NZ; k++) w[k * nV + vi] = wb;
    return;
  }
  double cint[NZ - 1];
#pragma unroll
  for (int k = 0; k < NZ - 1; k++) cint[k] = 0.0;
  const int n = M.nC[vi];
  const double x0 = V[vi], y0 = V[nV + vi];
  for (int ci = 0; ci < n; ci++) {
    const size_t o = (size_t)ci * nV + vi;
    const int vj = M.C[o] - 1, ei = M.VE[o] - 1;
    const int til = M.ETri[ei], tir = M.ETri[M.nE + ei];
    const double dS = M.Cw[o];
    double n0 = V[vj] - x0, n1 = V[nV + vj] - y0;
    const double nn = sqrt(n0 * n0 + n1 * n1);
    n0 = n0 / nn; n1 = n1 / nn;
    double up = 0.0, vp = 0.0;     // edge velocity of the layer below (k+1), carried through the loop
#pragma unroll
    for (int k = NZ - 1; k >= 0; k--) {
      double uc, vc;               // map_velocities_from_b_to_c_3D
      if (til == 0) { uc = u3b[k * nT + tir - 1]; vc = v3b[k * nT + tir - 1]; }
      else if (tir == 0) { uc = u3b[k * nT + til - 1]; vc = v3b[k * nT + til - 1]; }
      else { uc = (u3b[k * nT + til - 1] + u3b[k * nT + tir - 1]) / 2.0; vc = (v3b[k * nT + til - 1] + v3b[k * nT + tir - 1]) / 2.0; }
      if (k < NZ - 1) {
        const double u_ks = 0.5 * (uc + up), v_ks = 0.5 * (vc + vp);
        cint[k] = cint[k] + (u_ks * n0 + v_ks * n1) * dS;
      }
      up = uc; vp = vc;
    }
  }
  const double A_i = M.A[vi];
  double wk = wb;
  w[(NZ - 1) * nV + vi] = wb;
  double u_below = u3[(NZ - 1) * nV + vi], v_below = v3[(NZ - 1) * nV + vi];
  double zx_b = zx[(NZ - 1) * nV + vi], zy_b = zy[(NZ - 1) * nV + vi], zz_b = zz[(NZ - 1) * nV + vi];
#pragma unroll
  for (int ks = NZ - 2; ks >= 0; ks--) {
    const double dzeta = zeta[ks + 1] - zeta[ks];
    const double grad_uv = cint[ks] / A_i;
    const double u_k = u3[ks * nV + vi], v_k = v3[ks * nV + vi];
    const double du = (u_below - u_k) / dzeta, dv = (v_below - v_k) / dzeta;
    const double zx_k = zx[ks * nV + vi], zy_k = zy[ks * nV + vi], zz_k = zz[ks * nV + vi];
    const double dzx = 0.5 * (zx_k + zx_b), dzy = 0.5 * (zy_k + zy_b), dzz = 0.5 * (zz_k + zz_b);
    const double dw = -1.0 / dzz * (grad_uv + dzx * du + dzy * dv);
    wk = wk - dzeta * dw;
    w[ks * nV + vi] = wk;
    u_below = u_k; v_below = v_k; zx_b = zx_k; zy_b = zy_k; zz_b = zz_k;
  }
}

// first use: build M_ddx_a_a / M_ddy_a_a on the device and allocate the work fields
static int thk_ensure_vv(const ThkCtx &c) {
  ThicknessState *t = c.t;
  if (t->aa_built) return UFE_OK;
  const size_t nV = t->nV, nz = c.dm->nz;
  UFE_TRY(ufe_build_operators_a_a(c.st, *c.dm, t->aa));
  for (double *&p : t->vv_in) UFE_TRY(talloc(&p, nV));
  for (double *&p : t->vv_dz) UFE_TRY(talloc(&p, nV * nz));
  for (int *&p : t->vv_mask) UFE_TRY(talloc(&p, nV));
  UFE_TRY(talloc(&t->w_3D, nV * nz));
  t->aa_built = true;
  return UFE_OK;
}

template <int NZ>
static void launch_vv(const ThkCtx &c, const ThkMesh &M) {
  ThicknessState *t = c.t;
  k_vertical_velocities<NZ><<<ufe_div_up(M.nV, 128), 128, 0, c.st>>>(
      M, c.dm->V, c.dm->zeta, t->aa.ptr, t->aa.ind, t->aa.val[0], t->aa.val[1], t->vv_in[0], t->vv_in[1], t->vv_in[2],
      t->vv_in[3], t->vv_in[4], t->vv_mask[0], t->vv_mask[1], t->vv_dz[0], t->vv_dz[1], t->vv_dz[2], c.hv.u_3D_b, c.hv.v_3D_b,
      c.hv.u_3D, c.hv.v_3D, t->w_3D);
}

extern "C" int ufe_calc_vertical_velocities(ufe_handle *h, const ufe_vertical_velocity_inputs *in, double *w_3D) {
  ThkCtx c;
  if (!in || !w_3D) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  if (!in->Hi || !in->Hib || !in->dHb_dt || !in->dHi_dt || !in->BMB || !in->mask_grounded_ice || !in->mask_floating_ice ||
      !in->dzeta_dx_ak || !in->dzeta_dy_ak || !in->dzeta_dz_ak) {
    ufe_set_error("ufe_vertical_velocity_inputs: a required input field is NULL"); return UFE_ERR_INVALID;
  }
  UFE_TRY(thk_ctx(h, c, true));
  if (!c.hv.u_3D || !c.hv.sec_current) {
    ufe_set_error("calc_vertical_velocities reads ice%%u_3D / v_3D: call ufe_calc_secondary_velocities after the velocity solve first");
    return UFE_ERR_INVALID;
  }
  ThicknessState *t = c.t;
  const size_t nV = t->nV, nz = c.dm->nz;
  UFE_TRY(thk_ensure_vv(c));
  const double *src[5] = {in->Hib, in->dHb_dt, in->dHi_dt, in->BMB, in->Hi};
  for (int i = 0; i < 5; i++) UFE_CUDA(cudaMemcpyAsync(t->vv_in[i], src[i], 8 * nV, cudaMemcpyHostToDevice, c.st));
  const double *dz[3] = {in->dzeta_dx_ak, in->dzeta_dy_ak, in->dzeta_dz_ak};
  for (int i = 0; i < 3; i++) UFE_CUDA(cudaMemcpyAsync(t->vv_dz[i], dz[i], 8 * nV * nz, cudaMemcpyHostToDevice, c.st));
  UFE_CUDA(cudaMemcpyAsync(t->vv_mask[0], in->mask_grounded_ice, 4 * nV, cudaMemcpyHostToDevice, c.st));
  UFE_CUDA(cudaMemcpyAsync(t->vv_mask[1], in->mask_floating_ice, 4 * nV, cudaMemcpyHostToDevice, c.st));
  const ThkMesh M = thk_mesh(c);
  switch ((int)nz) {
    case 4: launch_vv<4>(c, M); break;
    case 8: launch_vv<8>(c, M); break;
    case 10: launch_vv<10>(c, M); break;
    case 12: launch_vv<12>(c, M); break;
    case 16: launch_vv<16>(c, M); break;
    case 20: launch_vv<20>(c, M); break;
    case 24: launch_vv<24>(c, M); break;
    case 32: launch_vv<32>(c, M); break;
    default: ufe_set_error("calc_vertical_velocities: nz = %d is not instantiated (4, 8, 10, 12, 16, 20, 24, 32)", (int)nz); return UFE_ERR_INVALID;
  }
  UFE_LAUNCH_CHECK();
  UFE_CUDA(cudaMemcpyAsync(w_3D, t->w_3D, 8 * nV * nz, cudaMemcpyDeviceToHost, c.st));
  UFE_CUDA(cudaStreamSynchronize(c.st));
  return UFE_OK;
}

// M_ddx_a_a / M_ddy_a_a as built on the device (which: 0 ddx, 1 ddy); query sizes with ind == NULL
extern "C" int ufe_mesh_get_operator_a_a(ufe_handle *h, int32_t which, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind,
                                         double *val) {
  ThkCtx c;
  UFE_TRY(thk_ctx(h, c, true));
  ThicknessState *t = c.t;
  if (which < 0 || which > 1) { ufe_set_error("bad operator id"); return UFE_ERR_INVALID; }
  UFE_TRY(thk_ensure_vv(c));
  if (m_loc) *m_loc = t->aa.m_loc;
  if (nnz) *nnz = t->aa.nnz;
  if (!ind) return UFE_OK;
  if (ptr) UFE_CUDA(cudaMemcpy(ptr, t->aa.ptr, sizeof(int) * ((size_t)t->aa.m_loc + 1), cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(ind, t->aa.ind, sizeof(int) * (size_t)t->aa.nnz, cudaMemcpyDeviceToHost));
  if (val) UFE_CUDA(cudaMemcpy(val, t->aa.val[which], sizeof(double) * (size_t)t->aa.nnz, cudaMemcpyDeviceToHost));
  return UFE_OK;
}
