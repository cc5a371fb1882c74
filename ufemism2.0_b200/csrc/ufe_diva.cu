// Handle, Picard drivers and the C ABI of include/ufe_diva.h.
// solve_DIVA  : src/UFEMISM/ice_dynamics/conservation_of_momentum/SSA_DIVA/DIVA_main.f90:88-262
// solve_SSA   : .../SSA_main.f90:87-242
// solve_SSA_DIVA_linearised : .../solve_linearised_SSA_DIVA.f90:23-178
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ufe_closures.cuh"

int ufe_build_operators(cudaStream_t st, const DevMesh &dm, int vi1, int vi2, int ti1, int ti2, bool need[3],
                        DevFamily fam[3]);
int ufe_colrange(cudaStream_t st, int nnz, const int *ind, int *jmin, int *jmax);
int ufe_kspmv_plain(cudaStream_t st, const DevSystem &S, const double *xg, double *y, KrylovWork &kw);
int ufe_kspmv_only(cudaStream_t st, const DevSystem &S, const double *xg, double *y, KrylovWork &kw);
static int make_plan(struct ufe_handle *h, HaloPlan &plan, int ntot, int own_lo, int own_hi, int need_lo, int need_hi);

// ------------------------------------------------------------------------------------
// errors, counters
// ------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
thread_local int64_t g_launch_count = 0;

void ufe_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
extern "C" const char *ufe_last_error_string(void) { return g_err; }
extern "C" int ufe_version(void) { return 200; }
extern "C" int ufe_sizeof_solve_info(void) { return (int)sizeof(ufe_solve_info); }

extern "C" void ufe_partition_list(int32_t ntot, int32_t i, int32_t n, int32_t *i1, int32_t *i2) {
  // mpi_distributed_memory.f90:42-68
  if (ntot > n * 2) {
    const int rem = ntot % n, slice = ntot / n;
    *i1 = slice * i + std::min(i, rem) + 1;
    *i2 = slice * (i + 1) + std::min(i + 1, rem);
  } else if (i == 0) { *i1 = 1; *i2 = ntot; }
  else { *i1 = 1; *i2 = 0; }
}

extern "C" int ufe_comm_get_unique_id(char id_out[128]) {
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  UFE_NCCL(ncclGetUniqueId(&id));
  memcpy(id_out, &id, 128);
  return UFE_OK;
}

static int check_device() {
  int dev = 0;
  UFE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  UFE_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    ufe_set_error("libufe_diva is built for sm_100a only; device %d is sm_%d%d (no fallback path)", dev, p.major, p.minor);
    return UFE_ERR_CUDA;
  }
  return UFE_OK;
}

template <typename T>
static int dalloc(T **p, size_t n) {
  *p = nullptr;
  UFE_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  // the handle's stream is non-blocking: a null-stream memset is not ordered before work on it, so wait here
  UFE_CUDA(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
  UFE_CUDA(cudaStreamSynchronize(0));
  return UFE_OK;
}
template <typename T>
static int dupload(T **p, const T *h, size_t n) {
  *p = nullptr;
  UFE_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  if (n) UFE_CUDA(cudaMemcpy(*p, h, sizeof(T) * n, cudaMemcpyHostToDevice));
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------
struct ufe_handle {
  ufe_config cfg;
  DevMesh dm;
  std::vector<double> hV, hGC, hzeta;          // host copies for the BC-table walk
  std::vector<int> hC, hnC, hiTri, hniTri, hTriBI;
  int vi1 = 1, vi2 = 0, ti1 = 1, ti2 = 0;      // ownership (1-based inclusive), partition_list
  Comm comm;
  int device = 0;
  cudaStream_t st = nullptr;
  DevFamily fam[3];                            // a_b, b_a, b_b(2nd)
  HaloPlan plan_b_for_a;                       // triangles needed by owned vertices (b_a rows)
  HaloPlan plan_a_for_b;                       // vertices needed by owned triangles (a_b rows)
  HaloPlan plan_b_for_b;                       // triangles needed by owned stiffness rows
  bool need_allgather_prev = false;            // copy BCs read u_prev anywhere
  // inputs (full length)
  double *Hi = nullptr, *Hs = nullptr, *Hib = nullptr, *SL = nullptr, *fraction_gr = nullptr,
         *fraction_gr_b = nullptr, *Neff = nullptr, *Ti = nullptr, *phi = nullptr, *alpha_sq = nullptr,
         *beta_sq = nullptr, *tys = nullptr, *bc_u = nullptr, *bc_v = nullptr;
  int *mask_gr = nullptr, *mask_fl = nullptr, *mask_land = nullptr, *bc_mask = nullptr;
  bool have_bc_prescr = false;
  int grounded_ice_exists = 1;
  DivaFields F;
  std::vector<double *> owned_ptrs;            // everything in F, for freeing
  // BC copy tables
  int *bc_slot = nullptr, *bc_copy_ti = nullptr;
  double *bc_copy_w = nullptr;
  // linear system
  DevSystem S;
  int *rowkind = nullptr;
  bool pattern_valid = false;
  KrylovWork kw;
  double *sym = nullptr;                       // symmetric peer buffer (several ranks): owns S.x, kw.pg, kw.sg
  SecondaryFields sec;                         // calc_secondary_velocities outputs (allocated on first use)
  bool sec_alloc = false, sec_current = false;
  int redundant_ranks = 0;                     // > 0: this many ranks each solve the whole (small) system redundantly
  bool outputs_gathered = false;               // several ranks: the solution fields are full-length on every rank
  int last_is_diva = 1;
  PcLU *pclu = nullptr;                        // UFE_PC_BJACOBI_LU workspace (tied to the cached pattern)
  int pc_used = -1;                            // resolved preconditioner for the cached pattern (-1 = undecided)
  int pc_used_l0 = -1;                         // ... of the last ufe_solve_matrix_equation_CSR call
  double *op_x = nullptr, *op_y = nullptr;     // staging buffers of ufe_mesh_apply_operator, grown on demand and kept
  size_t op_x_n = 0, op_y_n = 0;
  int pc_age = -1, pc_last_its = 0;            // bjacobi_lu reuse: solves since the last factorisation (-1 = never), its of the last solve
  int64_t pc_factorisations = 0;
  // reductions for the Picard residual
  double *red_partials = nullptr, *red_out = nullptr;
  unsigned *red_counter = nullptr;
  double *red_host = nullptr;                  // pinned
  cudaEvent_t ev[8];
  void *flush_buf = nullptr;
  size_t flush_bytes = 0;
  struct ThicknessState *thk = nullptr;        // ice-thickness path (ufe_thickness.cu), created by ufe_mesh_set_edges
};

// ufe_thickness.cu reaches the mesh, the stream and the resident velocities through this view
struct ThicknessState;
void ufe_thickness_free(ThicknessState *t);
struct ThkHandleView {
  DevMesh *dm; cudaStream_t st; int nranks, device;
  double *u_vav_b, *v_vav_b, *u_3D_b, *v_3D_b, *u_3D, *v_3D;   // u_3D / v_3D: nullptr until calc_secondary_velocities ran
  bool sec_current;                                             // ... for the most recent solve
  ThicknessState **slot;
};
int ufe_handle_thickness_view(ufe_handle *h, ThkHandleView *v) {
  v->dm = &h->dm; v->st = h->st; v->nranks = h->comm.nranks; v->device = h->device;
  v->u_vav_b = h->F.u_vav_b; v->v_vav_b = h->F.v_vav_b; v->u_3D_b = h->F.u_3D_b; v->v_3D_b = h->F.v_3D_b;
  v->u_3D = h->sec_alloc ? h->sec.u_3D : nullptr; v->v_3D = h->sec_alloc ? h->sec.v_3D : nullptr;
  v->sec_current = h->sec_current;
  v->slot = &h->thk;
  return UFE_OK;
}

static ClosureParams make_params(const ufe_handle *h, double eps_sq_0_applied) {
  ClosureParams P;
  memset(&P, 0, sizeof P);
  const ufe_config &c = h->cfg;
  for (int k = 0; k < h->dm.nz; k++) P.zeta[k] = h->hzeta[k];
  P.visc_eff_min = c.visc_eff_min;
  const double A_min = 1e-18, n = c.Glens_flow_law_exponent;
  // DIVA_main.f90:431: eta_max depends on the *applied* eps0^2
  P.eta_max = 0.5 * pow(A_min, -1.0 / n) * pow(eps_sq_0_applied, (1.0 - n) / (2.0 * n));
  P.eps_sq_0 = eps_sq_0_applied; P.n_Glen = n; P.vel_max = c.vel_max;
  P.uniform_A = c.uniform_Glens_flow_factor; P.m_enh_sheet = c.m_enh_sheet; P.m_enh_shelf = c.m_enh_shelf;
  P.slid_delta_v = c.slid_delta_v; P.slid_beta_max = c.slid_beta_max; P.slid_Weertman_m = c.slid_Weertman_m;
  P.slid_Budd_q = c.slid_Budd_q_plastic; P.slid_Budd_u = c.slid_Budd_u_threshold;
  P.slid_ZI_p = c.slid_ZI_p; P.slid_ZI_ut = c.slid_ZI_ut;
  P.subgrid_exponent = c.subgrid_friction_exponent_on_B_grid; P.Hi_min = c.Hi_min;
  P.icestream_Hi = c.refgeo_idealised_SSA_icestream_Hi; P.icestream_dhdx = c.refgeo_idealised_SSA_icestream_dhdx;
  P.icestream_L = c.refgeo_idealised_SSA_icestream_L; P.icestream_m = c.refgeo_idealised_SSA_icestream_m;
  P.ISMIP_HOM_L = c.refgeo_idealised_ISMIP_HOM_L;
  P.rheology = c.choice_ice_rheology_Glen; P.enh_transition = c.choice_enhancement_factor_transition;
  P.sliding_law = c.choice_sliding_law; P.idealised_law = c.choice_idealised_sliding_law;
  P.do_GL_subgrid_friction = c.do_GL_subgrid_friction;
  return P;
}

static AssemblyParams make_asm_params(const ufe_handle *h) {
  AssemblyParams A;
  A.crossterms = h->cfg.do_include_SSADIVA_crossterms;
  A.pc = h->cfg.krylov_pc >= UFE_PC_BJACOBI_LU ? UFE_PC_BJACOBI2 : h->cfg.krylov_pc;   // scaling folded into the matrix
  for (int s = 0; s < 4; s++) { A.bc_u[s] = h->cfg.BC_u[s]; A.bc_v[s] = h->cfg.BC_v[s]; }
  A.visc_it_relax = h->cfg.visc_it_relax;
  return A;
}

static int validate_config(const ufe_config *c) {
  for (int s = 0; s < 4; s++) {
    if (c->BC_u[s] < 1 || c->BC_u[s] > 4) { ufe_set_error("unknown choice_BC_u (code %d)!", c->BC_u[s]); return UFE_ERR_INVALID; }
    // solve_linearised_SSA_DIVA.f90:586-637 has no 'infinite_SSA_icestream' case for v: the reference crashes there
    if (c->BC_v[s] == UFE_BC_INFINITE_SSA_ICESTREAM) { ufe_set_error("unknown choice_BC_v \"infinite_SSA_icestream\"!"); return UFE_ERR_INVALID; }
    if (c->BC_v[s] < 1 || c->BC_v[s] > 3) { ufe_set_error("unknown choice_BC_v (code %d)!", c->BC_v[s]); return UFE_ERR_INVALID; }
  }
  if (c->choice_sliding_law < 0 || c->choice_sliding_law > 7) { ufe_set_error("unknown choice_sliding_law (code %d)", c->choice_sliding_law); return UFE_ERR_INVALID; }
  if (c->choice_sliding_law == UFE_SLID_IDEALISED &&
      !(c->choice_idealised_sliding_law == 1 || c->choice_idealised_sliding_law == 2 ||
        c->choice_idealised_sliding_law == 3 || c->choice_idealised_sliding_law == 5)) {
    ufe_set_error("unknown choice_idealised_sliding_law (code %d)", c->choice_idealised_sliding_law); return UFE_ERR_INVALID;
  }
  if (c->choice_ice_rheology_Glen < 0 || c->choice_ice_rheology_Glen > 1) { ufe_set_error("unknown choice_ice_rheology_Glen (code %d)!", c->choice_ice_rheology_Glen); return UFE_ERR_INVALID; }
  if (c->choice_enhancement_factor_transition < 0 || c->choice_enhancement_factor_transition > 1) { ufe_set_error("unknown choice_enhancement_factor_transition!"); return UFE_ERR_INVALID; }
  if (c->do_subgrid_friction_on_A_grid) { ufe_set_error("do_subgrid_friction_on_A_grid = .true. is not supported (needs Hs_slope and grounding-line masks)"); return UFE_ERR_INVALID; }
  if (c->krylov_method < 0 || c->krylov_method > 1 || c->krylov_pc < 0 || c->krylov_pc > UFE_PC_ND_LU) { ufe_set_error("unknown krylov method / preconditioner"); return UFE_ERR_INVALID; }
  if (c->choice_sliding_law == UFE_SLID_IDEALISED && c->choice_idealised_sliding_law == UFE_IDEAL_SSA_ICESTREAM &&
      c->Glens_flow_law_exponent != 3.0) { ufe_set_error("Schoof only derived a solution for the case of n=3!"); return UFE_ERR_INVALID; }
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// BC copy tables on the host (find_ti_copy_*, src/UPSY/mesh/mesh_utilities.f90:2623-2730;
// find_containing_vertex :1368-1412).  Purely geometric, so built once per mesh instead
// of once per BC row per Picard iteration.
// ------------------------------------------------------------------------------------
static double h_norm2(double a, double b) {
  double scale = 1.0, ssq = 0.0, v[2] = {a, b};
  for (int i = 0; i < 2; i++)
    if (v[i] != 0.0) {
      const double ax = fabs(v[i]);
      if (scale < ax) { const double t = scale / ax; ssq = 1.0 + ssq * t * t; scale = ax; }
      else { const double t = ax / scale; ssq += t * t; }
    }
  return scale * sqrt(ssq);
}

static int build_bc_tables(ufe_handle *h) {
  const int nV = h->dm.nV, nTri = h->dm.nTri, ncm = h->dm.nC_mem;
  bool any_copy = false;
  for (int s = 0; s < 4; s++)
    if (h->cfg.BC_u[s] >= 3 || h->cfg.BC_v[s] >= 3) any_copy = true;
  h->need_allgather_prev = any_copy;
  std::vector<int> slot(nTri, -1), copy_ti;
  std::vector<double> copy_w;
  int nslots = 0;
  if (any_copy) {
    if (nV < 5) { ufe_set_error("copy boundary conditions need a mesh with at least 5 vertices"); return UFE_ERR_INVALID; }
    const double *V = h->hV.data(), *GC = h->hGC.data();
    for (int ti = 0; ti < nTri; ti++) {
      const int bi = h->hTriBI[ti];
      if (bi == 0) continue;
      const int side = (bi <= 2) ? 0 : (bi <= 4) ? 1 : (bi <= 6) ? 2 : 3;
      // the u and v rows may use different copy kinds; tables are keyed on the u choice
      // first, and both choices must agree when both are copy BCs.
      const int cu = h->cfg.BC_u[side], cv = h->cfg.BC_v[side];
      const int kind = cu >= 3 ? cu : (cv >= 3 ? cv : 0);
      if (kind == 0) continue;
      if (cu >= 3 && cv >= 3 && cu != cv) { ufe_set_error("BC_u and BC_v use different copy boundary conditions on one border"); return UFE_ERR_INVALID; }
      const double gx = GC[ti], gy = GC[nTri + ti];
      double px, py;
      if (kind == UFE_BC_PERIODIC_ISMIP_HOM) {
        const double L = h->cfg.refgeo_idealised_ISMIP_HOM_L;
        px = gx > 0.0 ? gx - L / 2.0 : gx + L / 2.0;
        py = gy > 0.0 ? gy - L / 2.0 : gy + L / 2.0;
      } else {
        py = gy;
        px = gx < 0.0 ? h->dm.xmin + (h->dm.xmax - h->dm.xmin) * 1.0 / 3.0
                      : h->dm.xmin + (h->dm.xmax - h->dm.xmin) * 2.0 / 3.0;
      }
      int vi = 5, vi_prev = 5;     // 1-based; the reference always starts the walk at vertex 5
      for (;;) {
        const double d = h_norm2(V[vi - 1] - px, V[nV + vi - 1] - py);
        double dcmin = d + 10.0; int vcmin = 0;
        for (int ci = 0; ci < h->hnC[vi - 1]; ci++) {
          const int vc = h->hC[(size_t)ci * nV + vi - 1];
          if (vc == vi_prev) continue;
          const double dc = h_norm2(V[vc - 1] - px, V[nV + vc - 1] - py);
          if (dc < dcmin) { dcmin = dc; vcmin = vc; }
        }
        if (dcmin < d) { vi_prev = vi; vi = vcmin; } else break;
      }
      slot[ti] = nslots++;
      copy_ti.resize((size_t)nslots * ncm, 0);
      copy_w.resize((size_t)nslots * ncm, 0.0);
      int *ct = &copy_ti[(size_t)(nslots - 1) * ncm];
      double *cw = &copy_w[(size_t)(nslots - 1) * ncm];
      const int nt = h->hniTri[vi - 1];
      double sum = 0.0;
      for (int iti = 0; iti < nt; iti++) {
        const int tj = h->hiTri[(size_t)iti * nV + vi - 1];
        const double dist = h_norm2(px - GC[tj - 1], py - GC[nTri + tj - 1]);
        ct[iti] = tj; cw[iti] = 1.0 / (dist * dist);
      }
      for (int i = 0; i < nt; i++) sum += cw[i];
      for (int i = 0; i < nt; i++) cw[i] = cw[i] / sum;
    }
  }
  UFE_TRY(dupload(&h->bc_slot, slot.data(), slot.size()));
  UFE_TRY(dupload(&h->bc_copy_ti, copy_ti.data(), copy_ti.size()));
  UFE_TRY(dupload(&h->bc_copy_w, copy_w.data(), copy_w.size()));
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------
static int upload_family(const ufe_csr *const *src, int nval, DevFamily &F, int row1, int m_loc) {
  const ufe_csr *A0 = src[0];
  if (A0->i1 != row1 || A0->m_loc != m_loc) { ufe_set_error("operator row range [%d,+%d) does not match the ownership range [%d,+%d)", A0->i1, A0->m_loc, row1, m_loc); return UFE_ERR_INVALID; }
  F.m_loc = m_loc; F.m = A0->m; F.n = A0->n; F.i1 = row1; F.nnz = A0->nnz; F.nval = nval;
  UFE_TRY(dupload(&F.ptr, A0->ptr, (size_t)m_loc + 1));
  UFE_TRY(dupload(&F.ind, A0->ind, (size_t)A0->nnz));
  for (int q = 0; q < nval; q++) {
    if (src[q]->nnz != A0->nnz) { ufe_set_error("operator family members have different nnz"); return UFE_ERR_INVALID; }
    UFE_TRY(dupload(&F.val[q], src[q]->val, (size_t)A0->nnz));
  }
  return UFE_OK;
}

static int make_plan(ufe_handle *h, HaloPlan &plan, int ntot, int own_lo, int own_hi, int need_lo, int need_hi) {
  const int P = h->comm.nranks;
  plan.nranks = P; plan.rank = h->comm.rank;
  plan.own_lo.assign(P, 0); plan.own_hi.assign(P, 0); plan.need_lo.assign(P, 0); plan.need_hi.assign(P, 0);
  if (need_lo > own_lo) need_lo = own_lo;
  if (need_hi < own_hi) need_hi = own_hi;
  if (P == 1) { plan.own_lo[0] = own_lo; plan.own_hi[0] = own_hi; plan.need_lo[0] = need_lo; plan.need_hi[0] = need_hi; return UFE_OK; }
  int mine[4] = {own_lo, own_hi, need_lo, need_hi};
  int *d = nullptr;
  UFE_CUDA(cudaMalloc(&d, sizeof(int) * 4 * (P + 1)));
  UFE_CUDA(cudaMemcpy(d + 4 * P, mine, sizeof mine, cudaMemcpyHostToDevice));
  UFE_NCCL(ncclAllGather(d + 4 * P, d, 4, ncclInt32, h->comm.nccl, h->st));
  UFE_CUDA(cudaStreamSynchronize(h->st));
  std::vector<int> all(4 * P);
  UFE_CUDA(cudaMemcpy(all.data(), d, sizeof(int) * 4 * P, cudaMemcpyDeviceToHost));
  cudaFree(d);
  for (int q = 0; q < P; q++) { plan.own_lo[q] = all[4 * q]; plan.own_hi[q] = all[4 * q + 1]; plan.need_lo[q] = all[4 * q + 2]; plan.need_hi[q] = all[4 * q + 3]; }
  (void)ntot;
  return UFE_OK;
}

static int alloc_fields(ufe_handle *h) {
  const size_t nV = h->dm.nV, nT = h->dm.nTri, nz = h->dm.nz;
  DivaFields &F = h->F;
  struct { double **p; size_t n; } list[] = {
      {&F.u_vav_b, nT}, {&F.v_vav_b, nT}, {&F.u_base_b, nT}, {&F.v_base_b, nT}, {&F.tau_bx_b, nT}, {&F.tau_by_b, nT},
      {&F.eta_3D_b, nT * nz}, {&F.u_3D_b, nT * nz}, {&F.v_3D_b, nT * nz},
      {&F.du_dx_a, nV}, {&F.du_dy_a, nV}, {&F.dv_dx_a, nV}, {&F.dv_dy_a, nV},
      {&F.du_dz_3D_a, nV * nz}, {&F.dv_dz_3D_a, nV * nz}, {&F.eta_3D_a, nV * nz}, {&F.N_a, nV},
      {&F.beta_a, nV}, {&F.beta_eff_a, nV},
      {&F.N_b, nT}, {&F.dN_dx_b, nT}, {&F.dN_dy_b, nT}, {&F.F1_3D_b, nT * nz}, {&F.F2_3D_b, nT * nz},
      {&F.beta_b, nT}, {&F.beta_eff_b, nT}, {&F.tau_dx_b, nT}, {&F.tau_dy_b, nT}, {&F.u_b_prev, nT}, {&F.v_b_prev, nT},
      {&F.rec_b, nT * (size_t)((4 + 2 * nz + 3) / 4 * 4)}, {&F.rec_a, nV * (size_t)((4 + 3 * nz + 3) / 4 * 4)}};
  F.RB = (int)((4 + 2 * nz + 3) / 4 * 4); F.RA = (int)((4 + 3 * nz + 3) / 4 * 4);
  for (auto &e : list) { UFE_TRY(dalloc(e.p, e.n)); h->owned_ptrs.push_back(*e.p); }
  UFE_TRY(dalloc(&h->Hi, nV)); UFE_TRY(dalloc(&h->Hs, nV)); UFE_TRY(dalloc(&h->Hib, nV)); UFE_TRY(dalloc(&h->SL, nV));
  UFE_TRY(dalloc(&h->fraction_gr, nV)); UFE_TRY(dalloc(&h->fraction_gr_b, nT)); UFE_TRY(dalloc(&h->Neff, nV));
  UFE_TRY(dalloc(&h->Ti, nV * nz)); UFE_TRY(dalloc(&h->phi, nV)); UFE_TRY(dalloc(&h->alpha_sq, nV));
  UFE_TRY(dalloc(&h->beta_sq, nV)); UFE_TRY(dalloc(&h->tys, nV)); UFE_TRY(dalloc(&h->bc_u, nT)); UFE_TRY(dalloc(&h->bc_v, nT));
  UFE_TRY(dalloc(&h->mask_gr, nV)); UFE_TRY(dalloc(&h->mask_fl, nV)); UFE_TRY(dalloc(&h->mask_land, nV));
  UFE_TRY(dalloc(&h->bc_mask, nT));
  return UFE_OK;
}

static void peer_teardown(ufe_handle *h) {
  if (!h->sym) return;
  if (h->comm.nccl && h->st) {            // nobody may still be reading this rank's buffer
    int *d = nullptr;
    if (cudaMalloc(&d, sizeof(int)) == cudaSuccess) {
      cudaMemsetAsync(d, 0, sizeof(int), h->st);
      ncclAllReduce(d, d, 1, ncclInt32, ncclSum, h->comm.nccl, h->st);
      cudaStreamSynchronize(h->st);
      cudaFree(d);
    }
  }
  PeerComm &pc = h->comm.peer;
  if (pc.on) for (int q = 0; q < pc.P; q++) if (q != pc.me && pc.base[q]) cudaIpcCloseMemHandle(pc.base[q]);
  pc.on = 0;
  cudaFree(h->sym);
  h->sym = nullptr; h->S.x = nullptr;
}

extern "C" int ufe_diva_destroy(ufe_handle *h) {
  if (!h) return UFE_OK;
  cudaSetDevice(h->device);
  if (h->st) cudaStreamSynchronize(h->st);
  peer_teardown(h);
  for (double *p : h->owned_ptrs) cudaFree(p);
  double *dl[] = {h->Hi, h->Hs, h->Hib, h->SL, h->fraction_gr, h->fraction_gr_b, h->Neff, h->Ti, h->phi, h->alpha_sq,
                  h->beta_sq, h->tys, h->bc_u, h->bc_v, h->bc_copy_w, h->red_partials, h->red_out, h->dm.V, h->dm.TriGC,
                  h->dm.zeta, h->S.val, h->S.valS, h->S.bb, h->S.bS, h->S.x, h->S.bell_val};
  for (double *p : dl) cudaFree(p);
  int *il[] = {h->mask_gr, h->mask_fl, h->mask_land, h->bc_mask, h->bc_slot, h->bc_copy_ti, h->rowkind, h->dm.Tri,
               h->dm.TriC, h->dm.C, h->dm.nC, h->dm.iTri, h->dm.niTri, h->dm.VBI, h->dm.TriBI, h->S.ptr, h->S.ind,
               h->S.bell_off, h->S.bell_col};
  for (int *p : il) cudaFree(p);
  for (int f = 0; f < 3; f++) {
    cudaFree(h->fam[f].ptr); cudaFree(h->fam[f].ind);
    for (int q = 0; q < 5; q++) cudaFree(h->fam[f].val[q]);
  }
  cudaFree(h->red_counter); cudaFree(h->flush_buf); cudaFree(h->op_x); cudaFree(h->op_y);
  if (h->red_host) cudaFreeHost(h->red_host);
  if (h->sec_alloc) { double **sp = reinterpret_cast<double **>(&h->sec); for (size_t i = 0; i < sizeof(SecondaryFields) / sizeof(double *); i++) cudaFree(sp[i]); }
  ufe_krylov_free(h->kw);
  ufe_pclu_free(h->pclu);
  ufe_thickness_free(h->thk);
  for (int i = 0; i < 8; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->comm.nccl) ncclCommDestroy(h->comm.nccl);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return UFE_OK;
}

extern "C" int ufe_diva_set_config(ufe_handle *h, const ufe_config *cfg) {
  UFE_TRY(validate_config(cfg));
  bool bc_changed = false;
  for (int s = 0; s < 4; s++) if (cfg->BC_u[s] != h->cfg.BC_u[s] || cfg->BC_v[s] != h->cfg.BC_v[s]) bc_changed = true;
  if (cfg->refgeo_idealised_ISMIP_HOM_L != h->cfg.refgeo_idealised_ISMIP_HOM_L) bc_changed = true;
  h->cfg = *cfg;
  h->pattern_valid = false;
  if (bc_changed) {
    UFE_CUDA(cudaSetDevice(h->device));
    cudaFree(h->bc_slot); cudaFree(h->bc_copy_ti); cudaFree(h->bc_copy_w);
    h->bc_slot = h->bc_copy_ti = nullptr; h->bc_copy_w = nullptr;
    UFE_TRY(build_bc_tables(h));
  }
  return UFE_OK;
}

// Symmetric buffer + CUDA-IPC mapping of the other ranks' buffers (PeerComm, ufe_internal.cuh).
// If any rank cannot map a peer (no P2P / IPC in this environment) every rank falls back to the
// NCCL halo exchange and all-reduces; UFE_COMM=nccl forces that path (A/B measurements).
static int peer_setup(ufe_handle *h) {
  PeerComm &pc = h->comm.peer;
  const int P = h->comm.nranks, me = h->comm.rank;
  const size_t N = (size_t)2 * h->dm.nTri;
  if (P > UFE_MAX_RANKS) { ufe_set_error("at most %d ranks are supported", UFE_MAX_RANKS); return UFE_ERR_INVALID; }
  pc.P = P; pc.me = me;
  pc.off_pg = 0; pc.off_sg = (long long)N; pc.off_x = 2 * (long long)N; pc.off_dots = 3 * (long long)N;
  pc.off_flags = pc.off_dots + 2LL * P * UFE_PEER_DOTS;
  const size_t total = (size_t)pc.off_flags + 64;
  UFE_CUDA(cudaMalloc(&h->sym, total * sizeof(double)));
  UFE_CUDA(cudaMemset(h->sym, 0, total * sizeof(double)));
  UFE_CUDA(cudaStreamSynchronize(0));
  h->S.x = h->sym + pc.off_x;
  for (int q = 0; q <= P; q++) {
    int i1, i2;
    if (q < P) { ufe_partition_list(h->dm.nTri, q, P, &i1, &i2); pc.bounds[q] = i1 - 1; }
    else pc.bounds[q] = h->dm.nTri;
  }
  for (int q = 1; q < P; q++) if (pc.bounds[q] < pc.bounds[q - 1]) pc.bounds[q] = pc.bounds[q - 1];   // ranks without rows
  const char *env = getenv("UFE_COMM");
  int ok = (env && strcmp(env, "nccl") == 0) ? 0 : 1;
  cudaIpcMemHandle_t mine, *all = nullptr;
  std::vector<cudaIpcMemHandle_t> hall(P);
  if (ok && cudaIpcGetMemHandle(&mine, h->sym) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  char *d = nullptr;
  UFE_CUDA(cudaMalloc(&d, sizeof(mine) * (P + 1)));
  UFE_CUDA(cudaMemcpy(d + sizeof(mine) * P, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  UFE_NCCL(ncclAllGather(d + sizeof(mine) * P, d, sizeof(mine), ncclChar, h->comm.nccl, h->st));
  UFE_CUDA(cudaStreamSynchronize(h->st));
  UFE_CUDA(cudaMemcpy(hall.data(), d, sizeof(mine) * P, cudaMemcpyDeviceToHost));
  (void)all;
  for (int q = 0; q < P; q++) pc.base[q] = nullptr;
  pc.base[me] = h->sym;
  for (int q = 0; q < P && ok; q++) {
    if (q == me) continue;
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, hall[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    pc.base[q] = static_cast<double *>(ptr);
  }
  // all ranks must agree
  int *dok = reinterpret_cast<int *>(d);
  UFE_CUDA(cudaMemcpy(dok, &ok, sizeof(int), cudaMemcpyHostToDevice));
  UFE_NCCL(ncclAllReduce(dok, dok, 1, ncclInt32, ncclMin, h->comm.nccl, h->st));
  UFE_CUDA(cudaStreamSynchronize(h->st));
  int all_ok = 0;
  UFE_CUDA(cudaMemcpy(&all_ok, dok, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d);
  if (!all_ok) {
    for (int q = 0; q < P; q++) if (q != me && pc.base[q]) { cudaIpcCloseMemHandle(pc.base[q]); pc.base[q] = nullptr; }
    pc.on = 0;
  } else pc.on = 1;
  return UFE_OK;
}

extern "C" int ufe_diva_create(const ufe_mesh *mesh, const ufe_config *cfg, const ufe_comm *comm, ufe_handle **out) {
  if (!mesh || !cfg || !out) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  *out = nullptr;
  UFE_TRY(validate_config(cfg));
  if (mesh->nz < 2 || mesh->nz > UFE_NZ_MAX) { ufe_set_error("nz = %d outside [2,%d]", mesh->nz, UFE_NZ_MAX); return UFE_ERR_INVALID; }
  ufe_handle *h = new ufe_handle();
  memset(&h->F, 0, sizeof h->F);
  for (int i = 0; i < 8; i++) h->ev[i] = nullptr;
  h->cfg = *cfg;
  h->comm.rank = comm ? comm->rank : 0;
  h->comm.nranks = comm ? comm->nranks : 1;
  // Small systems are not partitioned: below ~1.3e5 unknowns every kernel of the path is launch-latency-bound
  // (profiles/r1_sweep_spmv_krylov_*.jsonl), so strips only add halo epochs and reductions.  Every rank then solves
  // the whole system redundantly and bit-identically, with no communication at all (PETSc's PCREDUNDANT idea applied
  // to the whole solve).  krylov_pc_strip_only = 1 or UFE_REDUNDANT_MAX_UNKNOWNS=0 keep the row partition.
  {
    long long max_unknowns = 131072;
    if (const char *e = getenv("UFE_REDUNDANT_MAX_UNKNOWNS")) max_unknowns = atoll(e);
    if (h->comm.nranks > 1 && !cfg->krylov_pc_strip_only && 2LL * mesh->nTri <= max_unknowns) {
      h->redundant_ranks = h->comm.nranks;
      h->comm.rank = 0; h->comm.nranks = 1;
    }
  }
  h->device = comm ? comm->device : 0;
  int rc = UFE_OK;
  auto fail = [&](int code) { ufe_diva_destroy(h); return code; };
  if (cudaSetDevice(h->device) != cudaSuccess) { ufe_set_error("cudaSetDevice(%d) failed: no usable CUDA device", h->device); return fail(UFE_ERR_CUDA); }
  if ((rc = check_device()) != UFE_OK) return fail(rc);
  if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { ufe_set_error("stream creation failed"); return fail(UFE_ERR_CUDA); }
  for (int i = 0; i < 8; i++) cudaEventCreate(&h->ev[i]);
  if (h->comm.nranks > 1) {
    if (!comm->nccl_unique_id) { ufe_set_error("nranks > 1 needs an NCCL unique id"); return fail(UFE_ERR_INVALID); }
    ncclUniqueId id;
    memcpy(&id, comm->nccl_unique_id, 128);
    if (ncclCommInitRank(&h->comm.nccl, h->comm.nranks, id, h->comm.rank) != ncclSuccess) { ufe_set_error("ncclCommInitRank failed"); return fail(UFE_ERR_CUDA); }
  }
  // mesh
  DevMesh &dm = h->dm;
  dm.nV = mesh->nV; dm.nTri = mesh->nTri; dm.nC_mem = mesh->nC_mem; dm.nz = mesh->nz;
  dm.xmin = mesh->xmin; dm.xmax = mesh->xmax; dm.ymin = mesh->ymin; dm.ymax = mesh->ymax;
  const size_t nV = dm.nV, nT = dm.nTri, ncm = dm.nC_mem;
#define UP(dst, src, n) if ((rc = dupload(&dst, src, n)) != UFE_OK) return fail(rc)
  UP(dm.V, mesh->V, nV * 2); UP(dm.TriGC, mesh->TriGC, nT * 2); UP(dm.zeta, mesh->zeta, (size_t)dm.nz);
  UP(dm.Tri, mesh->Tri, nT * 3); UP(dm.TriC, mesh->TriC, nT * 3); UP(dm.C, mesh->C, nV * ncm); UP(dm.nC, mesh->nC, nV);
  UP(dm.iTri, mesh->iTri, nV * ncm); UP(dm.niTri, mesh->niTri, nV); UP(dm.VBI, mesh->VBI, nV); UP(dm.TriBI, mesh->TriBI, nT);
#undef UP
  h->hV.assign(mesh->V, mesh->V + nV * 2); h->hGC.assign(mesh->TriGC, mesh->TriGC + nT * 2);
  h->hC.assign(mesh->C, mesh->C + nV * ncm); h->hnC.assign(mesh->nC, mesh->nC + nV);
  h->hiTri.assign(mesh->iTri, mesh->iTri + nV * ncm); h->hniTri.assign(mesh->niTri, mesh->niTri + nV);
  h->hTriBI.assign(mesh->TriBI, mesh->TriBI + nT);
  h->hzeta.assign(mesh->zeta, mesh->zeta + dm.nz);
  // ownership ranges: determine_ownership_ranges_equal -> partition_list (mesh_parallelisation.f90:167-203)
  ufe_partition_list(dm.nV, h->comm.rank, h->comm.nranks, &h->vi1, &h->vi2);
  ufe_partition_list(dm.nTri, h->comm.rank, h->comm.nranks, &h->ti1, &h->ti2);
  const int nv_loc = h->vi2 - h->vi1 + 1, nt_loc = h->ti2 - h->ti1 + 1;
  // operators: received or built on the device
  bool need[3] = {mesh->M_a_b[0] == nullptr, mesh->M_b_a[0] == nullptr, mesh->M2_b_b[0] == nullptr};
  if (!need[0] && (rc = upload_family(mesh->M_a_b, 3, h->fam[0], h->ti1, nt_loc)) != UFE_OK) return fail(rc);
  if (!need[1] && (rc = upload_family(mesh->M_b_a, 3, h->fam[1], h->vi1, nv_loc)) != UFE_OK) return fail(rc);
  if (!need[2] && (rc = upload_family(mesh->M2_b_b, 5, h->fam[2], h->ti1, nt_loc)) != UFE_OK) return fail(rc);
  if ((rc = ufe_build_operators(h->st, dm, h->vi1, h->vi2, h->ti1, h->ti2, need, h->fam)) != UFE_OK) return fail(rc);
  // halo plans from the column ranges of the owned rows (calc_j_node_range, CSR_matrix_basics.f90:529-572)
  {
    int lo, hi;
    if ((rc = ufe_colrange(h->st, h->fam[1].nnz, h->fam[1].ind, &lo, &hi)) != UFE_OK) return fail(rc);
    if (h->fam[1].nnz == 0) { lo = h->ti1; hi = h->ti1 - 1; }
    if ((rc = make_plan(h, h->plan_b_for_a, dm.nTri, h->ti1 - 1, h->ti2, lo - 1, hi)) != UFE_OK) return fail(rc);
    if ((rc = ufe_colrange(h->st, h->fam[0].nnz, h->fam[0].ind, &lo, &hi)) != UFE_OK) return fail(rc);
    if (h->fam[0].nnz == 0) { lo = h->vi1; hi = h->vi1 - 1; }
    if ((rc = make_plan(h, h->plan_a_for_b, dm.nV, h->vi1 - 1, h->vi2, lo - 1, hi)) != UFE_OK) return fail(rc);
    if ((rc = ufe_colrange(h->st, h->fam[2].nnz, h->fam[2].ind, &lo, &hi)) != UFE_OK) return fail(rc);
    if (h->fam[2].nnz == 0) { lo = h->ti1; hi = h->ti1 - 1; }
    if ((rc = make_plan(h, h->plan_b_for_b, dm.nTri, h->ti1 - 1, h->ti2, lo - 1, hi)) != UFE_OK) return fail(rc);
  }
  if ((rc = alloc_fields(h)) != UFE_OK) return fail(rc);
  if ((rc = build_bc_tables(h)) != UFE_OK) return fail(rc);
  if (h->comm.nranks > 1) {
    if ((rc = peer_setup(h)) != UFE_OK) return fail(rc);
    if ((rc = ufe_krylov_alloc(h->kw, 2 * dm.nTri, 2 * nt_loc, true, h->sym + h->comm.peer.off_pg, h->sym + h->comm.peer.off_sg)) != UFE_OK) return fail(rc);
    if (h->comm.peer.on) {      // device-side description of the peers for the fused reductions / signals
      const PeerComm &pc = h->comm.peer;
      PeerDev pd;
      memset(&pd, 0, sizeof pd);
      pd.P = pc.P; pd.me = pc.me;
      for (int q = 0; q < pc.P; q++) { pd.flags[q] = reinterpret_cast<int *>(pc.base[q] + pc.off_flags); pd.dots[q] = pc.base[q] + pc.off_dots; }
      if (cudaMalloc(&h->kw.peer_dev, sizeof pd) != cudaSuccess ||
          cudaMemcpy(h->kw.peer_dev, &pd, sizeof pd, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(reinterpret_cast<char *>(h->kw.sc) + offsetof(KrylovScalars, peer), &h->kw.peer_dev, sizeof(PeerDev *),
                     cudaMemcpyHostToDevice) != cudaSuccess) { ufe_set_error("peer description upload failed"); return fail(UFE_ERR_CUDA); }
    }
  } else {
    if ((rc = dalloc(&h->S.x, (size_t)2 * nT)) != UFE_OK) return fail(rc);
    if ((rc = ufe_krylov_alloc(h->kw, 2 * dm.nTri, 2 * nt_loc, true)) != UFE_OK) return fail(rc);
  }
  if ((rc = dalloc(&h->red_partials, (size_t)UFE_RED_BLOCKS * 2)) != UFE_OK) return fail(rc);
  if ((rc = dalloc(&h->red_out, 4)) != UFE_OK) return fail(rc);
  if ((rc = dalloc(&h->red_counter, 1)) != UFE_OK) return fail(rc);
  if (cudaMallocHost(&h->red_host, sizeof(double) * 4) != cudaSuccess) { ufe_set_error("pinned allocation failed"); return fail(UFE_ERR_CUDA); }
  *out = h;
  return UFE_OK;
}

extern "C" int ufe_get_ownership(ufe_handle *h, int32_t *vi1, int32_t *vi2, int32_t *ti1, int32_t *ti2) {
  *vi1 = h->vi1; *vi2 = h->vi2; *ti1 = h->ti1; *ti2 = h->ti2;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// uploads / downloads
// ------------------------------------------------------------------------------------
#define H2D(dst, src, n) do { if (src) UFE_CUDA(cudaMemcpyAsync(dst, src, sizeof(*(dst)) * (size_t)(n), cudaMemcpyHostToDevice, h->st)); } while (0)
#define D2H(dst, src, n) do { if (dst) UFE_CUDA(cudaMemcpyAsync(dst, src, sizeof(*(src)) * (size_t)(n), cudaMemcpyDeviceToHost, h->st)); } while (0)

static int upload_inputs(ufe_handle *h, const ufe_ice_inputs *ice, int64_t *bytes) {
  const size_t nV = h->dm.nV, nT = h->dm.nTri, nz = h->dm.nz;
  if (!ice->Hi || !ice->Hs || !ice->mask_grounded_ice) { ufe_set_error("ice inputs: Hi, Hs and mask_grounded_ice are required"); return UFE_ERR_INVALID; }
  H2D(h->Hi, ice->Hi, nV); H2D(h->Hs, ice->Hs, nV); H2D(h->Hib, ice->Hib, nV); H2D(h->SL, ice->SL, nV);
  H2D(h->fraction_gr, ice->fraction_gr, nV); H2D(h->fraction_gr_b, ice->fraction_gr_b, nT);
  H2D(h->Neff, ice->effective_pressure, nV);
  H2D(h->mask_gr, ice->mask_grounded_ice, nV); H2D(h->mask_fl, ice->mask_floating_ice, nV);
  H2D(h->mask_land, ice->mask_icefree_land, nV);
  int64_t b = 8 * (6 * nV + nT) + 4 * 3 * nV;
  if (h->cfg.choice_ice_rheology_Glen == UFE_RHEO_HUYBRECHTS1992) {
    if (!ice->Ti) { ufe_set_error("Huybrechts1992 rheology needs Ti"); return UFE_ERR_INVALID; }
    H2D(h->Ti, ice->Ti, nV * nz); b += 8 * nV * nz;
  }
  H2D(h->phi, ice->till_friction_angle, nV); H2D(h->alpha_sq, ice->alpha_sq, nV); H2D(h->beta_sq, ice->beta_sq, nV);
  b += 8 * 3 * nV;
  const bool have = ice->BC_prescr_mask_b || ice->BC_prescr_u_b || ice->BC_prescr_v_b;
  if (have && !(ice->BC_prescr_mask_b && ice->BC_prescr_u_b && ice->BC_prescr_v_b)) {
    ufe_set_error("need to provide prescribed u,v fields and mask!");   // DIVA_main.f90:139-141
    return UFE_ERR_INVALID;
  }
  if (have) {
    H2D(h->bc_mask, ice->BC_prescr_mask_b, nT); H2D(h->bc_u, ice->BC_prescr_u_b, nT); H2D(h->bc_v, ice->BC_prescr_v_b, nT);
    b += 20 * nT;
  } else if (h->have_bc_prescr) {
    UFE_CUDA(cudaMemsetAsync(h->bc_mask, 0, sizeof(int) * nT, h->st));
  }
  if (have || h->have_bc_prescr) h->pattern_valid = false;
  h->have_bc_prescr = have;
  // grounded_ice_exists = any(mask_grounded_ice), DIVA_main.f90:123-124 (global arrays on every rank)
  int any = 0;
  for (size_t i = 0; i < nV; i++) if (ice->mask_grounded_ice[i]) { any = 1; break; }
  h->grounded_ice_exists = any;
  if (bytes) *bytes += b;
  return UFE_OK;
}

static int upload_state(ufe_handle *h, const ufe_diva_state *s, int64_t *bytes) {
  const size_t nT = h->dm.nTri, nz = h->dm.nz;
  H2D(h->F.u_vav_b, s->u_vav_b, nT); H2D(h->F.v_vav_b, s->v_vav_b, nT);
  H2D(h->F.tau_bx_b, s->tau_bx_b, nT); H2D(h->F.tau_by_b, s->tau_by_b, nT);
  H2D(h->F.eta_3D_b, s->eta_3D_b, nT * nz);
  H2D(h->F.u_base_b, s->u_base_b, nT); H2D(h->F.v_base_b, s->v_base_b, nT);
  if (bytes) *bytes += 8 * (6 * nT + nT * nz);
  return UFE_OK;
}

static int download_state(ufe_handle *h, ufe_diva_state *s, int64_t *bytes) {
  const size_t nV = h->dm.nV, nT = h->dm.nTri, nz = h->dm.nz;
  const DivaFields &F = h->F;
  D2H(s->u_vav_b, F.u_vav_b, nT); D2H(s->v_vav_b, F.v_vav_b, nT); D2H(s->tau_bx_b, F.tau_bx_b, nT);
  D2H(s->tau_by_b, F.tau_by_b, nT); D2H(s->eta_3D_b, F.eta_3D_b, nT * nz);
  D2H(s->u_base_b, F.u_base_b, nT); D2H(s->v_base_b, F.v_base_b, nT);
  D2H(s->u_3D_b, F.u_3D_b, nT * nz); D2H(s->v_3D_b, F.v_3D_b, nT * nz);
  D2H(s->du_dx_a, F.du_dx_a, nV); D2H(s->du_dy_a, F.du_dy_a, nV); D2H(s->dv_dx_a, F.dv_dx_a, nV); D2H(s->dv_dy_a, F.dv_dy_a, nV);
  D2H(s->du_dz_3D_a, F.du_dz_3D_a, nV * nz); D2H(s->dv_dz_3D_a, F.dv_dz_3D_a, nV * nz);
  D2H(s->eta_3D_a, F.eta_3D_a, nV * nz);
  D2H(s->basal_friction_coefficient_a, F.beta_a, nV);
  if (bytes) {
    int64_t b = 8 * (6 * nT + nT * nz);
    if (s->u_3D_b) b += 8 * nT * nz; if (s->v_3D_b) b += 8 * nT * nz;
    if (s->du_dx_a) b += 8 * 4 * nV; if (s->du_dz_3D_a) b += 8 * 2 * nV * nz;
    if (s->eta_3D_a) b += 8 * nV * nz; if (s->basal_friction_coefficient_a) b += 8 * nV;
    *bytes += b;
  }
  UFE_CUDA(cudaStreamSynchronize(h->st));
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// the linearised solve on resident data (assembly + Krylov), solve_linearised_SSA_DIVA.f90:23-178
// ------------------------------------------------------------------------------------
static VertexInputs vertex_inputs(const ufe_handle *h) {
  return VertexInputs{h->dm.V, h->Hi, h->Hib, h->SL, h->fraction_gr, h->Neff, h->Ti, h->tys, h->alpha_sq, h->beta_sq,
                      h->mask_gr, h->mask_fl};
}

static int ensure_pattern(ufe_handle *h) {
  if (h->pattern_valid) return UFE_OK;
  ufe_pclu_free(h->pclu); h->pclu = nullptr; h->pc_used = -1; h->pc_age = -1;
  const int nt = h->ti2 - h->ti1 + 1;
  double *x_keep = h->S.x;
  UFE_TRY(ufe_build_stiffness_pattern(h->st, h->ti1 - 1, nt, h->dm.nTri, make_asm_params(h), view_of(h->fam[2]),
                                      h->dm.TriBI, h->dm.TriC, h->have_bc_prescr ? h->bc_mask : nullptr, h->S,
                                      &h->rowkind));
  h->S.x = x_keep;
  h->pattern_valid = true;
  return UFE_OK;
}

static int gather_prev(ufe_handle *h) {
  // gather_to_all(u_b, u_b_prev), (v_b, v_b_prev) -- solve_linearised_SSA_DIVA.f90:54-55.
  // Free rows only read the local ti; the full gather is only needed by the copy BCs.
  const size_t nT = h->dm.nTri;
  if (h->comm.nranks > 1 && h->need_allgather_prev) {
    HaloPlan all = h->plan_b_for_b;
    for (int q = 0; q < all.nranks; q++) { all.need_lo[q] = 0; all.need_hi[q] = (int)nT; }
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, all, h->F.u_vav_b, 0, 1, 1));
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, all, h->F.v_vav_b, 0, 1, 1));
  }
  UFE_CUDA(cudaMemcpyAsync(h->F.u_b_prev, h->F.u_vav_b, sizeof(double) * nT, cudaMemcpyDeviceToDevice, h->st));
  UFE_CUDA(cudaMemcpyAsync(h->F.v_b_prev, h->F.v_vav_b, sizeof(double) * nT, cudaMemcpyDeviceToDevice, h->st));
  return UFE_OK;
}

static int linearised_resident(ufe_handle *h, double rtol, double abstol, int *n_its, int *flags, float *ms_asm,
                               float *ms_kry) {
  const int nt = h->ti2 - h->ti1 + 1;
  UFE_TRY(ensure_pattern(h));
  UFE_TRY(gather_prev(h));
  BCTables T{h->bc_slot, h->bc_copy_ti, h->bc_copy_w, h->dm.nC_mem};
  cudaEventRecord(h->ev[2], h->st);
  UFE_TRY(ufe_launch_assemble(h->st, h->ti1 - 1, nt, h->dm.nTri, make_asm_params(h), view_of(h->fam[2]), h->dm.TriC,
                              h->rowkind, T, h->bc_mask, h->bc_u, h->bc_v, h->F, h->S, 1));
  cudaEventRecord(h->ev[3], h->st);
  PcLU *pc = nullptr;
  if (h->pc_used < 0) {                               // once per cached pattern
    h->pc_used = h->cfg.krylov_pc;
    const double *gcx = h->hGC.data(), *gcy = h->hGC.data() + h->dm.nTri;
    const bool pow2 = (h->comm.nranks & (h->comm.nranks - 1)) == 0;
    if (h->cfg.krylov_pc == UFE_PC_AUTO) {
      // auto: the multifrontal nested-dissection factorisation (exact, any mesh shape, 1 / 2 / 4 / 8 ranks) when its
      // fronts fit; else the banded exact block solve; else 2x2 block Jacobi.  UFE_AUTO_ND=0 skips the first choice.
      // The choice made is reported in ufe_solve_info.krylov_pc_used.
      const char *e = getenv("UFE_AUTO_ND");
      int rc = UFE_ERR_INVALID;
      if (!(e && atoi(e) == 0) && pow2 && !h->cfg.krylov_pc_strip_only) {
        rc = ufe_pclu_setup_nd(h->st, h->S, &h->comm, h->dm.nTri, gcx, gcy, &h->pclu);
        if (rc == UFE_ERR_CUDA) return rc;
        if (rc == UFE_OK) h->pc_used = UFE_PC_ND_LU; else h->pclu = nullptr;
      }
      if (rc != UFE_OK) {
        if (h->comm.nranks > 1 && !h->cfg.krylov_pc_strip_only) {      // replicated exact solve when the whole system fits
          HaloPlan all = h->plan_b_for_b;
          for (int q = 0; q < all.nranks; q++) { all.need_lo[q] = 0; all.need_hi[q] = h->dm.nTri; }
          rc = ufe_pclu_setup(h->st, h->S, 1, (size_t)24 << 30, &h->pclu, &h->comm, &all);
          if (rc == UFE_ERR_CUDA) return rc;
        }
        if (rc != UFE_OK) rc = ufe_pclu_setup(h->st, h->S, 0, (size_t)24 << 30, &h->pclu);
        if (rc == UFE_ERR_CUDA) return rc;
        h->pc_used = rc == UFE_OK ? UFE_PC_BJACOBI_LU : UFE_PC_BJACOBI2;
      }
    } else if (h->cfg.krylov_pc == UFE_PC_BJACOBI_LU) {
      int rc = UFE_ERR_INVALID;
      if (h->comm.nranks > 1 && !h->cfg.krylov_pc_strip_only) {
        HaloPlan all = h->plan_b_for_b;
        for (int q = 0; q < all.nranks; q++) { all.need_lo[q] = 0; all.need_hi[q] = h->dm.nTri; }
        rc = ufe_pclu_setup(h->st, h->S, 1, (size_t)24 << 30, &h->pclu, &h->comm, &all);
        if (rc == UFE_ERR_CUDA) return rc;
      }
      if (rc != UFE_OK) rc = ufe_pclu_setup(h->st, h->S, 0, (size_t)100 << 30, &h->pclu);
      if (rc != UFE_OK) { h->pc_used = -1; return rc; }
    } else if (h->cfg.krylov_pc == UFE_PC_ND_LU) {    // multifrontal nested dissection, one analysis per pattern
      if (!pow2) { h->pc_used = -1; ufe_set_error("krylov_pc nd_lu: the number of ranks must be 1, 2, 4 or 8"); return UFE_ERR_INVALID; }
      const int rc = ufe_pclu_setup_nd(h->st, h->S, &h->comm, h->dm.nTri, gcx, gcy, &h->pclu);
      if (rc != UFE_OK) { h->pc_used = -1; return rc; }
    }
  }
  if (h->pc_used == UFE_PC_BJACOBI_LU || h->pc_used == UFE_PC_ND_LU) {
    // PCSetUp.  The factorisation of an earlier Picard iteration's matrix is still a good
    // preconditioner while the viscosity changes slowly, so it is reused (krylov_pc_lag > 0) until
    // the Krylov count shows it has aged: refactorise when the previous solve needed more than
    // krylov_pc_lag iterations.  A solve that fails with a reused factorisation is repeated with a
    // fresh one, so the result always satisfies the same stopping rule.
    const int lag = h->cfg.krylov_pc_lag;
    bool fresh = false;
    if (h->pc_age < 0 || lag <= 0 || h->pc_last_its > lag) {
      UFE_TRY(ufe_pclu_factor(h->st, h->S, h->pclu));
      h->pc_age = 0; h->pc_factorisations++; fresh = true;
    } else { h->pc_age++; ufe_pclu_mark_reused(h->pclu); }
    pc = h->pclu;
    int its1 = 0, fl1 = 0;
    const int cap = fresh ? h->cfg.krylov_maxits : std::max(4 * lag, 8);
    UFE_TRY(ufe_krylov_run(h->st, h->S, h->kw, h->comm, h->comm.nranks > 1 ? &h->plan_b_for_b : nullptr,
                           h->cfg.krylov_method, rtol, abstol, cap, h->cfg.krylov_guess_nonzero, &its1, &fl1, pc));
    if (!fresh && fl1 != 0) {
      int its2 = 0;
      UFE_TRY(ufe_pclu_factor(h->st, h->S, h->pclu));
      h->pc_age = 0; h->pc_factorisations++;
      UFE_TRY(ufe_krylov_run(h->st, h->S, h->kw, h->comm, h->comm.nranks > 1 ? &h->plan_b_for_b : nullptr,
                             h->cfg.krylov_method, rtol, abstol, h->cfg.krylov_maxits, h->cfg.krylov_guess_nonzero,
                             &its2, &fl1, pc));
      h->pc_last_its = its2;
      its1 += its2;
    } else h->pc_last_its = its1;
    if (n_its) *n_its = its1;
    if (flags) *flags = fl1;
  } else {
    UFE_TRY(ufe_krylov_run(h->st, h->S, h->kw, h->comm, h->comm.nranks > 1 ? &h->plan_b_for_b : nullptr,
                           h->cfg.krylov_method, rtol, abstol, h->cfg.krylov_maxits, h->cfg.krylov_guess_nonzero,
                           n_its, flags, nullptr));
  }
  cudaEventRecord(h->ev[4], h->st);
  UFE_CUDA(cudaEventSynchronize(h->ev[4]));
  if (ms_asm) cudaEventElapsedTime(ms_asm, h->ev[2], h->ev[3]);
  if (ms_kry) cudaEventElapsedTime(ms_kry, h->ev[3], h->ev[4]);
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// Picard driver (DIVA and SSA share it)
// ------------------------------------------------------------------------------------
static int picard_resident(ufe_handle *h, int is_diva, ufe_solve_info *info) {
  const ufe_config &c = h->cfg;
  const int nV = h->dm.nV, nTri = h->dm.nTri, nz = h->dm.nz;
  const int nv = h->vi2 - h->vi1 + 1, nt = h->ti2 - h->ti1 + 1, v0 = h->vi1 - 1, t0 = h->ti1 - 1;
  DivaFields &F = h->F;
  memset(info, 0, sizeof *info);
  const int64_t launches0 = g_launch_count;
  h->sec_current = false;
  h->outputs_gathered = false;
  h->last_is_diva = is_diva;
  cudaEventRecord(h->ev[0], h->st);
  if (!h->grounded_ice_exists) {           // DIVA_main.f90:123-134
    const size_t nT = nTri;
    UFE_CUDA(cudaMemsetAsync(F.u_vav_b, 0, 8 * nT, h->st)); UFE_CUDA(cudaMemsetAsync(F.v_vav_b, 0, 8 * nT, h->st));
    if (is_diva) {
      UFE_CUDA(cudaMemsetAsync(F.u_base_b, 0, 8 * nT, h->st)); UFE_CUDA(cudaMemsetAsync(F.v_base_b, 0, 8 * nT, h->st));
      UFE_CUDA(cudaMemsetAsync(F.u_3D_b, 0, 8 * nT * nz, h->st)); UFE_CUDA(cudaMemsetAsync(F.v_3D_b, 0, 8 * nT * nz, h->st));
    }
    UFE_CUDA(cudaStreamSynchronize(h->st));
    return UFE_OK;
  }
  ClosureParams P = make_params(h, c.Glens_flow_law_epsilon_sq_0);
  const bool multi = h->comm.nranks > 1;
  // velocity-independent pieces, once per solve
  if (c.choice_sliding_law == UFE_SLID_COULOMB || c.choice_sliding_law == UFE_SLID_BUDD ||
      c.choice_sliding_law == UFE_SLID_ZOET_IVERSON)
    UFE_TRY(ufe_launch_till(h->st, nV, P, h->Neff, h->phi, h->mask_land, h->mask_gr, h->dm.C, h->dm.nC, h->tys));
  UFE_TRY(ufe_launch_driving_stress(h->st, t0, nt, view_of(h->fam[0]), h->Hi, h->Hs, F.tau_dx_b, F.tau_dy_b));

  double L2_uv = 1e9, L2_prev;
  int nit_diverg_consec = 0;
  double relax = c.visc_it_relax, eps0 = c.Glens_flow_law_epsilon_sq_0;
  int it = 0, n_Axb = 0, flags = 0;
  bool converged = false;
  double ms_clo = 0, ms_asm = 0, ms_kry = 0;
  VertexInputs VI = vertex_inputs(h);
  while (!converged) {
    it++;
    cudaEventRecord(h->ev[5], h->st);
    if (is_diva) {
      // b-grid gather records of the owned triangles (velocities + du/dz, dv/dz per layer), then ONE halo message per peer
      UFE_TRY(ufe_launch_pack_b(h->st, t0, nt, nTri, nz, P, F));
      if (multi) UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, F.rec_b, 0, 1, F.RB));
    } else if (multi) {
      UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, F.u_vav_b, 0, 1, 1));
      UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, F.v_vav_b, 0, 1, 1));
    }
    UFE_TRY(ufe_launch_vertex(h->st, is_diva, v0, nv, nV, nTri, nz, P, view_of(h->fam[1]), VI, F));
    if (multi) {
      if (is_diva) UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_a_for_b, F.rec_a, 0, 1, F.RA));   // N, beta, beta_eff, (eta, F1, F2) per layer
      else {
        UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_a_for_b, F.N_a, 0, 1, 1));
        UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_a_for_b, F.beta_a, 0, 1, 1));
      }
    }
    UFE_TRY(ufe_launch_triangle(h->st, is_diva, t0, nt, nV, nTri, nz, P, view_of(h->fam[0]), h->fraction_gr_b, F));
    cudaEventRecord(h->ev[6], h->st);
    int its = 0, fl = 0;
    float a = 0, k = 0, cl = 0;
    UFE_TRY(linearised_resident(h, c.stress_balance_PETSc_rtol, c.stress_balance_PETSc_abstol, &its, &fl, &a, &k));
    cudaEventElapsedTime(&cl, h->ev[5], h->ev[6]);
    ms_clo += cl; ms_asm += a; ms_kry += k;
    n_Axb += its; flags |= fl;
    UFE_TRY(ufe_launch_post_picard(h->st, t0, nt, nTri, is_diva, P, relax, h->S.x, F, h->red_partials, h->red_counter,
                                   h->red_out));
    if (multi) { UFE_NCCL(ncclAllReduce(h->red_out, h->red_out, 2, ncclDouble, ncclSum, h->comm.nccl, h->st)); g_launch_count++; }
    UFE_CUDA(cudaMemcpyAsync(h->red_host, h->red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    UFE_CUDA(cudaStreamSynchronize(h->st));
    L2_prev = L2_uv;
    L2_uv = 2.0 * h->red_host[0] / std::max(h->red_host[1], 1e-8);       // SSA_DIVA_utilities.f90:178
    if (L2_uv > L2_prev) nit_diverg_consec++; else nit_diverg_consec = 0;
    if (nit_diverg_consec > 2) {
      nit_diverg_consec = 0;
      relax *= 0.9; eps0 *= 1.2;
      P = make_params(h, eps0);
    }
    if (relax <= 0.05 || eps0 >= 1e-5) {
      if (relax < 0.05) { ufe_set_error("viscosity iteration still diverges even with very low relaxation factor!"); return UFE_ERR_PICARD_DIVERGED; }
      else if (eps0 > 1e-5) { ufe_set_error("viscosity iteration still diverges even with very high effective strain rate regularisation!"); return UFE_ERR_PICARD_DIVERGED; }
    }
    converged = L2_uv < c.visc_it_norm_dUV_tol;
    if (it > c.visc_it_nit) { flags |= UFE_FLAG_PICARD_MAXIT; break; }
  }
  if (is_diva) UFE_TRY(ufe_launch_vel3d(h->st, t0, nt, nTri, nz, P, F));
  cudaEventRecord(h->ev[1], h->st);
  UFE_CUDA(cudaEventSynchronize(h->ev[1]));
  float tot = 0;
  cudaEventElapsedTime(&tot, h->ev[0], h->ev[1]);
  info->n_visc_its = it; info->n_Axb_its = n_Axb; info->flags = flags; info->L2_uv = L2_uv;
  info->visc_it_relax_applied = relax; info->Glens_flow_law_epsilon_sq_0_applied = eps0;
  info->ms_total = tot; info->ms_closures = ms_clo; info->ms_assembly = ms_asm; info->ms_krylov = ms_kry;
  info->gpu_launches = g_launch_count - launches0;
  info->krylov_pc_used = h->pc_used;
  info->reserved = h->redundant_ranks > 0 ? 2 : h->comm.peer.on;        // 1: peer-memory halo reads + reductions inside the Krylov loop
  return UFE_OK;
}

// all ranks end up with full-length results (the reference keeps distributed slices; the
// C ABI exchanges full arrays, so gather the owned slices of the outputs at the end)
static int gather_outputs(ufe_handle *h, int is_diva) {
  if (h->comm.nranks <= 1 || h->outputs_gathered) return UFE_OK;
  const int nTri = h->dm.nTri, nV = h->dm.nV, nz = h->dm.nz;
  HaloPlan pb = h->plan_b_for_b, pa = h->plan_a_for_b;
  for (int q = 0; q < pb.nranks; q++) { pb.need_lo[q] = 0; pb.need_hi[q] = nTri; pa.need_lo[q] = 0; pa.need_hi[q] = nV; }
  DivaFields &F = h->F;
  double *b1[] = {F.u_vav_b, F.v_vav_b, F.tau_bx_b, F.tau_by_b, F.u_base_b, F.v_base_b};
  for (int i = 0; i < (is_diva ? 6 : 2); i++) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pb, b1[i], 0, 1, 1));
  double *a1[] = {F.du_dx_a, F.du_dy_a, F.dv_dx_a, F.dv_dy_a, F.beta_a};
  for (double *p : a1) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pa, p, 0, 1, 1));
  if (is_diva) {
    double *b3[] = {F.eta_3D_b, F.u_3D_b, F.v_3D_b};
    for (double *p : b3) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pb, p, nTri, nz, 1));
    double *a3[] = {F.du_dz_3D_a, F.dv_dz_3D_a, F.eta_3D_a};
    for (double *p : a3) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pa, p, nV, nz, 1));
  }
  h->outputs_gathered = true;
  return UFE_OK;
}
// several ranks: make the resident solution fields full-length on this rank (collective)
int ufe_handle_gather_outputs(ufe_handle *h) { return gather_outputs(h, h->last_is_diva); }

extern "C" int ufe_diva_upload(ufe_handle *h, const ufe_ice_inputs *ice, const ufe_diva_state *state) {
  UFE_CUDA(cudaSetDevice(h->device));
  if (ice) UFE_TRY(upload_inputs(h, ice, nullptr));
  if (state) UFE_TRY(upload_state(h, state, nullptr));
  UFE_CUDA(cudaStreamSynchronize(h->st));
  return UFE_OK;
}
extern "C" int ufe_diva_solve_resident(ufe_handle *h, ufe_solve_info *info) {
  UFE_CUDA(cudaSetDevice(h->device));
  ufe_solve_info tmp;
  return picard_resident(h, 1, info ? info : &tmp);
}
extern "C" int ufe_diva_reset_state(ufe_handle *h) {
  UFE_CUDA(cudaSetDevice(h->device));
  const size_t nT = h->dm.nTri, nz = h->dm.nz;
  DivaFields &F = h->F;
  double *b1[] = {F.u_vav_b, F.v_vav_b, F.tau_bx_b, F.tau_by_b, F.u_base_b, F.v_base_b};
  for (double *p : b1) UFE_CUDA(cudaMemsetAsync(p, 0, 8 * nT, h->st));
  UFE_CUDA(cudaMemsetAsync(F.eta_3D_b, 0, 8 * nT * nz, h->st));
  return UFE_OK;
}
extern "C" int ufe_diva_download(ufe_handle *h, ufe_diva_state *state) {
  UFE_CUDA(cudaSetDevice(h->device));
  UFE_TRY(gather_outputs(h, 1));
  return download_state(h, state, nullptr);
}

extern "C" int ufe_diva_solve(ufe_handle *h, const ufe_ice_inputs *ice, ufe_diva_state *state, ufe_solve_info *info) {
  if (!h || !ice || !state) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  ufe_solve_info tmp;
  if (!info) info = &tmp;
  cudaEvent_t e0 = h->ev[7];
  cudaEventRecord(e0, h->st);
  UFE_TRY(upload_inputs(h, ice, nullptr));
  UFE_TRY(upload_state(h, state, nullptr));
  UFE_CUDA(cudaStreamSynchronize(h->st));
  float up = 0, dn = 0;
  cudaEventRecord(h->ev[6], h->st); cudaEventSynchronize(h->ev[6]);
  cudaEventElapsedTime(&up, e0, h->ev[6]);
  UFE_TRY(picard_resident(h, 1, info));
  cudaEventRecord(e0, h->st);
  UFE_TRY(gather_outputs(h, 1));
  UFE_TRY(download_state(h, state, nullptr));
  cudaEventRecord(h->ev[6], h->st); cudaEventSynchronize(h->ev[6]);
  cudaEventElapsedTime(&dn, e0, h->ev[6]);
  info->ms_h2d = up; info->ms_d2h = dn;
  return UFE_OK;
}

extern "C" int ufe_ssa_solve(ufe_handle *h, const ufe_ice_inputs *ice, ufe_ssa_state *state, ufe_solve_info *info) {
  if (!h || !ice || !state) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  ufe_solve_info tmp;
  if (!info) info = &tmp;
  const size_t nT = h->dm.nTri, nV = h->dm.nV;
  UFE_TRY(upload_inputs(h, ice, nullptr));
  H2D(h->F.u_vav_b, state->u_b, nT); H2D(h->F.v_vav_b, state->v_b, nT);
  UFE_TRY(picard_resident(h, 0, info));
  UFE_TRY(gather_outputs(h, 0));
  D2H(state->u_b, h->F.u_vav_b, nT); D2H(state->v_b, h->F.v_vav_b, nT);
  D2H(state->basal_friction_coefficient_a, h->F.beta_a, nV);
  UFE_CUDA(cudaStreamSynchronize(h->st));
  return UFE_OK;
}

extern "C" int ufe_ssa_diva_linearised(ufe_handle *h, double *u_b, double *v_b, const double *N_b, const double *dN_dx_b,
                                       const double *dN_dy_b, const double *beta_b, const double *tau_dx_b,
                                       const double *tau_dy_b, double *u_b_prev, double *v_b_prev, double rtol,
                                       double abstol, int32_t *n_Axb_its, const int32_t *bc_mask, const double *bc_u,
                                       const double *bc_v) {
  if (!h || !u_b || !v_b || !N_b || !dN_dx_b || !dN_dy_b || !beta_b || !tau_dx_b || !tau_dy_b) {
    ufe_set_error("null argument: every array of solve_SSA_DIVA_linearised is required"); return UFE_ERR_INVALID;
  }
  if (bc_mask && (!bc_u || !bc_v)) { ufe_set_error("BC_prescr_mask_b given without BC_prescr_u_b / BC_prescr_v_b"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  const size_t nT = h->dm.nTri;
  DivaFields &F = h->F;
  H2D(F.u_vav_b, u_b, nT); H2D(F.v_vav_b, v_b, nT); H2D(F.N_b, N_b, nT); H2D(F.dN_dx_b, dN_dx_b, nT);
  H2D(F.dN_dy_b, dN_dy_b, nT); H2D(F.beta_eff_b, beta_b, nT); H2D(F.tau_dx_b, tau_dx_b, nT); H2D(F.tau_dy_b, tau_dy_b, nT);
  const bool have = bc_mask != nullptr;
  if (have) { H2D(h->bc_mask, bc_mask, nT); H2D(h->bc_u, bc_u, nT); H2D(h->bc_v, bc_v, nT); }
  else if (h->have_bc_prescr) UFE_CUDA(cudaMemsetAsync(h->bc_mask, 0, sizeof(int) * nT, h->st));
  if (have || h->have_bc_prescr) h->pattern_valid = false;
  h->have_bc_prescr = have;
  int its = 0, fl = 0;
  UFE_TRY(linearised_resident(h, rtol, abstol, &its, &fl, nullptr, nullptr));
  // de-interleave (:163-173) into full-length u_b, v_b
  std::vector<double> x(2 * nT);
  if (h->comm.nranks > 1) {
    HaloPlan all = h->plan_b_for_b;
    for (int q = 0; q < all.nranks; q++) { all.need_lo[q] = 0; all.need_hi[q] = (int)nT; }
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, all, h->S.x, 0, 1, 2));
  }
  UFE_CUDA(cudaMemcpyAsync(x.data(), h->S.x, sizeof(double) * 2 * nT, cudaMemcpyDeviceToHost, h->st));
  D2H(u_b_prev, F.u_b_prev, nT); D2H(v_b_prev, F.v_b_prev, nT);
  UFE_CUDA(cudaStreamSynchronize(h->st));
  for (size_t t = 0; t < nT; t++) { u_b[t] = x[2 * t]; v_b[t] = x[2 * t + 1]; }
  if (n_Axb_its) *n_Axb_its = its;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// calc_secondary_velocities on the resident result of the last DIVA solve
// ------------------------------------------------------------------------------------
extern "C" int ufe_calc_secondary_velocities(ufe_handle *h, ufe_secondary_velocities *out) {
  if (!h || !out) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  const size_t nV = h->dm.nV, nT = h->dm.nTri, nz = h->dm.nz;
  SecondaryFields &O = h->sec;
  if (!h->sec_alloc) {
    memset(&O, 0, sizeof O);
    double **b9[] = {&O.u_surf_b, &O.v_surf_b, &O.uabs_surf_b, &O.u_base_b, &O.v_base_b, &O.uabs_base_b, &O.u_vav_b, &O.v_vav_b, &O.uabs_vav_b};
    for (double **p : b9) UFE_TRY(dalloc(p, nT));
    UFE_TRY(dalloc(&O.u_3D, nV * nz)); UFE_TRY(dalloc(&O.v_3D, nV * nz));
    double **a10[] = {&O.u_surf, &O.v_surf, &O.uabs_surf, &O.u_base, &O.v_base, &O.uabs_base, &O.u_vav, &O.v_vav, &O.uabs_vav, &O.R_shear};
    for (double **p : a10) UFE_TRY(dalloc(p, nV));
    h->sec_alloc = true;
  }
  const int nv = h->vi2 - h->vi1 + 1, nt = h->ti2 - h->ti1 + 1;
  ClosureParams P = make_params(h, h->cfg.Glens_flow_law_epsilon_sq_0);
  UFE_TRY(ufe_launch_secondary_b(h->st, h->ti1 - 1, nt, (int)nT, (int)nz, P, h->F.u_3D_b, h->F.v_3D_b, O));
  if (h->comm.nranks > 1) {     // the b->a maps read the triangles around the owned vertices
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, h->F.u_3D_b, (long long)nT, (int)nz, 1));
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, h->F.v_3D_b, (long long)nT, (int)nz, 1));
    double *b6[] = {O.u_surf_b, O.v_surf_b, O.u_base_b, O.v_base_b, O.u_vav_b, O.v_vav_b};
    for (double *p : b6) UFE_TRY(ufe_halo_exchange(h->st, h->comm, h->plan_b_for_a, p, 0, 1, 1));
  }
  UFE_TRY(ufe_launch_secondary_a(h->st, h->vi1 - 1, nv, (int)nV, (int)nT, (int)nz, view_of(h->fam[1]), h->F.u_3D_b, h->F.v_3D_b, O));
  if (h->comm.nranks > 1) {     // full-length results on every rank, like the other outputs of the C ABI
    HaloPlan pb = h->plan_b_for_b, pa = h->plan_a_for_b;
    for (int q = 0; q < pb.nranks; q++) { pb.need_lo[q] = 0; pb.need_hi[q] = (int)nT; pa.need_lo[q] = 0; pa.need_hi[q] = (int)nV; }
    double *b9[] = {O.u_surf_b, O.v_surf_b, O.uabs_surf_b, O.u_base_b, O.v_base_b, O.uabs_base_b, O.u_vav_b, O.v_vav_b, O.uabs_vav_b};
    for (double *p : b9) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pb, p, 0, 1, 1));
    double *a10[] = {O.u_surf, O.v_surf, O.uabs_surf, O.u_base, O.v_base, O.uabs_base, O.u_vav, O.v_vav, O.uabs_vav, O.R_shear};
    for (double *p : a10) UFE_TRY(ufe_halo_exchange(h->st, h->comm, pa, p, 0, 1, 1));
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, pa, O.u_3D, (long long)nV, (int)nz, 1));
    UFE_TRY(ufe_halo_exchange(h->st, h->comm, pa, O.v_3D, (long long)nV, (int)nz, 1));
  }
  D2H(out->u_surf_b, O.u_surf_b, nT); D2H(out->v_surf_b, O.v_surf_b, nT); D2H(out->uabs_surf_b, O.uabs_surf_b, nT);
  D2H(out->u_base_b, O.u_base_b, nT); D2H(out->v_base_b, O.v_base_b, nT); D2H(out->uabs_base_b, O.uabs_base_b, nT);
  D2H(out->u_vav_b, O.u_vav_b, nT); D2H(out->v_vav_b, O.v_vav_b, nT); D2H(out->uabs_vav_b, O.uabs_vav_b, nT);
  D2H(out->u_3D, O.u_3D, nV * nz); D2H(out->v_3D, O.v_3D, nV * nz);
  D2H(out->u_surf, O.u_surf, nV); D2H(out->v_surf, O.v_surf, nV); D2H(out->uabs_surf, O.uabs_surf, nV);
  D2H(out->u_base, O.u_base, nV); D2H(out->v_base, O.v_base, nV); D2H(out->uabs_base, O.uabs_base, nV);
  D2H(out->u_vav, O.u_vav, nV); D2H(out->v_vav, O.v_vav, nV); D2H(out->uabs_vav, O.uabs_vav, nV);
  D2H(out->R_shear, O.R_shear, nV);
  UFE_CUDA(cudaStreamSynchronize(h->st));
  h->sec_current = true;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// operator access
// ------------------------------------------------------------------------------------
extern "C" int ufe_mesh_get_operator(ufe_handle *h, int32_t family, int32_t which, int32_t *m_loc, int32_t *nnz,
                                     int32_t *ptr, int32_t *ind, double *val) {
  if (family < 0 || family > 2 || which < 0 || which >= h->fam[family].nval) { ufe_set_error("bad operator id"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  const DevFamily &F = h->fam[family];
  *m_loc = F.m_loc; *nnz = F.nnz;
  if (!ind) return UFE_OK;
  UFE_CUDA(cudaMemcpy(ptr, F.ptr, sizeof(int) * (F.m_loc + 1), cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(ind, F.ind, sizeof(int) * F.nnz, cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(val, F.val[which], sizeof(double) * F.nnz, cudaMemcpyDeviceToHost));
  return UFE_OK;
}

extern "C" int ufe_mesh_apply_operator(ufe_handle *h, int32_t family, int32_t which, const double *x, double *y,
                                       int32_t nlayers) {
  if (family < 0 || family > 2 || which < 0 || which >= h->fam[family].nval) { ufe_set_error("bad operator id"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  const DevFamily &F = h->fam[family];
  if (!x || !y || nlayers < 1) { ufe_set_error("apply_operator: null argument"); return UFE_ERR_INVALID; }
  const size_t nx = (size_t)F.n * nlayers, ny = (size_t)F.m * nlayers;
  if (nx > h->op_x_n) { cudaFree(h->op_x); h->op_x = nullptr; h->op_x_n = 0; UFE_TRY(dalloc(&h->op_x, nx)); h->op_x_n = nx; }
  if (ny > h->op_y_n) { cudaFree(h->op_y); h->op_y = nullptr; h->op_y_n = 0; UFE_TRY(dalloc(&h->op_y, ny)); h->op_y_n = ny; }
  UFE_CUDA(cudaMemcpyAsync(h->op_x, x, sizeof(double) * nx, cudaMemcpyHostToDevice, h->st));
  UFE_CUDA(cudaMemsetAsync(h->op_y, 0, sizeof(double) * ny, h->st));
  UFE_TRY(ufe_spmv_launch(h->st, F.m_loc, F.nnz, F.ptr, F.ind, F.val[which], h->op_x, F.n, h->op_y + (F.i1 - 1), F.m, nlayers));
  if (cudaMemcpyAsync(y, h->op_y, sizeof(double) * ny, cudaMemcpyDeviceToHost, h->st) != cudaSuccess ||
      cudaStreamSynchronize(h->st) != cudaSuccess) { ufe_set_error("apply_operator copy failed"); return UFE_ERR_CUDA; }
  return UFE_OK;
}

extern "C" int ufe_get_stiffness_csr(ufe_handle *h, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind, double *val,
                                     double *bb) {
  if (!h->pattern_valid) { ufe_set_error("no stiffness matrix assembled yet"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  *m_loc = h->S.m_loc; *nnz = h->S.nnz;
  if (!ind) return UFE_OK;
  UFE_CUDA(cudaMemcpy(ptr, h->S.ptr, sizeof(int) * (h->S.m_loc + 1), cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(ind, h->S.ind, sizeof(int) * h->S.nnz, cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(val, h->S.val, sizeof(double) * h->S.nnz, cudaMemcpyDeviceToHost));
  UFE_CUDA(cudaMemcpy(bb, h->S.bb, sizeof(double) * h->S.m_loc, cudaMemcpyDeviceToHost));
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// roofline helper: the Krylov MatMult kernel on the resident stiffness matrix
// ------------------------------------------------------------------------------------
extern "C" int ufe_bench_spmv(ufe_handle *h, int32_t reps, int32_t flush_l2, double *ms_per_launch,
                              double *algorithmic_bytes) {
  UFE_CUDA(cudaSetDevice(h->device));
  if (!h->pattern_valid) { ufe_set_error("no stiffness matrix assembled yet"); return UFE_ERR_INVALID; }
  if (flush_l2 && !h->flush_buf) { h->flush_bytes = (size_t)512 << 20; UFE_CUDA(cudaMalloc(&h->flush_buf, h->flush_bytes)); }
  const DevSystem &S = h->S;
  for (int w = 0; w < 3; w++) UFE_TRY(ufe_kspmv_plain(h->st, S, h->kw.pg, h->kw.t, h->kw));
  double total = 0.0;
  for (int r = 0; r < reps; r++) {
    if (flush_l2) UFE_CUDA(cudaMemsetAsync(h->flush_buf, r & 0xff, h->flush_bytes, h->st));
    cudaEventRecord(h->ev[2], h->st);
    UFE_TRY(ufe_kspmv_only(h->st, S, h->kw.pg, h->kw.t, h->kw));
    cudaEventRecord(h->ev[3], h->st);
    UFE_CUDA(cudaEventSynchronize(h->ev[3]));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    total += ms;
  }
  *ms_per_launch = total / reps;
  // B_spmv = 12 nnz + 4 (m+1) + 8 n + 8 m  (SURVEY.md 8d); n = columns touched by the owned rows
  const double ncols = (double)(S.jmax - S.jmin + 1);
  *algorithmic_bytes = 12.0 * S.nnz + 4.0 * (S.m_loc + 1) + 8.0 * ncols + 8.0 * S.m_loc;
  return UFE_OK;
}

// ------------------------------------------------------------------------------------
// L0 with the handle's communicator: the drop-in for solve_matrix_equation_CSR_PETSc as the reference calls it --
// every rank passes ITS rows of A (type_sparse_matrix_CSR_dp: rows i1..i2, local 1-based ptr, global 1-based ind),
// its slice of bb and of xx (petsc_basic.f90:32-64; call sites solve_linearised_SSA_DIVA.f90:159,
// conservation_of_mass_semiimplicit.f90:155).  Method / preconditioner / maxits / initial guess come from the handle's
// config: for the stiffness system of the handle's mesh (m = 2 nTri, rows = the rank's triangle range) every
// preconditioner is available, including the exact multifrontal one (auto, nd_lu); other systems get point Jacobi or
// the banded block solve.  Collective over the ranks of the handle.
// ------------------------------------------------------------------------------------
extern "C" int ufe_solve_matrix_equation_CSR(ufe_handle *h, const ufe_csr *A, const double *bb, double *xx, double rtol,
                                             double abstol, int32_t *n_Axb_its, int32_t *flags) {
  if (!h || !A || !bb || !xx || !A->ptr || (A->nnz > 0 && (!A->ind || !A->val))) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  if (A->m != A->n || A->m_loc < 0 || A->i1 < 1 || A->i1 - 1 + A->m_loc > A->m) { ufe_set_error("matrix and vector sub-sizes dont match!"); return UFE_ERR_INVALID; }
  UFE_CUDA(cudaSetDevice(h->device));
  const int N = A->m, m_loc = A->m_loc, r0 = A->i1 - 1, P = h->comm.nranks;
  if (P == 1 && m_loc != N) { ufe_set_error("one rank must hold all rows of the matrix"); return UFE_ERR_INVALID; }
  cudaStream_t st = h->st;
  DevSystem T;
  T.N = N; T.m_loc = m_loc; T.r1 = A->i1; T.nnz = A->nnz;
  KrylovWork kw;
  PcLU *pc = nullptr;
  int rc = UFE_OK;
  auto cleanup = [&]() {
    ufe_pclu_free(pc); ufe_krylov_free(kw);
    cudaFree(T.ptr); cudaFree(T.ind); cudaFree(T.val); cudaFree(T.valS); cudaFree(T.bb); cudaFree(T.bS); cudaFree(T.x);
  };
#define L0_TRY(call) do { rc = (call); if (rc != UFE_OK) { cleanup(); return rc; } } while (0)
  L0_TRY(dupload(&T.ptr, A->ptr, (size_t)m_loc + 1)); L0_TRY(dupload(&T.ind, A->ind, (size_t)A->nnz));
  L0_TRY(dupload(&T.val, A->val, (size_t)A->nnz)); L0_TRY(dalloc(&T.valS, (size_t)A->nnz));
  L0_TRY(dupload(&T.bb, bb, (size_t)m_loc)); L0_TRY(dalloc(&T.bS, (size_t)m_loc));
  L0_TRY(dalloc(&T.x, (size_t)N));
  if (m_loc > 0 && cudaMemcpy(T.x + r0, xx, sizeof(double) * m_loc, cudaMemcpyHostToDevice) != cudaSuccess) { cleanup(); ufe_set_error("upload of xx failed"); return UFE_ERR_CUDA; }
  int lo = A->i1, hi = A->i1 - 1;
  if (A->nnz > 0) L0_TRY(ufe_colrange(st, A->nnz, T.ind, &lo, &hi));
  T.jmin = lo; T.jmax = hi;
  HaloPlan plan;
  L0_TRY(make_plan(h, plan, N, r0, r0 + m_loc, lo - 1, hi));
  plan.mult = 1;
  if (P > 1) {                      // the strips must tile 1..m in rank order
    int next = 0;
    for (int q = 0; q < P; q++) { if (plan.own_lo[q] != next) { cleanup(); ufe_set_error("the row ranges of the ranks must be contiguous and in rank order"); return UFE_ERR_INVALID; } next = plan.own_hi[q]; }
    if (next != N) { cleanup(); ufe_set_error("the row ranges of the ranks do not cover the matrix"); return UFE_ERR_INVALID; }
  }
  Comm c2 = h->comm;                // generic CSR rows: NCCL halo exchange + all-reduce (the peer-memory path reads the blocked layout)
  c2.peer.on = 0;
  L0_TRY(ufe_krylov_alloc(kw, N, m_loc, true));
  L0_TRY(ufe_launch_scale_generic(st, T));
  const int want = h->cfg.krylov_pc;
  const bool stiff = N == 2 * h->dm.nTri && A->i1 == 2 * h->ti1 - 1 && m_loc == 2 * (h->ti2 - h->ti1 + 1);
  const bool pow2 = (P & (P - 1)) == 0;
  int used = UFE_PC_JACOBI;
  if ((want == UFE_PC_AUTO || want == UFE_PC_ND_LU) && stiff && pow2) {
    rc = ufe_pclu_setup_nd(st, T, &c2, h->dm.nTri, h->hGC.data(), h->hGC.data() + h->dm.nTri, &pc);
    if (rc == UFE_OK) { ufe_pclu_set_point_scaling(pc); used = UFE_PC_ND_LU; }
    else if (rc == UFE_ERR_CUDA || want == UFE_PC_ND_LU) { cleanup(); return rc; }
    else pc = nullptr;
  } else if (want == UFE_PC_ND_LU) {
    cleanup(); ufe_set_error("krylov_pc nd_lu needs the stiffness system of the handle's mesh (2 nTri unknowns, the rank's triangle rows)"); return UFE_ERR_INVALID;
  }
  if (!pc && (want == UFE_PC_AUTO || want == UFE_PC_BJACOBI_LU)) {
    rc = ufe_pclu_setup(st, T, 0, (size_t)24 << 30, &pc);
    if (rc == UFE_OK) used = UFE_PC_BJACOBI_LU;
    else if (rc == UFE_ERR_CUDA || want == UFE_PC_BJACOBI_LU) { cleanup(); return rc; }
    else pc = nullptr;
  }
  if (pc) L0_TRY(ufe_pclu_factor(st, T, pc));
  int its = 0, fl = 0;
  L0_TRY(ufe_krylov_run(st, T, kw, c2, P > 1 ? &plan : nullptr, h->cfg.krylov_method, rtol, abstol, h->cfg.krylov_maxits,
                        h->cfg.krylov_guess_nonzero, &its, &fl, pc));
  if (m_loc > 0 && (cudaMemcpyAsync(xx, T.x + r0, sizeof(double) * m_loc, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess)) { cleanup(); ufe_set_error("copy back of xx failed"); return UFE_ERR_CUDA; }
#undef L0_TRY
  cleanup();
  h->pc_used_l0 = used;
  if (n_Axb_its) *n_Axb_its = its;
  if (flags) *flags = fl;
  return UFE_OK;
}
extern "C" int ufe_last_l0_preconditioner(const ufe_handle *h) { return h ? h->pc_used_l0 : -1; }

// ------------------------------------------------------------------------------------
// L0: generic CSR entry points (single GPU)
// ------------------------------------------------------------------------------------
extern "C" int ufe_spmv(const ufe_csr *A, const double *x, double *y, int32_t nlayers) {
  if (!A || !x || !y) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  UFE_TRY(check_device());
  int *ptr = nullptr, *ind = nullptr; double *val = nullptr, *dx = nullptr, *dy = nullptr;
  UFE_TRY(dupload(&ptr, A->ptr, (size_t)A->m_loc + 1)); UFE_TRY(dupload(&ind, A->ind, (size_t)A->nnz));
  UFE_TRY(dupload(&val, A->val, (size_t)A->nnz)); UFE_TRY(dupload(&dx, x, (size_t)A->n * nlayers));
  UFE_TRY(dalloc(&dy, (size_t)A->m_loc * nlayers));
  int rc = ufe_spmv_launch(0, A->m_loc, A->nnz, ptr, ind, val, dx, A->n, dy, A->m_loc, nlayers);
  if (rc == UFE_OK && cudaMemcpy(y, dy, sizeof(double) * (size_t)A->m_loc * nlayers, cudaMemcpyDeviceToHost) != cudaSuccess) {
    ufe_set_error("ufe_spmv: copy back failed: %s", cudaGetErrorString(cudaGetLastError())); rc = UFE_ERR_CUDA;
  }
  cudaFree(ptr); cudaFree(ind); cudaFree(val); cudaFree(dx); cudaFree(dy);
  return rc;
}

extern "C" int ufe_krylov_solve(const ufe_csr *A, const double *b, double *x, double rtol, double abstol, int32_t method,
                                int32_t maxits, int32_t guess_nonzero, int32_t *n_its, int32_t *flags) {
  if (!A || !b || !x) { ufe_set_error("null argument"); return UFE_ERR_INVALID; }
  if (A->m != A->n || A->m_loc != A->m) { ufe_set_error("ufe_krylov_solve needs the whole square matrix on one GPU"); return UFE_ERR_INVALID; }
  UFE_TRY(check_device());
  DevSystem S;
  S.N = A->m; S.m_loc = A->m; S.r1 = 1; S.nnz = A->nnz; S.jmin = 1; S.jmax = A->n;
  UFE_TRY(dupload(&S.ptr, A->ptr, (size_t)A->m_loc + 1)); UFE_TRY(dupload(&S.ind, A->ind, (size_t)A->nnz));
  UFE_TRY(dupload(&S.val, A->val, (size_t)A->nnz)); UFE_TRY(dalloc(&S.valS, (size_t)A->nnz));
  UFE_TRY(dupload(&S.bb, b, (size_t)A->m)); UFE_TRY(dalloc(&S.bS, (size_t)A->m));
  UFE_TRY(dupload(&S.x, x, (size_t)A->m));
  KrylovWork kw;
  Comm comm;
  int rc = ufe_krylov_alloc(kw, A->m, A->m, true);
  cudaStream_t st = 0;
  if (rc == UFE_OK) rc = ufe_launch_scale_generic(st, S);
  if (rc == UFE_OK) rc = ufe_krylov_run(st, S, kw, comm, nullptr, method, rtol, abstol, maxits, guess_nonzero, n_its, flags);
  if (rc == UFE_OK && cudaMemcpy(x, S.x, sizeof(double) * A->m, cudaMemcpyDeviceToHost) != cudaSuccess) {
    ufe_set_error("ufe_krylov_solve: copy back failed"); rc = UFE_ERR_CUDA;
  }
  ufe_krylov_free(kw);
  cudaFree(S.ptr); cudaFree(S.ind); cudaFree(S.val); cudaFree(S.valS); cudaFree(S.bb); cudaFree(S.bS); cudaFree(S.x);
  return rc;
}
