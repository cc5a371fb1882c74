/*
 * ufe_diva.h -- C ABI of the B200-native DIVA/SSA ice-velocity solve.
 *
 * Drop-in boundary for ONE path of IMAU-paleo/UFEMISM2.0: the DIVA/SSA velocity solve.
 * The reference has no FFI; the interface this replaces is its Fortran module-procedure
 * call surface.  Each entry point cites the reference procedure it stands in for
 * (paths relative to the reference root).  The Fortran host binds these with
 * ISO_C_BINDING (INTEGRATION.md shows the shim); tests bind them with ctypes.
 *
 * Conventions (identical to the reference so that patterns stay bit-exact):
 *   - all arrays column-major, contiguous, fp64 / int32;
 *   - all indices stored 1-based; 0 = "none" in TriC, padding in C / iTri;
 *   - Fortran LOGICAL masks are passed as int32 (0 / non-zero);
 *   - CSR = type_sparse_matrix_CSR_dp: ptr = process-local 1-based offsets for rows
 *     i1..i2 (m_loc+1 entries), ind = global 1-based columns (unsorted), val;
 *   - unknown ordering of the stiffness matrix: row = 2*(ti-1)+uv (tiuv2n).
 *   - host arrays are borrowed for the duration of a call; the handle owns all device
 *     memory; nothing is retained past a call except what ufe_diva_create is given.
 *   - every function returns 0 on success, >0 on error (never aborts);
 *     ufe_last_error_string() describes the last error on the calling thread.  The
 *     Fortran shim maps non-zero to crash() (control_resources_and_error_messaging.f90:377).
 *   - collective semantics as in the reference (pure SPMD): with nranks > 1 every rank
 *     calls every function in the same order.  One process per GPU.
 */
#ifndef UFE_DIVA_H
#define UFE_DIVA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ---------------------------------------------------------------- */
enum {
  UFE_OK = 0,
  UFE_ERR_INVALID = 1,        /* bad argument / unknown config choice (reference: crash('unknown ...')) */
  UFE_ERR_CUDA = 2,           /* CUDA runtime / NCCL failure, or no usable sm_100 device */
  UFE_ERR_PICARD_DIVERGED = 3,/* DIVA_main.f90:222-230: relax < 0.05 or eps0^2 > 1e-5 */
  UFE_ERR_OPERATOR = 4,       /* neighbourhood grew past the local stack while building operators */
  UFE_ERR_KRYLOV_BREAKDOWN = 5
};

/* flags returned in ufe_solve_info.flags (warnings, not errors; DIVA_main.f90:246-249) */
enum {
  UFE_FLAG_PICARD_MAXIT = 1,  /* 'viscosity iteration failed to converge within N iterations' */
  UFE_FLAG_KRYLOV_MAXIT = 2,  /* a linear solve hit maxits (the reference never checks KSPGetConvergedReason) */
  UFE_FLAG_KRYLOV_DIVERGED = 4,
  UFE_FLAG_NEGATIVE_HI = 8    /* calc_dHi_dt: 'encountered negative values for Hi_tplusdt - time step too large?' */
};

/* ---- type_sparse_matrix_CSR_dp (src/UPSY/basic/CSR_sparse_matrix_type.f90:15-38) -- */
typedef struct ufe_csr {
  int32_t m, n;               /* global size */
  int32_t m_loc, n_loc;       /* owned rows / columns */
  int32_t i1, i2, j1, j2;     /* owned row / column ranges, 1-based inclusive */
  int32_t nnz;
  const int32_t *ptr;         /* (m_loc+1) local 1-based offsets */
  const int32_t *ind;         /* (nnz) global 1-based columns */
  const double *val;          /* (nnz) */
} ufe_csr;

/* ---- type_mesh subset (src/UPSY/types/mesh_types.f90:16-279) --------------------- */
typedef struct ufe_mesh {
  int32_t nV, nTri, nC_mem, nz;
  double xmin, xmax, ymin, ymax;
  const double *V;            /* (nV,2) */
  const int32_t *Tri;         /* (nTri,3) */
  const int32_t *TriC;        /* (nTri,3) */
  const int32_t *C;           /* (nV,nC_mem) */
  const int32_t *nC;          /* (nV) */
  const int32_t *iTri;        /* (nV,nC_mem) */
  const int32_t *niTri;       /* (nV) */
  const int32_t *VBI;         /* (nV) */
  const int32_t *TriBI;       /* (nTri) */
  const double *TriGC;        /* (nTri,2) */
  const double *zeta;         /* (nz) */
  /* Operator matrices of calc_all_matrix_operators_mesh
   * (mesh_disc_calc_matrix_operators_2D.f90:26-59).  Each family shares one pattern.
   * Pass NULL to have the library build the family on the GPU (same algorithm:
   * BFS neighbourhood + weighted least squares).  Order within a family:
   *   a_b: map, ddx, ddy;  b_a: map, ddx, ddy;  b_b (2nd order): ddx, ddy, d2dx2, d2dxdy, d2dy2. */
  const ufe_csr *M_a_b[3];
  const ufe_csr *M_b_a[3];
  const ufe_csr *M2_b_b[5];
} ufe_mesh;

/* ---- config keys read by the path (src/UFEMISM/basic/model_configuration.f90) ----
 * Member names are the reference's C%<name>; string choices are passed as codes. */
enum { UFE_BC_INFINITE = 1, UFE_BC_ZERO = 2, UFE_BC_PERIODIC_ISMIP_HOM = 3, UFE_BC_INFINITE_SSA_ICESTREAM = 4 };
enum { UFE_SLID_NO_SLIDING = 0, UFE_SLID_IDEALISED = 1, UFE_SLID_WEERTMAN = 2, UFE_SLID_COULOMB = 3,
       UFE_SLID_BUDD = 4, UFE_SLID_TSAI2015 = 5, UFE_SLID_SCHOOF2005 = 6, UFE_SLID_ZOET_IVERSON = 7 };
enum { UFE_IDEAL_NONE = 0, UFE_IDEAL_SSA_ICESTREAM = 1, UFE_IDEAL_ISMIP_HOM_C = 2, UFE_IDEAL_ISMIP_HOM_D = 3,
       UFE_IDEAL_ISMIP_HOM_F = 5 };
enum { UFE_RHEO_UNIFORM = 0, UFE_RHEO_HUYBRECHTS1992 = 1 };
enum { UFE_ENH_SEPARATE = 0, UFE_ENH_INTERP = 1 };
enum { UFE_KRYLOV_BICGSTAB = 0, UFE_KRYLOV_GMRES = 1 };
/* UFE_PC_BJACOBI_LU: block Jacobi over contiguous row ranges (like PETSc's per-rank blocks) with an
 * exact block-tridiagonal LU solve inside each block (the reference uses ILU(0) there). */
enum { UFE_PC_JACOBI = 0, UFE_PC_BJACOBI2 = 1, UFE_PC_BJACOBI_LU = 2,
       UFE_PC_AUTO = 3 /* ND_LU when the number of ranks is 1, 2, 4 or 8 and its fronts fit the device memory (env UFE_AUTO_ND=0
                          skips it), else BJACOBI_LU when its dense blocks fit the memory budget, else BJACOBI2;
                          ufe_solve_info.krylov_pc_used reports the choice */,
       UFE_PC_ND_LU = 4 /* exact: multifrontal nested-dissection LU of the whole matrix (csrc/ufe_nd_numeric.cu), factorised
                           once per Picard iteration (krylov_pc_lag applies), sub-trees distributed over 1, 2, 4 or 8 ranks.
                           With fresh factors the linear solve is one Richardson step x = M^-1 b checked with the
                           reference's stopping rule on the true scaled residual (n_Axb_its = 1); the Krylov method only
                           continues from that x if the check fails */ };

typedef struct ufe_config {
  int32_t do_include_SSADIVA_crossterms;           /* :276 */
  double visc_it_norm_dUV_tol;                     /* :307 */
  int32_t visc_it_nit;                             /* :308 */
  double visc_it_relax;                            /* :309 */
  double visc_eff_min;                             /* :310 */
  double vel_max;                                  /* :311 */
  double stress_balance_PETSc_rtol;                /* :312 */
  double stress_balance_PETSc_abstol;              /* :313 */
  int32_t BC_u[4], BC_v[4];                        /* north, east, south, west  (:316-323) */
  int32_t choice_sliding_law;                      /* :329 */
  int32_t choice_idealised_sliding_law;            /* :330 */
  double slid_Weertman_m, slid_Budd_q_plastic, slid_Budd_u_threshold, slid_ZI_p, slid_ZI_ut; /* :333-337 */
  int32_t do_GL_subgrid_friction;                  /* :340 */
  int32_t do_subgrid_friction_on_A_grid;           /* :344 (only .false. supported) */
  double subgrid_friction_exponent_on_B_grid;      /* :345 */
  double slid_beta_max, slid_delta_v;              /* :348-349 */
  double Hi_min;                                   /* :437 */
  double Glens_flow_law_exponent;                  /* :591 */
  double Glens_flow_law_epsilon_sq_0;              /* :592 */
  int32_t choice_ice_rheology_Glen;                /* :595 */
  double uniform_Glens_flow_factor;                /* :596 */
  int32_t choice_enhancement_factor_transition;    /* :599 */
  double m_enh_sheet, m_enh_shelf;                 /* :600-601 */
  double refgeo_idealised_SSA_icestream_Hi, refgeo_idealised_SSA_icestream_dhdx,
         refgeo_idealised_SSA_icestream_L, refgeo_idealised_SSA_icestream_m;   /* :190-193 */
  double refgeo_idealised_ISMIP_HOM_L;             /* :195 */
  /* extensions (no reference key; defaults = reference behaviour where one exists) */
  int32_t krylov_method;        /* UFE_KRYLOV_* ; reference: PETSc default GMRES(30) */
  int32_t krylov_pc;            /* UFE_PC_*     ; reference: block-Jacobi / ILU(0)   */
  int32_t krylov_maxits;        /* PETSc default 10000 */
  int32_t krylov_guess_nonzero; /* 0 = KSP default (x0 = 0, petsc_basic.f90:99-128) */
  int32_t krylov_pc_lag;        /* UFE_PC_BJACOBI_LU: reuse a factorisation across Picard iterations until a solve
                                 * needs more than this many Krylov iterations; 0 = factorise every iteration */
  int32_t krylov_pc_strip_only; /* several ranks.  0: systems of at most 131072 unknowns (env UFE_REDUNDANT_MAX_UNKNOWNS) are
                                 * not partitioned at all -- every rank solves the whole system redundantly and
                                 * bit-identically without communication (every kernel is launch-latency-bound at that
                                 * size); larger ones are row-partitioned, and UFE_PC_BJACOBI_LU replicates the exact
                                 * factorisation on every rank when the whole system's dense blocks fit, else one strip
                                 * block per rank.  1: always partition the rows, one strip block per rank (PETSc's
                                 * bjacobi structure).  Read by ufe_diva_create. */
} ufe_config;

/* ---- inputs read from type_ice_model / type_bed_roughness_model ------------------
 * (src/UFEMISM/types/ice_model_types.f90:208+; bed_roughness_model_types.f90:65-85).
 * All (nV) unless noted; full-length (global) arrays on every rank. */
typedef struct ufe_ice_inputs {
  const double *Hi, *Hs, *Hib, *SL;
  const double *fraction_gr;            /* (nV)   */
  const double *fraction_gr_b;          /* (nTri) */
  const double *effective_pressure;
  const int32_t *mask_grounded_ice, *mask_floating_ice, *mask_icefree_land;
  const double *Ti;                     /* (nV,nz); read only for Huybrechts1992 */
  const double *till_friction_angle, *alpha_sq, *beta_sq;
  /* optional prescribed-velocity BC (solve_DIVA optional args, DIVA_main.f90:99-101); NULL = absent */
  const int32_t *BC_prescr_mask_b;      /* (nTri) */
  const double *BC_prescr_u_b, *BC_prescr_v_b;
} ufe_ice_inputs;

/* ---- type_ice_velocity_solver_DIVA subset (ice_model_types.f90:63-109) ------------
 * inout: the seven restart fields carried across solves (DIVA_main.f90:840-846).
 * out (may be NULL to skip the copy): fields read by set_ice_velocities_to_DIVA_results
 * (conservation_of_momentum_main.f90:486-500).  Full-length arrays. */
typedef struct ufe_diva_state {
  double *u_vav_b, *v_vav_b, *tau_bx_b, *tau_by_b;   /* (nTri)   inout */
  double *eta_3D_b;                                  /* (nTri,nz) inout */
  double *u_base_b, *v_base_b;                       /* (nTri)   inout */
  double *u_3D_b, *v_3D_b;                           /* (nTri,nz) out */
  double *du_dx_a, *du_dy_a, *dv_dx_a, *dv_dy_a;     /* (nV) out */
  double *du_dz_3D_a, *dv_dz_3D_a;                   /* (nV,nz) out */
  double *eta_3D_a;                                  /* (nV,nz) out */
  double *basal_friction_coefficient_a;              /* (nV) out (ice%basal_friction_coefficient) */
} ufe_diva_state;

/* type_ice_velocity_solver_SSA subset (ice_model_types.f90:30-61) */
typedef struct ufe_ssa_state {
  double *u_b, *v_b;                                 /* (nTri) inout */
  double *basal_friction_coefficient_a;              /* (nV) out, may be NULL */
} ufe_ssa_state;

/* outputs of calc_secondary_velocities (type_ice_model members, ice_model_types.f90); every pointer may
 * be NULL (that field is then not copied back).  Full-length arrays. */
typedef struct ufe_secondary_velocities {
  double *u_surf_b, *v_surf_b, *uabs_surf_b, *u_base_b, *v_base_b, *uabs_base_b, *u_vav_b, *v_vav_b, *uabs_vav_b; /* (nTri) */
  double *u_3D, *v_3D;                                                                                             /* (nV,nz) */
  double *u_surf, *v_surf, *uabs_surf, *u_base, *v_base, *uabs_base, *u_vav, *v_vav, *uabs_vav, *R_shear;          /* (nV) */
} ufe_secondary_velocities;

typedef struct ufe_solve_info {
  int32_t n_visc_its, n_Axb_its;   /* solve_DIVA outputs, DIVA_main.f90:97-98 */
  int32_t flags;                   /* UFE_FLAG_* */
  double L2_uv;                    /* last Picard residual (calc_L2_norm_uv) */
  double visc_it_relax_applied, Glens_flow_law_epsilon_sq_0_applied;
  /* device-time split, milliseconds (CUDA events) */
  double ms_total, ms_closures, ms_assembly, ms_krylov, ms_h2d, ms_d2h;
  int64_t gpu_launches;            /* kernels launched by this library during the call */
  int32_t krylov_pc_used;          /* UFE_PC_* actually applied (resolves UFE_PC_AUTO) */
  int32_t reserved;                /* several ranks: 0 NCCL, 1 peer memory inside the Krylov loop, 2 redundant (not partitioned) */
} ufe_solve_info;

/* multi-GPU communicator description: one process per GPU, NCCL underneath.
 * nccl_unique_id: 128 bytes from ufe_comm_get_unique_id() on rank 0, broadcast by the
 * host (MPI_Bcast in the Fortran driver).  nranks == 1: pass NULL for the whole struct. */
typedef struct ufe_comm {
  int32_t rank, nranks, device;    /* device = CUDA ordinal to use (local rank) */
  const char *nccl_unique_id;
} ufe_comm;

typedef struct ufe_handle ufe_handle;

const char *ufe_last_error_string(void);
int ufe_comm_get_unique_id(char id_out[128]);
int ufe_version(void);
/* sizeof(ufe_solve_info) as this library was built: a binding asserts it against its own declaration of the struct
 * (the entry points memset and fill the whole struct) */
int ufe_sizeof_solve_info(void);

/* partition_list (src/UPSY/basic/mpi_parallelisation/mpi_distributed_memory.f90:42-68) */
void ufe_partition_list(int32_t ntot, int32_t i, int32_t n, int32_t *i1, int32_t *i2);

/* L0 -- replaces solve_matrix_equation_CSR_PETSc (src/UPSY/basic/petsc_basic.f90:32-64;
 * call site solve_linearised_SSA_DIVA.f90:159) without a handle: one GPU, A holds all rows, point Jacobi
 * (ufe_solve_matrix_equation_CSR below is the distributed form with every preconditioner).
 * x in: initial guess (used only if guess_nonzero), out: solution.
 * method/pc: UFE_KRYLOV_* / UFE_PC_JACOBI. */
int ufe_krylov_solve(const ufe_csr *A, const double *b, double *x, double rtol, double abstol,
                     int32_t method, int32_t maxits, int32_t guess_nonzero, int32_t *n_its,
                     int32_t *flags);

/* L0 as the reference calls it -- replaces solve_matrix_equation_CSR_PETSc( A_CSR, bb, xx, rtol, abstol, n_Axb_its)
 * (petsc_basic.f90:32-64; call sites solve_linearised_SSA_DIVA.f90:159 and conservation_of_mass_semiimplicit.f90:155)
 * on the ranks of a handle: every rank passes ITS rows of A (i1..i2; ptr: m_loc + 1 local 1-based offsets; ind: global
 * 1-based columns), its slice bb(m_loc) and its slice xx(m_loc) (in: initial guess, used only with
 * krylov_guess_nonzero; out: solution).  The row ranges must tile 1..m in rank order (partition_list does).  Krylov
 * method, preconditioner, maxits come from the handle's ufe_config: for the stiffness system of the handle's mesh
 * (m = 2 nTri, the rank's rows = its triangle range) every preconditioner is available, including the exact
 * multifrontal one (UFE_PC_AUTO, UFE_PC_ND_LU); other systems get the banded block solve or point Jacobi.
 * Collective; 'matrix and vector sub-sizes dont match!' (petsc_basic.f90:91) for inconsistent sizes.
 * ufe_last_l0_preconditioner: the UFE_PC_* the last call applied (resolves UFE_PC_AUTO). */
int ufe_solve_matrix_equation_CSR(ufe_handle *h, const ufe_csr *A, const double *bb, double *xx, double rtol, double abstol,
                                  int32_t *n_Axb_its, int32_t *flags);
int ufe_last_l0_preconditioner(const ufe_handle *h);

/* multiply_CSR_matrix_with_vector_1D / _2D
 * (src/UPSY/basic/CSR_matrix_algebra/CSR_matrix_vector_multiplication.f90:198,336).
 * x: (n, nlayers) over all columns, y: (m_loc, nlayers). Single-GPU helper. */
int ufe_spmv(const ufe_csr *A, const double *x, double *y, int32_t nlayers);

/* lifecycle -- initialise_DIVA_solver / allocate_DIVA_solver (DIVA_main.f90:37,752):
 * uploads the mesh, builds or receives the operator CSRs, builds BC tables, halo lists
 * and the NCCL communicator. */
int ufe_diva_create(const ufe_mesh *mesh, const ufe_config *cfg, const ufe_comm *comm,
                    ufe_handle **out);
int ufe_diva_destroy(ufe_handle *h);
int ufe_diva_set_config(ufe_handle *h, const ufe_config *cfg);

/* L2 -- replaces solve_DIVA (DIVA_main.f90:88-262; call site
 * conservation_of_momentum_main.f90:142-144).  Host buffers in, host buffers out. */
int ufe_diva_solve(ufe_handle *h, const ufe_ice_inputs *ice, ufe_diva_state *state,
                   ufe_solve_info *info);
/* replaces solve_SSA (SSA_main.f90:87-242) */
int ufe_ssa_solve(ufe_handle *h, const ufe_ice_inputs *ice, ufe_ssa_state *state,
                  ufe_solve_info *info);

/* device-resident variants (same computation, no host<->device field traffic): inputs
 * are uploaded once, the solve is repeated on resident data, results fetched on demand. */
int ufe_diva_upload(ufe_handle *h, const ufe_ice_inputs *ice, const ufe_diva_state *state);
int ufe_diva_solve_resident(ufe_handle *h, ufe_solve_info *info);
int ufe_diva_download(ufe_handle *h, ufe_diva_state *state);
/* initialise_DIVA_solver with choice_initial_velocity = 'zero' (DIVA_main.f90:60-68) on the
 * resident state: the seven restart fields are zeroed on the device (no host traffic). */
int ufe_diva_reset_state(ufe_handle *h);

/* SURVEY.md 8(f) rank 1 -- replaces set_ice_velocities_to_DIVA_results + calc_secondary_velocities
 * (conservation_of_momentum_main.f90:470-510, 176-245), the step right after solve_DIVA, on the
 * device-resident u_3D_b / v_3D_b of the most recent DIVA solve: surface, base and vertically
 * averaged velocities on the b- and a-grid, u_3D / v_3D on the a-grid, absolute values, R_shear. */
int ufe_calc_secondary_velocities(ufe_handle *h, ufe_secondary_velocities *out);

/* ---- ice-thickness rates of change (conservation of mass) ----------------
 * The caller's other PETSc call site (conservation_of_mass_semiimplicit.f90:161) and the explicit
 * scheme it builds on, on the same handle / device as the velocity solve, so that u_vav_b, v_vav_b
 * can stay resident between solve_DIVA and the thickness update of the predictor-corrector step. */
enum { UFE_BC_H_INFINITE = 1, UFE_BC_H_ZERO = 2 };   /* C%BC_H_*: 'infinite' | 'zero' */

/* type_mesh members (src/UPSY/types/mesh_types.f90) read by calc_ice_flux_divergence_matrix_upwind
 * (conservation_of_mass_utilities.f90:21-131) and map_velocities_from_b_to_c_2D
 * (map_velocities_to_c_grid.f90:17-69), as built by construct_mesh_edges (edges/mesh_edges.f90:19-194),
 * calc_Voronoi_cell_areas, calc_connection_widths, calc_connection_lengths (mesh_secondary.f90:137-366). */
typedef struct ufe_mesh_edges {
  int32_t nE;
  const int32_t *VE;                   /* (nV,nC_mem) edge of connection ci */
  const int32_t *ETri;                 /* (nE,2) [til, tir], 0 = none */
  const double *A;                     /* (nV) Voronoi cell areas */
  const double *Cw, *D_x, *D_y, *D;    /* (nV,nC_mem) */
} ufe_mesh_edges;

typedef struct ufe_thickness_config {
  double dHi_semiimplicit_fs;          /* model_configuration.f90:356 */
  double dHi_PETSc_rtol, dHi_PETSc_abstol;  /* :357-358 */
  int32_t BC_H[4];                     /* north, east, south, west (:361-364), UFE_BC_H_* */
  double dt_ice_max, dt_ice_min;       /* :391-392 */
  double Hi_min;                       /* :437 */
  int32_t krylov_method, krylov_maxits;/* extensions as in ufe_config (reference: PETSc GMRES(30), 10000) */
} ufe_thickness_config;

/* dummy arguments of calc_dHi_dt_explicit / calc_dHi_dt_semiimplicit
 * (conservation_of_mass_explicit.f90:23-72, conservation_of_mass_semiimplicit.f90:24-98).  All (nV)
 * unless noted, full-length.  mask_noice is a Fortran LOGICAL passed as int32. */
typedef struct ufe_thickness_fields {
  const double *Hi, *Hb, *SL;
  const double *u_vav_b, *v_vav_b;     /* (nTri); both NULL = the device-resident velocities of the handle's
                                          most recent solve_DIVA / solve_SSA */
  const double *SMB, *BMB, *LMB, *fraction_margin, *dHi_dt_target;
  const int32_t *mask_noice;
  const int32_t *BC_prescr_mask;       /* optional pair, NULL = absent */
  const double *BC_prescr_Hi;
  double *AMB, *dHi_dt, *Hi_tplusdt, *divQ;   /* out; any may be NULL (not copied back) */
} ufe_thickness_fields;

/* upload the edge / Voronoi data once per mesh (the handle must come from ufe_diva_create on the same mesh) */
int ufe_mesh_set_edges(ufe_handle *h, const ufe_mesh_edges *edges);
/* replaces calc_dHi_dt_explicit (conservation_of_mass_explicit.f90:23-138); dt inout (flux-limited) */
int ufe_calc_dHi_dt_explicit(ufe_handle *h, const ufe_thickness_config *cfg, ufe_thickness_fields *f, double *dt);
/* replaces calc_dHi_dt_semiimplicit (conservation_of_mass_semiimplicit.f90:24-173); dt is not changed
 * there.  n_Axb_its: Krylov iterations of the (1 + dt f_s M_divQ) solve; flags: UFE_FLAG_KRYLOV_*.
 * The Krylov solve starts from x0 = 0 like the reference's (KSP default; its 'initial guess'
 * Hi + dt*dHi_dt is never handed to PETSc as non-zero, petsc_basic.f90:99-128). */
int ufe_calc_dHi_dt_semiimplicit(ufe_handle *h, const ufe_thickness_config *cfg, ufe_thickness_fields *f, double dt,
                                 int32_t *n_Axb_its, int32_t *flags);
/* replaces calc_dHi_dt (conservation_of_mass_main.f90:22-109), the entry the predictor-corrector scheme calls:
 * dispatch on C%choice_ice_integration_method, then clip negative thicknesses (UFE_FLAG_NEGATIVE_HI in *flags
 * where the reference warns), AMB = AMB + (Hi_tplusdt - Hi)/dt - dHi_dt, dHi_dt = (Hi_tplusdt - Hi)/dt.
 * dt inout (only the explicit scheme changes it). */
enum { UFE_THK_NONE = 0, UFE_THK_EXPLICIT = 1, UFE_THK_SEMI_IMPLICIT = 2 };   /* 'none' | 'explicit' | 'semi-implicit' */
int ufe_calc_dHi_dt(ufe_handle *h, const ufe_thickness_config *cfg, int32_t choice_ice_integration_method,
                    ufe_thickness_fields *f, double *dt, int32_t *n_Axb_its, int32_t *flags);
/* which = 0: M_divQ of the most recent call; 1: the stiffness matrix AA and load vector bb of the most
 * recent semi-implicit call.  Reference CSR layout (row vi = [vi, C(vi,1..nC)]).  Query sizes with ind == NULL. */
int ufe_get_thickness_csr(ufe_handle *h, int32_t which, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind,
                          double *val, double *bb);
/* ---- SURVEY.md 8(f) rank 1, last part: replaces calc_vertical_velocities
 * (src/UFEMISM/ice_dynamics/conservation_of_mass/vertical_velocities.f90:18-210), called after dHi_dt is known.
 * Reads the device-resident u_3D_b / v_3D_b of the most recent solve_DIVA and ice%u_3D / v_3D of the
 * ufe_calc_secondary_velocities call that followed it; needs ufe_mesh_set_edges.  All (nV) unless noted. */
typedef struct ufe_vertical_velocity_inputs {
  const double *Hi, *Hib, *dHb_dt, *dHi_dt, *BMB;
  const int32_t *mask_grounded_ice, *mask_floating_ice;
  const double *dzeta_dx_ak, *dzeta_dy_ak, *dzeta_dz_ak;    /* (nV,nz), calc_zeta_gradients (zeta_gradients.f90:129-131) */
} ufe_vertical_velocity_inputs;
int ufe_calc_vertical_velocities(ufe_handle *h, const ufe_vertical_velocity_inputs *in, double *w_3D /* (nV,nz) out */);
/* mesh%M_ddx_a_a (which = 0) / M_ddy_a_a (1), calc_matrix_operators_mesh_a_a
 * (mesh_disc_calc_matrix_operators_2D.f90:60-196), as built on the device.  Query sizes with ind == NULL. */
int ufe_mesh_get_operator_a_a(ufe_handle *h, int32_t which, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind,
                              double *val);

/* device time (CUDA events on the handle's stream) of the most recent thickness call, milliseconds:
 * ms[0] inputs host->device, [1] k_thk_divq (M_divQ, divQ, explicit dH/dt, time-step limit), [2] rest of the explicit
 * scheme + border BCs + system assembly, [3] Krylov solve, [4] finishing kernel, [5] outputs device->host; and the
 * algorithmic bytes of one k_thk_divq launch (104 B per connection + 96 B per vertex). */
int ufe_get_thickness_timing(ufe_handle *h, double ms[6], double *divq_algorithmic_bytes);

/* L1 -- replaces solve_SSA_DIVA_linearised (solve_linearised_SSA_DIVA.f90:23-178; call
 * sites DIVA_main.f90:189-192, SSA_main.f90:178-181).  Full-length (nTri) arrays.
 * u_b, v_b inout; u_b_prev, v_b_prev out (the gathered previous solution). */
int ufe_ssa_diva_linearised(ufe_handle *h, double *u_b, double *v_b, const double *N_b,
                            const double *dN_dx_b, const double *dN_dy_b,
                            const double *basal_friction_coefficient_b, const double *tau_dx_b,
                            const double *tau_dy_b, double *u_b_prev, double *v_b_prev,
                            double PETSc_rtol, double PETSc_abstol, int32_t *n_Axb_its,
                            const int32_t *BC_prescr_mask_b, const double *BC_prescr_u_b,
                            const double *BC_prescr_v_b);

/* operator access -- mesh%M_* as built/held on the device.
 * family: 0 a_b, 1 b_a, 2 b_b(2nd order); which: index within the family.
 * Query sizes with ind == NULL; then pass buffers (ptr: m_loc+1, ind/val: nnz). */
int ufe_mesh_get_operator(ufe_handle *h, int32_t family, int32_t which, int32_t *m_loc,
                          int32_t *nnz, int32_t *ptr, int32_t *ind, double *val);
/* map_a_b_2D/3D, ddx_a_b_2D, ... (src/UPSY/mesh/discretisation/mesh_disc_apply_operators.f90
 * :121-431): y = M x, x full-length (n, nlayers), y full-length (m, nlayers) with the
 * owned rows filled. */
int ufe_mesh_apply_operator(ufe_handle *h, int32_t family, int32_t which, const double *x,
                            double *y, int32_t nlayers);
/* the stiffness matrix and right-hand side of the most recent linearised solve, in the
 * reference's CSR layout (rows 2*ti1-1 .. 2*ti2). Query with ind == NULL first. */
int ufe_get_stiffness_csr(ufe_handle *h, int32_t *m_loc, int32_t *nnz, int32_t *ptr, int32_t *ind,
                          double *val, double *bb);

/* ---- staged for the multifrontal exact solver on wide meshes ---------------------------
 * Host-only symbolic analysis: nested dissection (recursive coordinate bisection on the triangle centroids) of the
 * block graph of the stiffness matrix -- one node per triangle's 2x2 (u,v) block.  Needs no GPU.  bptr / bind: block
 * pattern, 0-based CSR over triangles.  The tree is stored in post-order (children before parents, root last); node i
 * eliminates the triangles sep[0..n_sep) and its dense front is [sep; bnd] x [sep; bnd]; up[k] is the position of
 * bnd[k] in the parent's [sep; bnd] list (extend-add of the Schur complement); the entry map gives, per block entry
 * of the analysed pattern, the front it is assembled into and its block row / column there. */
typedef struct ufe_nd_tree ufe_nd_tree;
int ufe_nd_analyse(int32_t nT, const double *centroid_x, const double *centroid_y, const int32_t *bptr,
                   const int32_t *bind, int32_t leaf_triangles, ufe_nd_tree **out);
int ufe_nd_tree_info(const ufe_nd_tree *T, int32_t *n_nodes, int32_t *n_levels, int32_t *max_front,
                     double *padded_front_bytes);
int ufe_nd_tree_node(const ufe_nd_tree *T, int32_t i, int32_t *level, int32_t *parent, int32_t *n_sep, int32_t *n_bnd,
                     const int32_t **sep, const int32_t **bnd, const int32_t **up);
int ufe_nd_tree_entry_map(const ufe_nd_tree *T, const int32_t **node, const int32_t **row, const int32_t **col);
/* distribution of the tree over nranks (1, 2, 4, 8, ...) ranks: owner[i] = rank that factorises node i (post-order index as
 * in ufe_nd_tree_node); rank r owns the sub-tree below the r-th node of level log2(nranks) and the nodes above it on
 * its left spine -- the tree analogue of the reference's strip partition (mesh_parallelisation.f90:90-127) */
int ufe_nd_tree_owners(const ufe_nd_tree *T, int32_t nranks, int32_t *owner);
void ufe_nd_tree_free(ufe_nd_tree *T);

/* Numeric phase on the device (csrc/ufe_nd_numeric.cu): exact solve of A x = b, the system the reference hands to
 * PETSc at solve_linearised_SSA_DIVA.f90:159 (petsc_basic.f90:32-64), by a multifrontal factorisation over the tree --
 * one dense front per tree node with its own shape, right-looking block LU (32-wide pivot blocks, 64-wide trailing
 * updates), Schur complements pulled into the parents in a fixed order.  ptr / ind / val: scalar CSR of A in the
 * reference's own convention (type_sparse_matrix_CSR_dp, CSR_sparse_matrix_type.f90:15-38: ptr(1) = 1, 1-based column
 * indices), N = 2 nTri rows, whose block pattern is the analysed one; host arrays, borrowed for the call.  `factor` may
 * be called again with new values (same pattern: one analysis per mesh, one factorisation per Picard iteration);
 * `solve` returns x = A^-1 b after n_refine steps of iterative refinement with the unscaled matrix and the relative
 * residual |b - A x| / |b|.  No CPU fallback: create fails with UFE_ERR_CUDA without a device. */
typedef struct ufe_nd_solver ufe_nd_solver;
int ufe_nd_solver_create(const ufe_nd_tree *T, int32_t N, const int32_t *ptr, const int32_t *ind, ufe_nd_solver **out);
int ufe_nd_solver_factor(ufe_nd_solver *S, const double *val);
int ufe_nd_solver_solve(ufe_nd_solver *S, const double *b, double *x, int32_t n_refine, double *relres);
int ufe_nd_solver_info(const ufe_nd_solver *S, double *factor_ms, double *solve_ms, double *front_bytes, double *factor_flops);
void ufe_nd_solver_free(ufe_nd_solver *S);

/* benchmark / roofline helpers: time `reps` launches of the stiffness-matrix SpMV (the
 * Krylov MatMult kernel) on the handle's resident matrix with CUDA events on the
 * launching stream; returns average ms per launch and the algorithmic bytes per launch
 * (12*nnz + 4*(m+1) + 8*n + 8*m). */
int ufe_bench_spmv(ufe_handle *h, int32_t reps, int32_t flush_l2, double *ms_per_launch,
                   double *algorithmic_bytes);
int ufe_get_ownership(ufe_handle *h, int32_t *vi1, int32_t *vi2, int32_t *ti1, int32_t *ti2);

#ifdef __cplusplus
}
#endif
#endif /* UFE_DIVA_H */
