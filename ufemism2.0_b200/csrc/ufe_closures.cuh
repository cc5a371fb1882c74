// Kernel-side views of the handle's device data (plain structs passed by value).
#pragma once
#include "ufe_internal.cuh"

struct DevFamilyView {
  const int *ptr, *ind;                 // ptr: local 1-based offsets of the owned rows
  const double *v0, *v1, *v2, *v3, *v4; // map/ddx/ddy  or  ddx/ddy/d2dx2/d2dxdy/d2dy2
};
static inline DevFamilyView view_of(const DevFamily &F) {
  return DevFamilyView{F.ptr, F.ind, F.val[0], F.val[1], F.val[2], F.val[3], F.val[4]};
}

struct ClosureParams {
  double zeta[UFE_NZ_MAX];
  double visc_eff_min, eta_max, eps_sq_0, n_Glen, vel_max;
  double uniform_A, m_enh_sheet, m_enh_shelf;
  double slid_delta_v, slid_beta_max, slid_Weertman_m, slid_Budd_q, slid_Budd_u, slid_ZI_p, slid_ZI_ut;
  double subgrid_exponent, Hi_min;
  double icestream_Hi, icestream_dhdx, icestream_L, icestream_m, ISMIP_HOM_L;
  int rheology, enh_transition, sliding_law, idealised_law, do_GL_subgrid_friction;
};

// full-length (global-indexed) a-grid inputs
struct VertexInputs {
  const double *V;                       // (nV,2)
  const double *Hi, *Hib, *SL, *fraction_gr, *Neff, *Ti, *tys, *alpha_sq, *beta_sq;
  const int *mask_gr, *mask_fl;
};

// full-length solver fields (type_ice_velocity_solver_DIVA, ice_model_types.f90:63-109)
struct DivaFields {
  double *u_vav_b, *v_vav_b, *u_base_b, *v_base_b, *tau_bx_b, *tau_by_b, *eta_3D_b, *u_3D_b, *v_3D_b;
  double *du_dx_a, *du_dy_a, *dv_dx_a, *dv_dy_a, *du_dz_3D_a, *dv_dz_3D_a, *eta_3D_a, *N_a;
  double *beta_a, *beta_eff_a;
  double *N_b, *dN_dx_b, *dN_dy_b, *F1_3D_b, *F2_3D_b, *beta_b, *beta_eff_b, *tau_dx_b, *tau_dy_b;
  double *u_b_prev, *v_b_prev;
  // DIVA gather records (one contiguous, sector-aligned record per triangle / vertex instead of nz strided layers: a
  // gather touches whole 32-byte sectors, and a halo is ONE contiguous message per peer and stage):
  //   rec_b [nTri][RB]: u_vav, v_vav, u_base, v_base, then per layer (du/dz, dv/dz) = tau_b zeta / max(eta_min, eta_3D_b)
  //   rec_a [nV][RA]:   N_a, beta_a, beta_eff_a, 0, then per layer (eta, F1, F2)
  double *rec_b, *rec_a;
  int RB, RA;
};

// outputs of calc_secondary_velocities (full-length device arrays)
struct SecondaryFields {
  double *u_surf_b, *v_surf_b, *uabs_surf_b, *u_base_b, *v_base_b, *uabs_base_b, *u_vav_b, *v_vav_b, *uabs_vav_b;   // (nTri)
  double *u_3D, *v_3D;                                                                                               // (nV,nz)
  double *u_surf, *v_surf, *uabs_surf, *u_base, *v_base, *uabs_base, *u_vav, *v_vav, *uabs_vav, *R_shear;            // (nV)
};
int ufe_launch_secondary_b(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P, const double *u3,
                           const double *v3, const SecondaryFields &O);
int ufe_launch_secondary_a(cudaStream_t st, int v0, int nv, int nV, int nTri, int nz, DevFamilyView ba, const double *u3,
                           const double *v3, const SecondaryFields &O);

int ufe_launch_driving_stress(cudaStream_t st, int t0, int nt, DevFamilyView ab, const double *Hi,
                              const double *Hs, double *tdx, double *tdy);
int ufe_launch_till(cudaStream_t st, int nV, const ClosureParams &P, const double *Neff, const double *phi,
                    const int *mask_land, const int *mask_gr, const int *C, const int *nC, double *tys);
int ufe_launch_pack_b(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P, const DivaFields &F);
int ufe_launch_vertex(cudaStream_t st, int is_diva, int v0, int nv, int nV, int nTri, int nz,
                      const ClosureParams &P, DevFamilyView ba, const VertexInputs &I, const DivaFields &F);
int ufe_launch_triangle(cudaStream_t st, int is_diva, int t0, int nt, int nV, int nTri, int nz,
                        const ClosureParams &P, DevFamilyView ab, const double *fraction_gr_b,
                        const DivaFields &F);
int ufe_launch_post_picard(cudaStream_t st, int t0, int nt, int nTri, int is_diva, const ClosureParams &P,
                           double relax, const double *xg, const DivaFields &F, double *partials,
                           unsigned *counter, double *out);
int ufe_launch_vel3d(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P,
                     const DivaFields &F);

// ---- assembly (ufe_assembly.cu) ----
struct AssemblyParams {
  int crossterms, pc;                    // pc: UFE_PC_*
  int bc_u[4], bc_v[4];                  // north, east, south, west
  double visc_it_relax;                  // C%visc_it_relax (config value, used by the copy BCs)
};
struct BCTables {                        // per border triangle: copy list of find_ti_copy_*
  const int *slot;                       // (nTri) -1 or slot index
  const int *copy_ti;                    // (nslots, nC_mem) 1-based, 0 = none
  const double *copy_w;                  // (nslots, nC_mem)
  int nC_mem;
};
int ufe_build_stiffness_pattern(cudaStream_t st, int t0, int nt, int nTri, const AssemblyParams &A,
                                DevFamilyView bb2, const int *TriBI, const int *TriC, const int *bc_mask,
                                DevSystem &S, int **rowkind_io);
int ufe_launch_assemble(cudaStream_t st, int t0, int nt, int nTri, const AssemblyParams &A, DevFamilyView bb2,
                        const int *TriC, const int *rowkind, const BCTables &T, const int *bc_mask,
                        const double *bc_u, const double *bc_v, const DivaFields &F, const DevSystem &S,
                        int write_x);
int ufe_launch_scale_generic(cudaStream_t st, const DevSystem &S);
