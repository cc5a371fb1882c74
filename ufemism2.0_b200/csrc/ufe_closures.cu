// Per-Picard-iteration closures of the DIVA / SSA solve, fused into gather kernels.
//
// The reference evaluates them as 71 separate CSR SpMVs + elementwise loops per Picard
// iteration.  Here every a-grid quantity of one vertex is produced by
// one thread in one pass over the vertex's M_*_b_a row (k_vertex_*), and every b-grid
// quantity of one triangle in one pass over its M_*_a_b row (k_triangle_*): each
// operator row is read once per Picard iteration, the nz layers stay in registers.
// Row sums run in the reference's order (k ascending, y = sum val*x) and the file is
// compiled with -fmad=false, so results differ from the CPU evaluation only through
// pow()/exp().
//
// Reference routines restated (src/UFEMISM/ice_dynamics/conservation_of_momentum/SSA_DIVA/):
//   calc_driving_stress, calc_horizontal_strain_rates      SSA_DIVA_utilities.f90:21-82
//   calc_vertical_shear_strain_rates                        DIVA_main.f90:375-410
//   calc_effective_viscosity                                DIVA_main.f90:412-479, SSA_main.f90:314-388
//   calc_F_integrals                                        DIVA_main.f90:481-520
//   calc_effective_basal_friction_coefficient               DIVA_main.f90:522-574
//   calc_basal_friction_coefficient + laws                  ../sliding_laws.f90:25-409
//   calc_ice_rheology_Glen, Glen viscosity                  ../../rheology/constitutive_equation.f90:25-163
//   vertical_average, integrate_from_zeta_is_one_...        src/UPSY/mesh/mesh_zeta.f90:163-283
//   apply_velocity_limits, relax, calc_L2_norm_uv           SSA_DIVA_utilities.f90:84-184
//   calc_basal_velocities, calc_basal_shear_stress, calc_3D_velocities  DIVA_main.f90:576-676
#include "ufe_closures.cuh"
#include "ufe_reduce.cuh"

#define ICE_DENSITY 910.0      // src/UPSY/basic/parameters.f90:52
#define GRAV 9.81              // :49
#define PI_REF 3.141592653589793
#define R_GAS 8.314            // :56

// ---------------------------------------------------------------------------------
__global__ void k_driving_stress(int t0, int nt, DevFamilyView ab, const double *__restrict__ Hi,
                                 const double *__restrict__ Hs, double *__restrict__ tau_dx,
                                 double *__restrict__ tau_dy) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  double Hi_b = 0.0, dx = 0.0, dy = 0.0;
  for (int k = ab.ptr[tl] - 1; k < ab.ptr[tl + 1] - 1; k++) {
    const int vj = ab.ind[k] - 1;
    Hi_b += ab.v0[k] * Hi[vj];
    dx += ab.v1[k] * Hs[vj];
    dy += ab.v2[k] * Hs[vj];
  }
  tau_dx[t0 + tl] = -ICE_DENSITY * GRAV * Hi_b * dx;
  tau_dy[t0 + tl] = -ICE_DENSITY * GRAV * Hi_b * dy;
}

// till yield stress incl. extend_till_yield_stress_to_neighbours (sliding_laws.f90:370-409);
// independent of the velocity, so evaluated once per solve instead of once per iteration.
__global__ void k_till_yield_stress(int nV, ClosureParams P, const double *__restrict__ Neff,
                                    const double *__restrict__ phi, const int *__restrict__ mask_land,
                                    const int *__restrict__ mask_gr, const int *__restrict__ C,
                                    const int *__restrict__ nC, double *__restrict__ tys) {
  const int vi = blockIdx.x * blockDim.x + threadIdx.x;
  if (vi >= nV) return;
  const double tp = tan(PI_REF / 180.0);
  double t = Neff[vi] * tp * phi[vi];
  if (mask_land[vi]) {
    bool found = false;
    double mn = 1000.0 * ICE_DENSITY * GRAV;
    for (int ci = 0; ci < nC[vi]; ci++) {
      const int vc = C[(size_t)ci * nV + vi] - 1;
      if (mask_gr[vc]) { mn = fmin(mn, Neff[vc] * tp * phi[vc]); found = true; }
    }
    t = found ? mn : P.Hi_min * ICE_DENSITY * GRAV;
  }
  tys[vi] = t;
}

__device__ __forceinline__ double flow_factor(const ClosureParams &P, double Ti, double enh) {
  double A;
  if (P.rheology == UFE_RHEO_UNIFORM) A = P.uniform_A;
  else if (Ti < 263.15) A = 1.14E-05 * exp(-6.0E+04 / (R_GAS * Ti));
  else A = 5.47E+10 * exp(-13.9E+04 / (R_GAS * Ti));
  return A * enh;
}

__device__ __forceinline__ double enhancement(const ClosureParams &P, const VertexInputs &I, int vi) {
  const bool gr = I.mask_gr[vi] != 0, fl = I.mask_fl[vi] != 0;
  if (P.enh_transition == UFE_ENH_INTERP) {
    if (I.Hi[vi] > 0.0 && I.Hib[vi] < I.SL[vi])
      return I.fraction_gr[vi] * P.m_enh_sheet + (1.0 - I.fraction_gr[vi]) * P.m_enh_shelf;
  }
  if (gr) return P.m_enh_sheet;
  if (fl) return P.m_enh_shelf;
  return 1.0;
}

__device__ __forceinline__ double sliding_beta(const ClosureParams &P, const VertexInputs &I, int vi,
                                               double u_a, double v_a, double x, double y) {
  const double uabs = sqrt(P.slid_delta_v * P.slid_delta_v + u_a * u_a + v_a * v_a);
  double beta = 0.0;
  switch (P.sliding_law) {
    case UFE_SLID_NO_SLIDING: beta = 0.0; break;
    case UFE_SLID_IDEALISED:
      switch (P.idealised_law) {
        case UFE_IDEAL_SSA_ICESTREAM: {
          // Schoof2006_icestream tau_yield (Schoof_SSA_solution.f90:36-46)
          const double f = -ICE_DENSITY * GRAV * P.icestream_Hi * P.icestream_dhdx;
          const double tys = f * pow(fabs(y / P.icestream_L), P.icestream_m);
          beta = tys / uabs;
          break;
        }
        case UFE_IDEAL_ISMIP_HOM_C:
          beta = 1000.0 + 1000.0 * sin(2.0 * PI_REF * x / P.ISMIP_HOM_L) * sin(2.0 * PI_REF * y / P.ISMIP_HOM_L);
          break;
        case UFE_IDEAL_ISMIP_HOM_D:
          beta = 1000.0 + 1000.0 * sin(2.0 * PI_REF * x / P.ISMIP_HOM_L);
          break;
        case UFE_IDEAL_ISMIP_HOM_F:
          beta = pow(P.uniform_A * 1000.0, -1.0);
          break;
      }
      break;
    case UFE_SLID_WEERTMAN:
      beta = I.beta_sq[vi] * pow(uabs, 1.0 / P.slid_Weertman_m - 1.0);
      break;
    case UFE_SLID_COULOMB:
      beta = I.tys[vi] / uabs;
      break;
    case UFE_SLID_BUDD:
      beta = I.tys[vi] * pow(uabs, P.slid_Budd_q - 1.0) / pow(P.slid_Budd_u, P.slid_Budd_q);
      break;
    case UFE_SLID_TSAI2015:
      beta = fmin(I.alpha_sq[vi] * I.Neff[vi], I.beta_sq[vi] * pow(uabs, 1.0 / P.slid_Weertman_m)) * pow(uabs, -1.0);
      break;
    case UFE_SLID_SCHOOF2005: {
      const double m = P.slid_Weertman_m, aN = I.alpha_sq[vi] * I.Neff[vi];
      beta = ((I.beta_sq[vi] * pow(uabs, 1.0 / m) * aN) /
              pow(pow(I.beta_sq[vi], m) * uabs + pow(aN, m), 1.0 / m)) * pow(uabs, -1.0);
      break;
    }
    case UFE_SLID_ZOET_IVERSON:
      beta = I.tys[vi] * pow(uabs, 1.0 / P.slid_ZI_p - 1.0) * pow(uabs + P.slid_ZI_ut, -1.0 / P.slid_ZI_p);
      break;
  }
  return fmin(P.slid_beta_max, beta);
}

// ---------------------------------------------------------------------------------
// b-grid gather record of a triangle: the four 2-D velocities the a-grid closures map, and
// calc_vertical_shear_strain_rates on the b-grid (DIVA_main.f90:375-410), du/dz_b = tau_bx zeta / max(eta_min, eta_3D_b).
// The reference evaluates that quotient per (triangle, layer) and maps it to the a-grid; evaluating it inside the vertex
// gather would repeat every fp64 division once per adjacent vertex, so it is done here, once, on the owned triangles.
// ---------------------------------------------------------------------------------
// a warp packs 32 consecutive triangles: lane = triangle while the layers are read (coalesced) and the quotients are
// formed, then the 32 records -- contiguous in memory -- leave through shared memory with coalesced stores
__global__ void __launch_bounds__(128)
k_pack_b(int t0, int nt, int nTri, int nz, ClosureParams P, DivaFields F) {
  extern __shared__ double pack_sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, RB = F.RB;
  const int tb = (blockIdx.x * 4 + w) * 32;              // first local triangle of this warp
  if (tb >= nt) return;
  double *tile = pack_sm + (size_t)w * 32 * RB;
  const int tl = tb + lane;
  if (tl < nt) {
    const int tj = t0 + tl;
    double *rec = tile + lane * RB;
    rec[0] = F.u_vav_b[tj]; rec[1] = F.v_vav_b[tj]; rec[2] = F.u_base_b[tj]; rec[3] = F.v_base_b[tj];
    const double tbx = F.tau_bx_b[tj], tby = F.tau_by_b[tj];
    for (int l = 0; l < nz; l++) {
      const double den = fmax(P.visc_eff_min, F.eta_3D_b[(size_t)l * nTri + tj]);
      rec[4 + 2 * l] = tbx * P.zeta[l] / den;
      rec[5 + 2 * l] = tby * P.zeta[l] / den;
    }
    for (int e = 4 + 2 * nz; e < RB; e++) rec[e] = 0.0;
  }
  __syncwarp();
  const int n = min(32, nt - tb) * RB;
  double *dst = F.rec_b + (size_t)(t0 + tb) * RB;
  for (int q = lane; q < n; q += 32) dst[q] = tile[q];
}

// ---------------------------------------------------------------------------------
// DIVA, a-grid: one thread per owned vertex
// ---------------------------------------------------------------------------------
template <int NZ>
__global__ void __launch_bounds__(128)
k_vertex_diva(int v0, int nv, int nV, int nTri, int nz_rt, ClosureParams P, DevFamilyView ba, VertexInputs I,
              DivaFields F) {
  const int vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= nv) return;
  const int vi = v0 + vl;
  const int nz = NZ > 0 ? NZ : nz_rt;
  constexpr int NZA = NZ > 0 ? NZ : UFE_NZ_MAX;
  double duz[NZA], dvz[NZA];
#pragma unroll
  for (int k = 0; k < NZA; k++) { duz[k] = 0.0; dvz[k] = 0.0; }
  double du_dx = 0.0, du_dy = 0.0, dv_dx = 0.0, dv_dy = 0.0, u_a = 0.0, v_a = 0.0;
  for (int k = ba.ptr[vl] - 1; k < ba.ptr[vl + 1] - 1; k++) {
    const int tj = ba.ind[k] - 1;
    const double wm = ba.v0[k], wx = ba.v1[k], wy = ba.v2[k];
    const double2 *rec = reinterpret_cast<const double2 *>(F.rec_b + (size_t)tj * F.RB);
    const double2 uv = rec[0], ub = rec[1];
    du_dx += wx * uv.x; du_dy += wy * uv.x; dv_dx += wx * uv.y; dv_dy += wy * uv.y;
    u_a += wm * ub.x; v_a += wm * ub.y;
#pragma unroll
    for (int l = 0; l < NZA; l++) {
      if (l < nz) {
        const double2 sh = rec[2 + l];
        duz[l] += wm * sh.x;
        dvz[l] += wm * sh.y;
      }
    }
  }
  F.du_dx_a[vi] = du_dx; F.du_dy_a[vi] = du_dy; F.dv_dx_a[vi] = dv_dx; F.dv_dy_a[vi] = dv_dy;

  const double nexp = P.n_Glen;
  const double enh = enhancement(P, I, vi);
  const double e1 = -1.0 / nexp, e2 = (1.0 - nexp) / (2.0 * nexp);
  const bool uniform = P.rheology == UFE_RHEO_UNIFORM;
  const double Apow_uniform = uniform ? pow(flow_factor(P, 0.0, enh), e1) : 0.0;      // the same for every layer
  double eta[NZA];
#pragma unroll
  for (int l = 0; l < NZA; l++) {
    if (l < nz) {
      F.du_dz_3D_a[(size_t)l * nV + vi] = duz[l];
      F.dv_dz_3D_a[(size_t)l * nV + vi] = dvz[l];
      const double Apow = uniform ? Apow_uniform : pow(flow_factor(P, I.Ti[(size_t)l * nV + vi], enh), e1);
      const double eps_sq = du_dx * du_dx + dv_dy * dv_dy + du_dx * dv_dy +
                            0.25 * ((du_dy + dv_dx) * (du_dy + dv_dx)) +
                            0.25 * (duz[l] * duz[l] + dvz[l] * dvz[l]) + P.eps_sq_0;
      double e = 0.5 * Apow * pow(eps_sq, e2);
      e = fmin(fmax(e, P.visc_eff_min), P.eta_max);
      eta[l] = e;
      F.eta_3D_a[(size_t)l * nV + vi] = e;
    }
  }
  // vertical_average (mesh_zeta.f90:257-283)
  double eta_vav = 0.0;
#pragma unroll
  for (int l = 0; l < NZA - 1; l++)
    if (l < nz - 1) eta_vav = eta_vav + 0.5 * (eta[l + 1] + eta[l]) * (P.zeta[l + 1] - P.zeta[l]);
  const double Hi = I.Hi[vi];
  double *ra = F.rec_a + (size_t)vi * F.RA;
  const double N_a = eta_vav * fmax((double)0.1f, Hi);      // max(0.1, Hi): default-real literal, DIVA_main.f90:468
  F.N_a[vi] = N_a;
  // F-integrals (DIVA_main.f90:481-520; integrate_from_zeta_is_one_to_zeta_is_zetap)
  const double Hd = -fmax(0.1, Hi);
  double i1 = 0.0, i2 = 0.0, F2_surf = 0.0;
  ra[4 + 3 * (nz - 1)] = eta[nz - 1]; ra[5 + 3 * (nz - 1)] = Hd * i1; ra[6 + 3 * (nz - 1)] = Hd * i2;
  if (nz == 1) F2_surf = Hd * i2;
  // the integrands zeta / eta and zeta^2 / eta of a layer enter two trapezoids: each quotient is evaluated once
  double f1a = P.zeta[nz - 1] / eta[nz - 1], f2a = (P.zeta[nz - 1] * P.zeta[nz - 1]) / eta[nz - 1];
#pragma unroll
  for (int l = NZA - 2; l >= 0; l--) {
    if (l < nz - 1) {
      const double dz = P.zeta[l + 1] - P.zeta[l];
      const double f1b = P.zeta[l] / eta[l];
      const double f2b = (P.zeta[l] * P.zeta[l]) / eta[l];
      i1 = i1 - 0.5 * (f1a + f1b) * dz;
      i2 = i2 - 0.5 * (f2a + f2b) * dz;
      f1a = f1b; f2a = f2b;
      ra[4 + 3 * l] = eta[l]; ra[5 + 3 * l] = Hd * i1; ra[6 + 3 * l] = Hd * i2;
      if (l == 0) F2_surf = Hd * i2;
    }
  }
  // basal friction (sliding_laws.f90:25-81) and beta_eff (DIVA_main.f90:538-550)
  const double beta = sliding_beta(P, I, vi, u_a, v_a, I.V[vi], I.V[(size_t)nV + vi]);
  const double beta_eff = (P.sliding_law == UFE_SLID_NO_SLIDING) ? 1.0 / F2_surf : beta / (1.0 + beta * F2_surf);
  F.beta_a[vi] = beta;
  F.beta_eff_a[vi] = beta_eff;
  ra[0] = N_a; ra[1] = beta; ra[2] = beta_eff; ra[3] = 0.0;
}

// rows with more than three entries (singular three-point fit -> widened neighbourhood): accumulate per layer.
// Rare, so the per-layer accumulators live in local memory (rolled loops) and do not set the kernel's register count.
template <int NZA>
__device__ __noinline__ void k_triangle_diva_general(int tl, int ti, int nV, int nTri, int nz, const ClosureParams &P,
                                                     const DevFamilyView &ab, const double *__restrict__ fraction_gr_b,
                                                     const DivaFields &F) {
  double eb[NZA], f1[NZA], f2[NZA];
#pragma unroll 1
  for (int l = 0; l < NZA; l++) { eb[l] = 0.0; f1[l] = 0.0; f2[l] = 0.0; }
  double N_b = 0.0, dNx = 0.0, dNy = 0.0, beta_b = 0.0, beta_eff_b = 0.0;
  for (int k = ab.ptr[tl] - 1; k < ab.ptr[tl + 1] - 1; k++) {
    const int vj = ab.ind[k] - 1;
    const double wm = ab.v0[k], wx = ab.v1[k], wy = ab.v2[k];
    const double *ra = F.rec_a + (size_t)vj * F.RA;
    const double Na = ra[0];
    N_b += wm * Na; dNx += wx * Na; dNy += wy * Na;
    beta_b += wm * ra[1];
    beta_eff_b += wm * ra[2];
#pragma unroll 1
    for (int l = 0; l < NZA; l++) {
      if (l < nz) {
        eb[l] += wm * ra[4 + 3 * l];
        f1[l] += wm * ra[5 + 3 * l];
        f2[l] += wm * ra[6 + 3 * l];
      }
    }
  }
  if (P.do_GL_subgrid_friction) beta_eff_b = beta_eff_b * pow(fraction_gr_b[ti], P.subgrid_exponent);
  F.N_b[ti] = N_b; F.dN_dx_b[ti] = dNx; F.dN_dy_b[ti] = dNy;
  F.beta_b[ti] = beta_b; F.beta_eff_b[ti] = beta_eff_b;
#pragma unroll 1
  for (int l = 0; l < NZA; l++) {
    if (l < nz) {
      F.eta_3D_b[(size_t)l * nTri + ti] = eb[l];
      F.F1_3D_b[(size_t)l * nTri + ti] = f1[l];
      F.F2_3D_b[(size_t)l * nTri + ti] = f2[l];
    }
  }
}


// DIVA, b-grid: one thread per owned triangle
template <int NZ>
__global__ void __launch_bounds__(128)
k_triangle_diva(int t0, int nt, int nV, int nTri, int nz_rt, ClosureParams P, DevFamilyView ab,
                const double *__restrict__ fraction_gr_b, DivaFields F) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  const int nz = NZ > 0 ? NZ : nz_rt;
  constexpr int NZA = NZ > 0 ? NZ : UFE_NZ_MAX;
  const int k0 = ab.ptr[tl] - 1, k1 = ab.ptr[tl + 1] - 1;
  if (k1 - k0 == 3) {
    // the usual row (the triangle's own three vertices): weights and columns stay in registers and the layer loop
    // is the outer one -- nine independent gathers per layer in flight, no per-layer accumulator arrays, so the
    // kernel runs at twice the occupancy.  Same summation order as the general path: ((0 + w0 f0) + w1 f1) + w2 f2.
    const int v0 = ab.ind[k0] - 1, v1 = ab.ind[k0 + 1] - 1, v2 = ab.ind[k0 + 2] - 1;
    const double w0 = ab.v0[k0], w1 = ab.v0[k0 + 1], w2 = ab.v0[k0 + 2];
    const double *r0 = F.rec_a + (size_t)v0 * F.RA, *r1 = F.rec_a + (size_t)v1 * F.RA, *r2 = F.rec_a + (size_t)v2 * F.RA;
    {
      const double x0 = ab.v1[k0], x1 = ab.v1[k0 + 1], x2 = ab.v1[k0 + 2];
      const double y0 = ab.v2[k0], y1 = ab.v2[k0 + 1], y2 = ab.v2[k0 + 2];
      const double N0 = r0[0], N1 = r1[0], N2 = r2[0];
      double N_b = 0.0, dNx = 0.0, dNy = 0.0, beta_b = 0.0, beta_eff_b = 0.0;
      N_b += w0 * N0; N_b += w1 * N1; N_b += w2 * N2;
      dNx += x0 * N0; dNx += x1 * N1; dNx += x2 * N2;
      dNy += y0 * N0; dNy += y1 * N1; dNy += y2 * N2;
      beta_b += w0 * r0[1]; beta_b += w1 * r1[1]; beta_b += w2 * r2[1];
      beta_eff_b += w0 * r0[2]; beta_eff_b += w1 * r1[2]; beta_eff_b += w2 * r2[2];
      if (P.do_GL_subgrid_friction) beta_eff_b = beta_eff_b * pow(fraction_gr_b[ti], P.subgrid_exponent);
      F.N_b[ti] = N_b; F.dN_dx_b[ti] = dNx; F.dN_dy_b[ti] = dNy;
      F.beta_b[ti] = beta_b; F.beta_eff_b[ti] = beta_eff_b;
    }
    // two layers = six doubles = three 16-byte loads per record
    const double2 *q0 = reinterpret_cast<const double2 *>(r0 + 4), *q1 = reinterpret_cast<const double2 *>(r1 + 4), *q2 = reinterpret_cast<const double2 *>(r2 + 4);
#pragma unroll 2
    for (int l = 0; l + 1 < nz; l += 2) {
      const int o = 3 * (l >> 1);
      const double2 A0 = q0[o], B0 = q0[o + 1], C0 = q0[o + 2];      // (eta_l, F1_l) (F2_l, eta_l+1) (F1_l+1, F2_l+1)
      const double2 A1 = q1[o], B1 = q1[o + 1], C1 = q1[o + 2];
      const double2 A2 = q2[o], B2 = q2[o + 1], C2 = q2[o + 2];
      double e = 0.0, a = 0.0, b = 0.0;
      e += w0 * A0.x; e += w1 * A1.x; e += w2 * A2.x;
      a += w0 * A0.y; a += w1 * A1.y; a += w2 * A2.y;
      b += w0 * B0.x; b += w1 * B1.x; b += w2 * B2.x;
      size_t ob = (size_t)l * nTri + ti;
      F.eta_3D_b[ob] = e; F.F1_3D_b[ob] = a; F.F2_3D_b[ob] = b;
      e = 0.0; a = 0.0; b = 0.0;
      e += w0 * B0.y; e += w1 * B1.y; e += w2 * B2.y;
      a += w0 * C0.x; a += w1 * C1.x; a += w2 * C2.x;
      b += w0 * C0.y; b += w1 * C1.y; b += w2 * C2.y;
      ob += nTri;
      F.eta_3D_b[ob] = e; F.F1_3D_b[ob] = a; F.F2_3D_b[ob] = b;
    }
    if (nz & 1) {
      const int l = nz - 1, o = 4 + 3 * l;
      double e = 0.0, a = 0.0, b = 0.0;
      e += w0 * r0[o]; e += w1 * r1[o]; e += w2 * r2[o];
      a += w0 * r0[o + 1]; a += w1 * r1[o + 1]; a += w2 * r2[o + 1];
      b += w0 * r0[o + 2]; b += w1 * r1[o + 2]; b += w2 * r2[o + 2];
      const size_t ob = (size_t)l * nTri + ti;
      F.eta_3D_b[ob] = e; F.F1_3D_b[ob] = a; F.F2_3D_b[ob] = b;
    }
    return;
  }
  k_triangle_diva_general<NZA>(tl, ti, nV, nTri, nz, P, ab, fraction_gr_b, F);
}

// ---------------------------------------------------------------------------------
// SSA (SSA_main.f90:314-430): 2-D viscosity with vertically averaged A, beta_b directly
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_vertex_ssa(int v0, int nv, int nV, int nz, ClosureParams P, DevFamilyView ba, VertexInputs I, DivaFields F) {
  const int vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= nv) return;
  const int vi = v0 + vl;
  double du_dx = 0.0, du_dy = 0.0, dv_dx = 0.0, dv_dy = 0.0, u_a = 0.0, v_a = 0.0;
  for (int k = ba.ptr[vl] - 1; k < ba.ptr[vl + 1] - 1; k++) {
    const int tj = ba.ind[k] - 1;
    const double uj = F.u_vav_b[tj], vj = F.v_vav_b[tj];
    du_dx += ba.v1[k] * uj; du_dy += ba.v2[k] * uj; dv_dx += ba.v1[k] * vj; dv_dy += ba.v2[k] * vj;
    u_a += ba.v0[k] * uj; v_a += ba.v0[k] * vj;
  }
  F.du_dx_a[vi] = du_dx; F.du_dy_a[vi] = du_dy; F.dv_dx_a[vi] = dv_dx; F.dv_dy_a[vi] = dv_dy;
  const double enh = enhancement(P, I, vi);
  double A_vav = 0.0;
  double Aprev = flow_factor(P, (P.rheology == UFE_RHEO_UNIFORM) ? 0.0 : I.Ti[vi], enh);
  for (int l = 0; l < nz - 1; l++) {
    const double Anext = flow_factor(P, (P.rheology == UFE_RHEO_UNIFORM) ? 0.0 : I.Ti[(size_t)(l + 1) * nV + vi], enh);
    A_vav = A_vav + 0.5 * (Anext + Aprev) * (P.zeta[l + 1] - P.zeta[l]);
    Aprev = Anext;
  }
  const double nexp = P.n_Glen;
  const double eps_sq = du_dx * du_dx + dv_dy * dv_dy + du_dx * dv_dy + 0.25 * ((du_dy + dv_dx) * (du_dy + dv_dx)) + P.eps_sq_0;
  double e = 0.5 * pow(A_vav, -1.0 / nexp) * pow(eps_sq, (1.0 - nexp) / (2.0 * nexp));
  e = fmin(fmax(e, P.visc_eff_min), P.eta_max);
  F.eta_3D_a[vi] = e;                                    // SSA%eta_a
  F.N_a[vi] = e * fmax(0.1, I.Hi[vi]);                   // 0.1_dp here, SSA_main.f90:382
  F.beta_a[vi] = sliding_beta(P, I, vi, u_a, v_a, I.V[vi], I.V[(size_t)nV + vi]);
}

__global__ void __launch_bounds__(128)
k_triangle_ssa(int t0, int nt, ClosureParams P, DevFamilyView ab, const double *__restrict__ fraction_gr_b,
               DivaFields F) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  double N_b = 0.0, dNx = 0.0, dNy = 0.0, beta_b = 0.0;
  for (int k = ab.ptr[tl] - 1; k < ab.ptr[tl + 1] - 1; k++) {
    const int vj = ab.ind[k] - 1;
    const double Na = F.N_a[vj];
    N_b += ab.v0[k] * Na; dNx += ab.v1[k] * Na; dNy += ab.v2[k] * Na;
    beta_b += ab.v0[k] * F.beta_a[vj];
  }
  if (P.do_GL_subgrid_friction) beta_b = beta_b * pow(fraction_gr_b[ti], P.subgrid_exponent);
  F.N_b[ti] = N_b; F.dN_dx_b[ti] = dNx; F.dN_dy_b[ti] = dNy;
  F.beta_b[ti] = beta_b; F.beta_eff_b[ti] = beta_b;     // the SSA passes beta_b to the linearised solve
}

// ---------------------------------------------------------------------------------
// after the linear solve: velocity limits, relaxation, basal velocities, basal shear
// stress and the Picard residual sums, in one pass (a15 + a16 of SURVEY.md 8a)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_post_picard(int t0, int nt, int nTri, int is_diva, ClosureParams P, double relax, const double *__restrict__ xg,
              DivaFields F, double *partials, unsigned *counter, double *out) {
  double acc[2] = {0.0, 0.0};
  for (int tl = blockIdx.x * blockDim.x + threadIdx.x; tl < nt; tl += gridDim.x * blockDim.x) {
    const int ti = t0 + tl;
    double u = xg[2 * (size_t)ti], v = xg[2 * (size_t)ti + 1];
    const double uabs = sqrt(u * u + v * v);
    if (uabs > P.vel_max) { u = u * P.vel_max / uabs; v = v * P.vel_max / uabs; }
    const double up = F.u_b_prev[ti], vp = F.v_b_prev[ti];
    u = (relax * u) + ((1.0 - relax) * up);
    v = (relax * v) + ((1.0 - relax) * vp);
    F.u_vav_b[ti] = u; F.v_vav_b[ti] = v;
    if (is_diva) {
      if (P.sliding_law == UFE_SLID_NO_SLIDING) { F.u_base_b[ti] = 0.0; F.v_base_b[ti] = 0.0; }
      else {
        const double dn = 1.0 + F.beta_b[ti] * F.F2_3D_b[ti];     // F2_3D_b(ti,1)
        F.u_base_b[ti] = u / dn; F.v_base_b[ti] = v / dn;
      }
      F.tau_bx_b[ti] = u * F.beta_eff_b[ti];
      F.tau_by_b[ti] = v * F.beta_eff_b[ti];
    }
    acc[0] += (u - up) * (u - up); acc[0] += (v - vp) * (v - vp);
    acc[1] += (u + up) * (u + up); acc[1] += (v + vp) * (v + vp);
  }
  reduce_publish<2>(acc, partials, counter, out);
}

__global__ void k_vel3d(int t0, int nt, int nTri, int nz, ClosureParams P, DivaFields F) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  for (int l = 0; l < nz; l++) {
    const double f1 = F.F1_3D_b[(size_t)l * nTri + ti];
    if (P.sliding_law == UFE_SLID_NO_SLIDING) {
      F.u_3D_b[(size_t)l * nTri + ti] = F.tau_bx_b[ti] * f1;
      F.v_3D_b[(size_t)l * nTri + ti] = F.tau_by_b[ti] * f1;
    } else {
      const double g = 1.0 + F.beta_b[ti] * f1;
      F.u_3D_b[(size_t)l * nTri + ti] = F.u_base_b[ti] * g;
      F.v_3D_b[(size_t)l * nTri + ti] = F.v_base_b[ti] * g;
    }
  }
}

// ---------------------------------------------------------------------------------
// calc_secondary_velocities (conservation_of_momentum_main.f90:176-245), the step right after the
// solve: surface / base / vertically averaged velocities on the b-grid from u_3D_b, v_3D_b, then the
// b->a maps (map_b_a_3D x2, map_b_a_2D x6) fused into one pass over the shared M_map_b_a pattern,
// absolute values and the slide/shear ratio.
// ---------------------------------------------------------------------------------
__global__ void k_secondary_b(int t0, int nt, int nTri, int nz, ClosureParams P, const double *__restrict__ u3,
                              const double *__restrict__ v3, SecondaryFields O) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int ti = t0 + tl;
  const double us = u3[ti], vs = v3[ti];
  const double ub = u3[(size_t)(nz - 1) * nTri + ti], vb = v3[(size_t)(nz - 1) * nTri + ti];
  O.u_surf_b[ti] = us; O.v_surf_b[ti] = vs; O.uabs_surf_b[ti] = sqrt(us * us + vs * vs);
  O.u_base_b[ti] = ub; O.v_base_b[ti] = vb; O.uabs_base_b[ti] = sqrt(ub * ub + vb * vb);
  double ua = 0.0, va = 0.0;                    // vertical_average, mesh_zeta.f90:257-283
  for (int k = 0; k < nz - 1; k++) {
    const double dz = P.zeta[k + 1] - P.zeta[k];
    ua = ua + 0.5 * (u3[(size_t)(k + 1) * nTri + ti] + u3[(size_t)k * nTri + ti]) * dz;
    va = va + 0.5 * (v3[(size_t)(k + 1) * nTri + ti] + v3[(size_t)k * nTri + ti]) * dz;
  }
  O.u_vav_b[ti] = ua; O.v_vav_b[ti] = va; O.uabs_vav_b[ti] = sqrt(ua * ua + va * va);
}

template <int NZ>
__global__ void __launch_bounds__(128)
k_secondary_a(int v0, int nv, int nV, int nTri, int nz_rt, DevFamilyView ba, const double *__restrict__ u3,
              const double *__restrict__ v3, SecondaryFields O) {
  const int vl = blockIdx.x * blockDim.x + threadIdx.x;
  if (vl >= nv) return;
  const int vi = v0 + vl, nz = NZ > 0 ? NZ : nz_rt;
  double us = 0.0, vs = 0.0, ub = 0.0, vb = 0.0, ua = 0.0, va = 0.0;
  const int k0 = ba.ptr[vl] - 1, k1 = ba.ptr[vl + 1] - 1;
  for (int l = 0; l < nz; l++) {                 // map_b_a_3D, one layer at a time in the row's entry order
    double su = 0.0, sv = 0.0;
    for (int k = k0; k < k1; k++) {
      const int tj = ba.ind[k] - 1;
      const double w = ba.v0[k];
      su += w * u3[(size_t)l * nTri + tj];
      sv += w * v3[(size_t)l * nTri + tj];
    }
    O.u_3D[(size_t)l * nV + vi] = su; O.v_3D[(size_t)l * nV + vi] = sv;
  }
  for (int k = k0; k < k1; k++) {                // map_b_a_2D of the six b-grid fields
    const int tj = ba.ind[k] - 1;
    const double w = ba.v0[k];
    us += w * O.u_surf_b[tj]; vs += w * O.v_surf_b[tj];
    ub += w * O.u_base_b[tj]; vb += w * O.v_base_b[tj];
    ua += w * O.u_vav_b[tj];  va += w * O.v_vav_b[tj];
  }
  O.u_surf[vi] = us; O.v_surf[vi] = vs; O.u_base[vi] = ub; O.v_base[vi] = vb; O.u_vav[vi] = ua; O.v_vav[vi] = va;
  const double as = sqrt(us * us + vs * vs), ab = sqrt(ub * ub + vb * vb), aa = sqrt(ua * ua + va * va);
  O.uabs_surf[vi] = as; O.uabs_base[vi] = ab; O.uabs_vav[vi] = aa;
  O.R_shear[vi] = (ab + 0.1) / (as + 0.1);
}

// ---------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------
int ufe_launch_driving_stress(cudaStream_t st, int t0, int nt, DevFamilyView ab, const double *Hi,
                              const double *Hs, double *tdx, double *tdy) {
  if (nt <= 0) return UFE_OK;
  k_driving_stress<<<ufe_div_up(nt, 256), 256, 0, st>>>(t0, nt, ab, Hi, Hs, tdx, tdy);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_till(cudaStream_t st, int nV, const ClosureParams &P, const double *Neff, const double *phi,
                    const int *mask_land, const int *mask_gr, const int *C, const int *nC, double *tys) {
  k_till_yield_stress<<<ufe_div_up(nV, 256), 256, 0, st>>>(nV, P, Neff, phi, mask_land, mask_gr, C, nC, tys);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_pack_b(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P, const DivaFields &F) {
  if (nt <= 0) return UFE_OK;
  const size_t smem = (size_t)128 * F.RB * sizeof(double);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) { UFE_CUDA(cudaFuncSetAttribute(k_pack_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
  k_pack_b<<<ufe_div_up(nt, 128), 128, smem, st>>>(t0, nt, nTri, nz, P, F);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

int ufe_launch_vertex(cudaStream_t st, int is_diva, int v0, int nv, int nV, int nTri, int nz,
                      const ClosureParams &P, DevFamilyView ba, const VertexInputs &I, const DivaFields &F) {
  if (nv <= 0) return UFE_OK;
  const int blocks = ufe_div_up(nv, 128);
  if (!is_diva) k_vertex_ssa<<<blocks, 128, 0, st>>>(v0, nv, nV, nz, P, ba, I, F);
  else if (nz == 12) k_vertex_diva<12><<<blocks, 128, 0, st>>>(v0, nv, nV, nTri, nz, P, ba, I, F);
  else k_vertex_diva<0><<<blocks, 128, 0, st>>>(v0, nv, nV, nTri, nz, P, ba, I, F);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_triangle(cudaStream_t st, int is_diva, int t0, int nt, int nV, int nTri, int nz,
                        const ClosureParams &P, DevFamilyView ab, const double *fraction_gr_b,
                        const DivaFields &F) {
  if (nt <= 0) return UFE_OK;
  const int blocks = ufe_div_up(nt, 128);
  if (!is_diva) k_triangle_ssa<<<blocks, 128, 0, st>>>(t0, nt, P, ab, fraction_gr_b, F);
  else if (nz == 12) k_triangle_diva<12><<<blocks, 128, 0, st>>>(t0, nt, nV, nTri, nz, P, ab, fraction_gr_b, F);
  else k_triangle_diva<0><<<blocks, 128, 0, st>>>(t0, nt, nV, nTri, nz, P, ab, fraction_gr_b, F);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_post_picard(cudaStream_t st, int t0, int nt, int nTri, int is_diva, const ClosureParams &P,
                           double relax, const double *xg, const DivaFields &F, double *partials,
                           unsigned *counter, double *out) {
  k_post_picard<<<UFE_RED_BLOCKS, UFE_RED_THREADS, 0, st>>>(t0, nt, nTri, is_diva, P, relax, xg, F, partials,
                                                            counter, out);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_vel3d(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P,
                     const DivaFields &F) {
  if (nt <= 0) return UFE_OK;
  k_vel3d<<<ufe_div_up(nt, 256), 256, 0, st>>>(t0, nt, nTri, nz, P, F);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

int ufe_launch_secondary_b(cudaStream_t st, int t0, int nt, int nTri, int nz, const ClosureParams &P, const double *u3,
                           const double *v3, const SecondaryFields &O) {
  if (nt <= 0) return UFE_OK;
  k_secondary_b<<<ufe_div_up(nt, 256), 256, 0, st>>>(t0, nt, nTri, nz, P, u3, v3, O);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
int ufe_launch_secondary_a(cudaStream_t st, int v0, int nv, int nV, int nTri, int nz, DevFamilyView ba, const double *u3,
                           const double *v3, const SecondaryFields &O) {
  if (nv <= 0) return UFE_OK;
  if (nz == 12) k_secondary_a<12><<<ufe_div_up(nv, 128), 128, 0, st>>>(v0, nv, nV, nTri, nz, ba, u3, v3, O);
  else k_secondary_a<0><<<ufe_div_up(nv, 128), 128, 0, st>>>(v0, nv, nV, nTri, nz, ba, u3, v3, O);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}
