"""Host-side mirror of the reference's call surface into the velocity-solver modules.

Same names, argument meaning and error behaviour as the Fortran procedures, on top of
the C ABI (include/ufe_diva.h):

  initialise_DIVA_solver / solve_DIVA / solve_SSA     DIVA_main.f90:37,88 ; SSA_main.f90:87
  remap_DIVA_solver                                   DIVA_main.f90:264
  solve_SSA_DIVA_linearised                           solve_linearised_SSA_DIVA.f90:23
  solve_matrix_equation_CSR_PETSc                     src/UPSY/basic/petsc_basic.f90:32
  multiply_CSR_matrix_with_vector_1D/_2D              CSR_matrix_vector_multiplication.f90:198,336
  map_a_b_2D/3D, ddx_a_b_2D, ... , ddy_b_a_2D          mesh_disc_apply_operators.f90:121-431
  partition_list                                      mpi_distributed_memory.f90:42
  calc_dHi_dt_explicit / calc_dHi_dt_semiimplicit     conservation_of_mass_explicit.f90:23 ; _semiimplicit.f90:24

``crash(...)`` in the reference becomes ``UfeError``; the "viscosity iteration failed to
converge" warning becomes ``info.flags & PICARD_MAXIT``.
"""
from __future__ import annotations

import ctypes as ct
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import UfeError, check, vp
from .config import (BC_CODES, BC_H_CODES, ENH_CODES, ICE_INTEGRATION_CODES, IDEALISED_SLIDING_CODES, RHEOLOGY_CODES, SLIDING_CODES, Config)
from .mesh_types import Mesh

PICARD_MAXIT, KRYLOV_MAXIT, KRYLOV_DIVERGED = 1, 2, 4
KRYLOV_METHODS = {"bicgstab": 0, "gmres": 1}
KRYLOV_PCS = {"jacobi": 0, "bjacobi2": 1, "bjacobi_lu": 2, "auto": 3, "nd_lu": 4}
FAMILIES = {"a_b": (0, ("map", "ddx", "ddy")), "b_a": (1, ("map", "ddx", "ddy")),
            "b_b": (2, ("ddx", "ddy", "d2dx2", "d2dxdy", "d2dy2"))}


def _code(table, key, what):
    try:
        return table[key]
    except KeyError:
        raise UfeError(1, f'unknown {what} "{key}"!') from None


def config_struct(C: Config) -> capi.ufe_config:
    s = capi.ufe_config()
    s.do_include_SSADIVA_crossterms = int(C.do_include_SSADIVA_crossterms)
    for k in ("visc_it_norm_dUV_tol", "visc_it_nit", "visc_it_relax", "visc_eff_min", "vel_max",
              "stress_balance_PETSc_rtol", "stress_balance_PETSc_abstol", "slid_Weertman_m",
              "slid_Budd_q_plastic", "slid_Budd_u_threshold", "slid_ZI_p", "slid_ZI_ut",
              "subgrid_friction_exponent_on_B_grid", "slid_beta_max", "slid_delta_v", "Hi_min",
              "Glens_flow_law_exponent", "Glens_flow_law_epsilon_sq_0", "uniform_Glens_flow_factor",
              "m_enh_sheet", "m_enh_shelf", "refgeo_idealised_SSA_icestream_Hi",
              "refgeo_idealised_SSA_icestream_dhdx", "refgeo_idealised_SSA_icestream_L",
              "refgeo_idealised_SSA_icestream_m", "refgeo_idealised_ISMIP_HOM_L"):
        setattr(s, k, getattr(C, k))
    for i, side in enumerate(("north", "east", "south", "west")):
        s.BC_u[i] = _code(BC_CODES, getattr(C, f"BC_u_{side}"), "choice_BC_u")
        s.BC_v[i] = _code(BC_CODES, getattr(C, f"BC_v_{side}"), "choice_BC_v")
    s.choice_sliding_law = _code(SLIDING_CODES, C.choice_sliding_law, "choice_sliding_law")
    s.choice_idealised_sliding_law = _code(IDEALISED_SLIDING_CODES, C.choice_idealised_sliding_law,
                                           "choice_idealised_sliding_law")
    s.do_GL_subgrid_friction = int(C.do_GL_subgrid_friction)
    s.do_subgrid_friction_on_A_grid = int(C.do_subgrid_friction_on_A_grid)
    s.choice_ice_rheology_Glen = _code(RHEOLOGY_CODES, C.choice_ice_rheology_Glen, "choice_ice_rheology_Glen")
    s.choice_enhancement_factor_transition = _code(ENH_CODES, C.choice_enhancement_factor_transition,
                                                   "choice_enhancement_factor_transition")
    if C.choice_flow_law != "Glen":
        raise UfeError(1, f'unknown choice_flow_law "{C.choice_flow_law}"!')
    s.krylov_method = _code(KRYLOV_METHODS, C.b200_krylov_method, "b200_krylov_method")
    s.krylov_pc = _code(KRYLOV_PCS, C.b200_krylov_pc, "b200_krylov_pc")
    s.krylov_maxits = C.b200_krylov_maxits
    s.krylov_guess_nonzero = int(C.b200_krylov_guess_nonzero)
    s.krylov_pc_lag = int(C.b200_krylov_pc_lag)
    s.krylov_pc_strip_only = int(C.b200_krylov_pc_strip_only)
    return s


def partition_list(ntot, i, n):
    i1, i2 = ct.c_int32(), ct.c_int32()
    capi.lib().ufe_partition_list(ntot, i, n, ct.byref(i1), ct.byref(i2))
    return i1.value, i2.value


@dataclass
class CSRMatrix:
    """type_sparse_matrix_CSR_dp (CSR_sparse_matrix_type.f90:15-38)."""
    m: int
    n: int
    i1: int
    i2: int
    ptr: np.ndarray
    ind: np.ndarray
    val: np.ndarray


def multiply_CSR_matrix_with_vector(A: CSRMatrix, x: np.ndarray) -> np.ndarray:
    """1-D (x shape (n,)) or 2-D (x shape (n,nz), column-major) SpMV on the GPU."""
    keep = []
    s = capi.csr_struct(A.m, A.n, A.i1, A.i2, A.ptr, A.ind, A.val, keep)
    x = np.asfortranarray(x, dtype=np.float64)
    nl = 1 if x.ndim == 1 else x.shape[1]
    y = np.zeros((A.i2 - A.i1 + 1,) if x.ndim == 1 else (A.i2 - A.i1 + 1, nl), order="F")
    check(capi.lib().ufe_spmv(ct.byref(s), vp(x), vp(y), nl))
    return y


def solve_matrix_equation_CSR(A: CSRMatrix, bb, xx, rtol, abstol, method="bicgstab", maxits=10000,
                              guess_nonzero=False):
    """Replaces solve_matrix_equation_CSR_PETSc (petsc_basic.f90:32-64). Returns (x, n_its, flags)."""
    keep = []
    s = capi.csr_struct(A.m, A.n, A.i1, A.i2, A.ptr, A.ind, A.val, keep)
    b = np.ascontiguousarray(bb, dtype=np.float64)
    x = np.array(xx, dtype=np.float64, copy=True)
    its, fl = ct.c_int32(), ct.c_int32()
    check(capi.lib().ufe_krylov_solve(ct.byref(s), vp(b), vp(x), ct.c_double(rtol), ct.c_double(abstol),
                                      KRYLOV_METHODS[method], maxits, int(guess_nonzero), ct.byref(its),
                                      ct.byref(fl)))
    return x, its.value, fl.value


@dataclass
class SolveInfo:
    n_visc_its: int
    n_Axb_its: int
    flags: int
    L2_uv: float
    visc_it_relax_applied: float
    Glens_flow_law_epsilon_sq_0_applied: float
    ms_total: float
    ms_closures: float
    ms_assembly: float
    ms_krylov: float
    ms_h2d: float
    ms_d2h: float
    gpu_launches: int
    krylov_pc_used: int = 0
    reserved: int = 0


def _info(s: capi.ufe_solve_info) -> SolveInfo:
    return SolveInfo(*[getattr(s, f[0]) for f in capi.ufe_solve_info._fields_])


class DIVASolver:
    """type_ice_velocity_solver_DIVA (+ _SSA) bound to one mesh; owns the device handle."""

    STATE_FIELDS_B = ("u_vav_b", "v_vav_b", "tau_bx_b", "tau_by_b", "u_base_b", "v_base_b")

    def __init__(self, mesh: Mesh, C: Config, comm=None, operators=None):
        """comm: None (one GPU) or (rank, nranks, device, nccl_unique_id_bytes).
        operators: None -> built on the GPU; or dict name -> CSRMatrix holding this rank's rows."""
        self.mesh, self.C = mesh, C
        self._keep = []
        m = capi.ufe_mesh()
        m.nV, m.nTri, m.nC_mem, m.nz = mesh.nV, mesh.nTri, mesh.nC_mem, mesh.nz
        m.xmin, m.xmax, m.ymin, m.ymax = mesh.xmin, mesh.xmax, mesh.ymin, mesh.ymax
        for name in ("V", "Tri", "TriC", "C", "nC", "iTri", "niTri", "VBI", "TriBI", "TriGC", "zeta"):
            a = getattr(mesh, name)
            a = np.asfortranarray(a)
            self._keep.append(a)
            setattr(m, name, vp(a))
        if operators is not None:
            self._csr_structs = []
            for fam, attr in (("a_b", "M_a_b"), ("b_a", "M_b_a"), ("b_b", "M2_b_b")):
                arr = getattr(m, attr)
                for i, w in enumerate(FAMILIES[fam][1]):
                    nm = ("M2_%s_b_b" % w) if fam == "b_b" else ("M_%s_%s" % (w, fam))
                    A = operators[nm]
                    s = capi.csr_struct(A.m, A.n, A.i1, A.i2, A.ptr, A.ind, A.val, self._keep)
                    self._csr_structs.append(s)
                    arr[i] = ct.pointer(s)
        cs = config_struct(C)
        cm = None
        if comm is not None:
            rank, nranks, device, uid = comm
            self._uid = ct.create_string_buffer(uid, 128) if uid is not None else None
            cm = capi.ufe_comm(rank, nranks, device, ct.cast(self._uid, ct.c_char_p) if self._uid else None)
        self._h = ct.c_void_p()
        check(capi.lib().ufe_diva_create(ct.byref(m), ct.byref(cs), ct.byref(cm) if cm else None,
                                         ct.byref(self._h)))
        nT, nV, nz = mesh.nTri, mesh.nV, mesh.nz
        # allocate_DIVA_solver (DIVA_main.f90:752-804) with 'zero' initial velocities
        z = np.zeros
        self.u_vav_b, self.v_vav_b = z(nT), z(nT)
        self.tau_bx_b, self.tau_by_b = z(nT), z(nT)
        self.u_base_b, self.v_base_b = z(nT), z(nT)
        self.eta_3D_b = z((nT, nz), order="F")
        self.u_3D_b, self.v_3D_b = z((nT, nz), order="F"), z((nT, nz), order="F")
        self.du_dx_a, self.du_dy_a, self.dv_dx_a, self.dv_dy_a = z(nV), z(nV), z(nV), z(nV)
        self.du_dz_3D_a, self.dv_dz_3D_a = z((nV, nz), order="F"), z((nV, nz), order="F")
        self.eta_3D_a = z((nV, nz), order="F")
        self.basal_friction_coefficient_a = z(nV)
        # SSA state
        self.u_b, self.v_b = z(nT), z(nT)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().ufe_diva_destroy(self._h)
            self._h = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- config
    def set_config(self, C: Config):
        cs = config_struct(C)
        check(capi.lib().ufe_diva_set_config(self._h, ct.byref(cs)))
        self.C = C

    def ownership(self):
        v = [ct.c_int32() for _ in range(4)]
        check(capi.lib().ufe_get_ownership(self._h, *[ct.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # ---- structs
    def _ice_struct(self, ice, bc=None):
        s = capi.ufe_ice_inputs()
        keep = []

        def put(name, arr, dtype):
            if arr is None:
                return
            a = np.asfortranarray(arr, dtype=dtype)
            keep.append(a)
            setattr(s, name, vp(a))
        for n in ("Hi", "Hs", "Hib", "SL", "fraction_gr", "fraction_gr_b", "effective_pressure", "Ti",
                  "till_friction_angle", "alpha_sq", "beta_sq"):
            put(n, getattr(ice, n, None), np.float64)
        for n in ("mask_grounded_ice", "mask_floating_ice", "mask_icefree_land"):
            put(n, getattr(ice, n, None), np.int32)
        if bc is not None:
            put("BC_prescr_mask_b", bc[0], np.int32)
            put("BC_prescr_u_b", bc[1], np.float64)
            put("BC_prescr_v_b", bc[2], np.float64)
        return s, keep

    def _state_struct(self, outputs=True):
        s = capi.ufe_diva_state()
        names = ["u_vav_b", "v_vav_b", "tau_bx_b", "tau_by_b", "eta_3D_b", "u_base_b", "v_base_b"]
        if outputs:
            names += ["u_3D_b", "v_3D_b", "du_dx_a", "du_dy_a", "dv_dx_a", "dv_dy_a", "du_dz_3D_a",
                      "dv_dz_3D_a", "eta_3D_a", "basal_friction_coefficient_a"]
        for n in names:
            setattr(s, n, vp(getattr(self, n)))
        return s

    # ---- L2
    def solve_DIVA(self, ice, BC_prescr_mask_b=None, BC_prescr_u_b=None, BC_prescr_v_b=None,
                   outputs=True) -> SolveInfo:
        """solve_DIVA(mesh, ice, bed_roughness, DIVA, n_visc_its, n_Axb_its [, BC_prescr_*])."""
        given = [a is not None for a in (BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b)]
        if any(given) and not all(given):
            raise UfeError(1, "need to provide prescribed u,v fields and mask!")
        bc = (BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b) if all(given) else None
        s_ice, keep = self._ice_struct(ice, bc)
        st = self._state_struct(outputs)
        info = capi.ufe_solve_info()
        check(capi.lib().ufe_diva_solve(self._h, ct.byref(s_ice), ct.byref(st), ct.byref(info)))
        return _info(info)

    def solve_SSA(self, ice, BC_prescr_mask_b=None, BC_prescr_u_b=None, BC_prescr_v_b=None) -> SolveInfo:
        given = [a is not None for a in (BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b)]
        if any(given) and not all(given):
            raise UfeError(1, "need to provide prescribed u,v fields and mask!")
        bc = (BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b) if all(given) else None
        s_ice, keep = self._ice_struct(ice, bc)
        st = capi.ufe_ssa_state(vp(self.u_b), vp(self.v_b), vp(self.basal_friction_coefficient_a))
        info = capi.ufe_solve_info()
        check(capi.lib().ufe_ssa_solve(self._h, ct.byref(s_ice), ct.byref(st), ct.byref(info)))
        return _info(info)

    # ---- device-resident variants
    def upload(self, ice=None, state=True):
        s_ice, keep = self._ice_struct(ice) if ice is not None else (None, None)
        st = self._state_struct(False) if state else None
        check(capi.lib().ufe_diva_upload(self._h, ct.byref(s_ice) if s_ice else None,
                                         ct.byref(st) if st else None))

    def solve_DIVA_resident(self) -> SolveInfo:
        info = capi.ufe_solve_info()
        check(capi.lib().ufe_diva_solve_resident(self._h, ct.byref(info)))
        return _info(info)

    def reset_state_resident(self):
        """'zero' initial velocities (DIVA_main.f90:60-68) on the device-resident state."""
        check(capi.lib().ufe_diva_reset_state(self._h))

    def download(self, outputs=True):
        st = self._state_struct(outputs)
        check(capi.lib().ufe_diva_download(self._h, ct.byref(st)))

    # ---- the step right after the solve
    def calc_secondary_velocities(self) -> dict:
        """set_ice_velocities_to_DIVA_results + calc_secondary_velocities
        (conservation_of_momentum_main.f90:470-510, 176-245) on the resident result of the last
        solve_DIVA.  Returns the ice%... fields by name."""
        nT, nV, nz = self.mesh.nTri, self.mesh.nV, self.mesh.nz
        out, st = {}, capi.ufe_secondary_velocities()
        for n in capi.SECONDARY_FIELDS:
            shape = (nV, nz) if n in ("u_3D", "v_3D") else ((nT,) if n.endswith("_b") else (nV,))
            out[n] = np.zeros(shape, order="F")
            setattr(st, n, vp(out[n]))
        check(capi.lib().ufe_calc_secondary_velocities(self._h, ct.byref(st)))
        return out

    # ---- ice-thickness rates of change
    def set_mesh_edges(self, edges):
        """Upload mesh%VE, ETri, A, Cw, D_x, D_y, D (``mesh_types.MeshEdges``) once per mesh."""
        e = capi.ufe_mesh_edges()
        e.nE = edges.nE
        keep = []
        for n, dt in (("VE", np.int32), ("ETri", np.int32), ("A", np.float64), ("Cw", np.float64),
                      ("D_x", np.float64), ("D_y", np.float64), ("D", np.float64)):
            a = np.asfortranarray(getattr(edges, n), dtype=dt)
            keep.append(a)
            setattr(e, n, vp(a))
        check(capi.lib().ufe_mesh_set_edges(self._h, ct.byref(e)))

    def _thickness_structs(self, fields: dict):
        C = self.C
        cfg = capi.ufe_thickness_config()
        cfg.dHi_semiimplicit_fs, cfg.dHi_PETSc_rtol, cfg.dHi_PETSc_abstol = C.dHi_semiimplicit_fs, C.dHi_PETSc_rtol, C.dHi_PETSc_abstol
        for i, side in enumerate(("north", "east", "south", "west")):
            cfg.BC_H[i] = _code(BC_H_CODES, getattr(C, "BC_H_" + side), "BC_H")
        cfg.dt_ice_max, cfg.dt_ice_min, cfg.Hi_min = C.dt_ice_max, C.dt_ice_min, C.Hi_min
        cfg.krylov_method = _code(KRYLOV_METHODS, C.b200_krylov_method, "b200_krylov_method")
        cfg.krylov_maxits = C.b200_krylov_maxits
        f = capi.ufe_thickness_fields()
        keep, out = [], {}
        for n in capi.THICKNESS_IN:
            a = fields.get(n)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                keep.append(a)
                setattr(f, n, vp(a))
        a = np.ascontiguousarray(fields["mask_noice"], dtype=np.int32)
        keep.append(a)
        f.mask_noice = vp(a)
        if fields.get("BC_prescr_mask") is not None:
            a = np.ascontiguousarray(fields["BC_prescr_mask"], dtype=np.int32)
            keep.append(a)
            f.BC_prescr_mask = vp(a)
        if fields.get("BC_prescr_Hi") is not None:
            a = np.ascontiguousarray(fields["BC_prescr_Hi"], dtype=np.float64)
            keep.append(a)
            f.BC_prescr_Hi = vp(a)
        for n in capi.THICKNESS_OUT:
            out[n] = np.zeros(self.mesh.nV)
            setattr(f, n, vp(out[n]))
        return cfg, f, keep, out

    def calc_dHi_dt(self, fields: dict, dt: float) -> dict:
        """calc_dHi_dt (conservation_of_mass_main.f90:22): dispatch on C%choice_ice_integration_method, clip negative
        thicknesses, final dHi_dt and AMB.  Returns the out-arguments + ``dt``, ``n_Axb_its``, ``flags``."""
        cfg, f, keep, out = self._thickness_structs(fields)
        method = _code(ICE_INTEGRATION_CODES, self.C.choice_ice_integration_method, "choice_ice_integration_method")
        d, its, fl = ct.c_double(dt), ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_calc_dHi_dt(self._h, ct.byref(cfg), method, ct.byref(f), ct.byref(d), ct.byref(its), ct.byref(fl)))
        out["dt"], out["n_Axb_its"], out["flags"] = d.value, its.value, fl.value
        return out

    def calc_dHi_dt_explicit(self, fields: dict, dt: float) -> dict:
        """calc_dHi_dt_explicit(mesh, Hi, Hb, SL, u_vav_b, v_vav_b, SMB, BMB, LMB, AMB, fraction_margin, mask_noice,
        dt, dHi_dt, Hi_tplusdt, divQ, dHi_dt_target [, BC_prescr_mask, BC_prescr_Hi])
        (conservation_of_mass_explicit.f90:23).  ``fields`` holds the in-arguments by name; u_vav_b / v_vav_b may be
        left out to use the resident result of the last velocity solve.  Returns the out-arguments + ``dt``."""
        cfg, f, keep, out = self._thickness_structs(fields)
        d = ct.c_double(dt)
        check(capi.lib().ufe_calc_dHi_dt_explicit(self._h, ct.byref(cfg), ct.byref(f), ct.byref(d)))
        out["dt"] = d.value
        return out

    def calc_dHi_dt_semiimplicit(self, fields: dict, dt: float) -> dict:
        """calc_dHi_dt_semiimplicit (conservation_of_mass_semiimplicit.f90:24).  Returns the out-arguments,
        ``n_Axb_its`` and ``flags``."""
        cfg, f, keep, out = self._thickness_structs(fields)
        its, fl = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_calc_dHi_dt_semiimplicit(self._h, ct.byref(cfg), ct.byref(f), ct.c_double(dt),
                                                      ct.byref(its), ct.byref(fl)))
        out["n_Axb_its"], out["flags"] = its.value, fl.value
        return out

    def get_thickness_matrix(self, which="M_divQ"):
        """M_divQ of the most recent call, or (AA, bb) of the most recent semi-implicit call."""
        w = {"M_divQ": 0, "AA": 1}[which]
        m_loc, nnz = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_get_thickness_csr(self._h, w, ct.byref(m_loc), ct.byref(nnz), None, None, None, None))
        ptr = np.zeros(m_loc.value + 1, dtype=np.int32)
        ind, val, bb = np.zeros(nnz.value, dtype=np.int32), np.zeros(nnz.value), np.zeros(m_loc.value)
        check(capi.lib().ufe_get_thickness_csr(self._h, w, ct.byref(m_loc), ct.byref(nnz), vp(ptr), vp(ind), vp(val), vp(bb)))
        A = CSRMatrix(self.mesh.nV, self.mesh.nV, 1, self.mesh.nV, ptr, ind, val)
        return A if w == 0 else (A, bb)

    def calc_vertical_velocities(self, ice: dict) -> np.ndarray:
        """calc_vertical_velocities(mesh, ice, BMB) (vertical_velocities.f90:18): ``ice`` holds Hi, Hib, dHb_dt, dHi_dt,
        BMB, mask_grounded_ice, mask_floating_ice, dzeta_dx_ak, dzeta_dy_ak, dzeta_dz_ak; the horizontal velocities are
        the resident results of solve_DIVA + calc_secondary_velocities.  Returns ice%w_3D (nV,nz)."""
        s = capi.ufe_vertical_velocity_inputs()
        keep = []
        for n in capi.VERTICAL_IN:
            a = np.asfortranarray(ice[n], dtype=np.int32 if n.startswith("mask") else np.float64)
            keep.append(a)
            setattr(s, n, vp(a))
        w = np.zeros((self.mesh.nV, self.mesh.nz), order="F")
        check(capi.lib().ufe_calc_vertical_velocities(self._h, ct.byref(s), vp(w)))
        return w

    def get_operator_a_a(self, which: str) -> CSRMatrix:
        """mesh%M_ddx_a_a / M_ddy_a_a as built on the device (needs set_mesh_edges)."""
        w = {"ddx": 0, "ddy": 1}[which]
        m_loc, nnz = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_mesh_get_operator_a_a(self._h, w, ct.byref(m_loc), ct.byref(nnz), None, None, None))
        ptr = np.zeros(m_loc.value + 1, dtype=np.int32)
        ind, val = np.zeros(nnz.value, dtype=np.int32), np.zeros(nnz.value)
        check(capi.lib().ufe_mesh_get_operator_a_a(self._h, w, ct.byref(m_loc), ct.byref(nnz), vp(ptr), vp(ind), vp(val)))
        return CSRMatrix(self.mesh.nV, self.mesh.nV, 1, self.mesh.nV, ptr, ind, val)

    def thickness_timing(self) -> dict:
        """Device-time split (ms) of the most recent thickness call + algorithmic bytes of k_thk_divq."""
        ms = (ct.c_double * 6)()
        by = ct.c_double()
        check(capi.lib().ufe_get_thickness_timing(self._h, ms, ct.byref(by)))
        keys = ("ms_h2d", "ms_divq", "ms_explicit_and_system", "ms_krylov", "ms_finish", "ms_d2h")
        return dict(zip(keys, list(ms)), divq_algorithmic_bytes=by.value)

    # ---- L1
    def solve_SSA_DIVA_linearised(self, u_b, v_b, N_b, dN_dx_b, dN_dy_b, basal_friction_coefficient_b,
                                  tau_dx_b, tau_dy_b, PETSc_rtol, PETSc_abstol, BC_prescr_mask_b=None,
                                  BC_prescr_u_b=None, BC_prescr_v_b=None):
        """Returns (u_b, v_b, u_b_prev, v_b_prev, n_Axb_its)."""
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        u, v = np.array(u_b, dtype=np.float64), np.array(v_b, dtype=np.float64)
        up, vpv = np.zeros_like(u), np.zeros_like(v)
        args = [c(a) for a in (N_b, dN_dx_b, dN_dy_b, basal_friction_coefficient_b, tau_dx_b, tau_dy_b)]
        its = ct.c_int32()
        m = None if BC_prescr_mask_b is None else np.ascontiguousarray(BC_prescr_mask_b, dtype=np.int32)
        bu = None if BC_prescr_u_b is None else c(BC_prescr_u_b)
        bv = None if BC_prescr_v_b is None else c(BC_prescr_v_b)
        check(capi.lib().ufe_ssa_diva_linearised(self._h, vp(u), vp(v), *[vp(a) for a in args], vp(up), vp(vpv),
                                                 ct.c_double(PETSc_rtol), ct.c_double(PETSc_abstol),
                                                 ct.byref(its), vp(m), vp(bu), vp(bv)))
        return u, v, up, vpv, its.value

    # ---- L0 on the ranks of this handle
    def solve_matrix_equation_CSR(self, A: CSRMatrix, bb, xx, rtol, abstol):
        """solve_matrix_equation_CSR_PETSc( A_CSR, bb, xx, rtol, abstol, n_Axb_its) as the reference calls it
        (petsc_basic.f90:32-64): every rank passes ITS rows of A and its slices of bb / xx; method, preconditioner,
        maxits and the initial-guess switch come from the solver's config.  Returns (xx, n_Axb_its, flags, pc_used)."""
        keep = []
        s = capi.csr_struct(A.m, A.n, A.i1, A.i2, A.ptr, A.ind, A.val, keep)
        b = np.ascontiguousarray(bb, dtype=np.float64)
        x = np.array(xx, dtype=np.float64, copy=True)
        if b.size != A.i2 - A.i1 + 1 or x.size != b.size:
            raise UfeError(1, "matrix and vector sub-sizes dont match!")
        its, fl = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_solve_matrix_equation_CSR(self._h, ct.byref(s), vp(b), vp(x), ct.c_double(rtol),
                                                       ct.c_double(abstol), ct.byref(its), ct.byref(fl)))
        return x, its.value, fl.value, capi.lib().ufe_last_l0_preconditioner(self._h)

    # ---- operators (mesh%M_*), as held on the device
    def get_operator(self, family: str, which: str) -> CSRMatrix:
        fid, names = FAMILIES[family]
        w = names.index(which)
        m_loc, nnz = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_mesh_get_operator(self._h, fid, w, ct.byref(m_loc), ct.byref(nnz), None, None, None))
        ptr = np.zeros(m_loc.value + 1, dtype=np.int32)
        ind = np.zeros(max(nnz.value, 1), dtype=np.int32)
        val = np.zeros(max(nnz.value, 1))
        check(capi.lib().ufe_mesh_get_operator(self._h, fid, w, ct.byref(m_loc), ct.byref(nnz), vp(ptr), vp(ind), vp(val)))
        vi1, vi2, ti1, ti2 = self.ownership()
        i1, i2 = (vi1, vi2) if family == "b_a" else (ti1, ti2)
        m = self.mesh.nV if family == "b_a" else self.mesh.nTri
        n = self.mesh.nV if family == "a_b" else self.mesh.nTri
        return CSRMatrix(m, n, i1, i2, ptr, ind[:nnz.value], val[:nnz.value])

    def _apply(self, family, which, d):
        fid, names = FAMILIES[family]
        x = np.asfortranarray(d, dtype=np.float64)
        nl = 1 if x.ndim == 1 else x.shape[1]
        m = self.mesh.nV if family == "b_a" else self.mesh.nTri
        y = np.zeros((m,) if x.ndim == 1 else (m, nl), order="F")
        check(capi.lib().ufe_mesh_apply_operator(self._h, fid, names.index(which), vp(x), vp(y), nl))
        return y

    # mesh_disc_apply_operators.f90:121-431
    def map_a_b_2D(self, d): return self._apply("a_b", "map", d)
    def map_a_b_3D(self, d): return self._apply("a_b", "map", d)
    def ddx_a_b_2D(self, d): return self._apply("a_b", "ddx", d)
    def ddy_a_b_2D(self, d): return self._apply("a_b", "ddy", d)
    def map_b_a_2D(self, d): return self._apply("b_a", "map", d)
    def map_b_a_3D(self, d): return self._apply("b_a", "map", d)
    def ddx_b_a_2D(self, d): return self._apply("b_a", "ddx", d)
    def ddy_b_a_2D(self, d): return self._apply("b_a", "ddy", d)

    def get_stiffness_matrix(self):
        """(A_CSR, bb) of the most recent linearised solve, reference layout."""
        m_loc, nnz = ct.c_int32(), ct.c_int32()
        check(capi.lib().ufe_get_stiffness_csr(self._h, ct.byref(m_loc), ct.byref(nnz), None, None, None, None))
        ptr = np.zeros(m_loc.value + 1, dtype=np.int32)
        ind = np.zeros(max(nnz.value, 1), dtype=np.int32)
        val = np.zeros(max(nnz.value, 1))
        bb = np.zeros(max(m_loc.value, 1))
        check(capi.lib().ufe_get_stiffness_csr(self._h, ct.byref(m_loc), ct.byref(nnz), vp(ptr), vp(ind), vp(val), vp(bb)))
        _, _, ti1, ti2 = self.ownership()
        N = 2 * self.mesh.nTri
        return CSRMatrix(N, N, 2 * ti1 - 1, 2 * ti2, ptr, ind[:nnz.value], val[:nnz.value]), bb[:m_loc.value]

    def bench_spmv(self, reps=50, flush_l2=False):
        ms, by = ct.c_double(), ct.c_double()
        check(capi.lib().ufe_bench_spmv(self._h, reps, int(flush_l2), ct.byref(ms), ct.byref(by)))
        return ms.value, by.value


def remap_DIVA_solver(solver_old: DIVASolver, mesh_new: Mesh, map_from_mesh_to_mesh, comm=None) -> DIVASolver:
    """remap_DIVA_solver(mesh_old, mesh_new, DIVA) (DIVA_main.f90:264-373): the seven fields that are re-used by the
    next solve (u_vav_b, v_vav_b, tau_bx_b, tau_by_b, eta_3D_b, u_3D_b, v_3D_b) go b -> a on the old mesh (device),
    a(old) -> a(new) through ``map_from_mesh_to_mesh(d_a_old) -> d_a_new`` (the reference's
    map_from_mesh_to_mesh_with_reallocation_2D/3D, '2nd_order_conservative': the caller's remapping subsystem, out of
    scope here), and a -> b on the new mesh (device); everything else is reallocated (zero).  Returns the solver for
    ``mesh_new``; the old one is closed."""
    names2, names3 = ("u_vav_b", "v_vav_b", "tau_bx_b", "tau_by_b"), ("eta_3D_b", "u_3D_b", "v_3D_b")
    on_a = {n: solver_old.map_b_a_2D(getattr(solver_old, n)) for n in names2}
    on_a.update({n: solver_old.map_b_a_3D(getattr(solver_old, n)) for n in names3})
    C = solver_old.C
    solver_old.close()
    new = DIVASolver(mesh_new, C, comm)
    for n, d in on_a.items():
        d_new = np.asfortranarray(map_from_mesh_to_mesh(d), dtype=np.float64)
        want = (mesh_new.nV,) if n in names2 else (mesh_new.nV, mesh_new.nz)
        if d_new.shape != want:
            raise UfeError(1, f"remap_DIVA_solver: remapped {n} has shape {d_new.shape}, expected {want}")
        out = new.map_a_b_2D(d_new) if n in names2 else new.map_a_b_3D(d_new)
        getattr(new, n)[...] = out
    return new


def initialise_DIVA_solver(mesh: Mesh, C: Config, comm=None, operators=None) -> DIVASolver:
    """initialise_DIVA_solver (DIVA_main.f90:37-86): allocate + 'zero' initial velocities."""
    if C.choice_initial_velocity != "zero":
        raise UfeError(1, f'unknown choice_initial_velocity "{C.choice_initial_velocity}"! '
                          "(read_from_file is NetCDF I/O, out of scope: set the state arrays instead)")
    return DIVASolver(mesh, C, comm, operators)
