"""ctypes binding of include/ufe_diva.h (libufe_diva.so, built in-tree by csrc/Makefile).

There is no CPU fallback: if the shared library is missing or no sm_100 device is usable,
every compute entry point raises ``UfeError``.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UFE_LIB_PATH") or os.path.join(_HERE, "libufe_diva.so")   # UFE_LIB_PATH: A/B runs of two builds

c_i32, c_f64 = ct.c_int32, ct.c_double
P_i32, P_f64 = ct.POINTER(ct.c_int32), ct.POINTER(ct.c_double)


class UfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ufe error {code}: {msg}")
        self.code = code


class ufe_csr(ct.Structure):
    _fields_ = [("m", c_i32), ("n", c_i32), ("m_loc", c_i32), ("n_loc", c_i32), ("i1", c_i32),
                ("i2", c_i32), ("j1", c_i32), ("j2", c_i32), ("nnz", c_i32),
                ("ptr", ct.c_void_p), ("ind", ct.c_void_p), ("val", ct.c_void_p)]


class ufe_mesh(ct.Structure):
    _fields_ = [("nV", c_i32), ("nTri", c_i32), ("nC_mem", c_i32), ("nz", c_i32),
                ("xmin", c_f64), ("xmax", c_f64), ("ymin", c_f64), ("ymax", c_f64),
                ("V", ct.c_void_p), ("Tri", ct.c_void_p), ("TriC", ct.c_void_p), ("C", ct.c_void_p),
                ("nC", ct.c_void_p), ("iTri", ct.c_void_p), ("niTri", ct.c_void_p),
                ("VBI", ct.c_void_p), ("TriBI", ct.c_void_p), ("TriGC", ct.c_void_p),
                ("zeta", ct.c_void_p),
                ("M_a_b", ct.POINTER(ufe_csr) * 3), ("M_b_a", ct.POINTER(ufe_csr) * 3),
                ("M2_b_b", ct.POINTER(ufe_csr) * 5)]


class ufe_config(ct.Structure):
    _fields_ = [("do_include_SSADIVA_crossterms", c_i32), ("visc_it_norm_dUV_tol", c_f64),
                ("visc_it_nit", c_i32), ("visc_it_relax", c_f64), ("visc_eff_min", c_f64),
                ("vel_max", c_f64), ("stress_balance_PETSc_rtol", c_f64),
                ("stress_balance_PETSc_abstol", c_f64), ("BC_u", c_i32 * 4), ("BC_v", c_i32 * 4),
                ("choice_sliding_law", c_i32), ("choice_idealised_sliding_law", c_i32),
                ("slid_Weertman_m", c_f64), ("slid_Budd_q_plastic", c_f64),
                ("slid_Budd_u_threshold", c_f64), ("slid_ZI_p", c_f64), ("slid_ZI_ut", c_f64),
                ("do_GL_subgrid_friction", c_i32), ("do_subgrid_friction_on_A_grid", c_i32),
                ("subgrid_friction_exponent_on_B_grid", c_f64), ("slid_beta_max", c_f64),
                ("slid_delta_v", c_f64), ("Hi_min", c_f64), ("Glens_flow_law_exponent", c_f64),
                ("Glens_flow_law_epsilon_sq_0", c_f64), ("choice_ice_rheology_Glen", c_i32),
                ("uniform_Glens_flow_factor", c_f64), ("choice_enhancement_factor_transition", c_i32),
                ("m_enh_sheet", c_f64), ("m_enh_shelf", c_f64),
                ("refgeo_idealised_SSA_icestream_Hi", c_f64), ("refgeo_idealised_SSA_icestream_dhdx", c_f64),
                ("refgeo_idealised_SSA_icestream_L", c_f64), ("refgeo_idealised_SSA_icestream_m", c_f64),
                ("refgeo_idealised_ISMIP_HOM_L", c_f64),
                ("krylov_method", c_i32), ("krylov_pc", c_i32), ("krylov_maxits", c_i32),
                ("krylov_guess_nonzero", c_i32), ("krylov_pc_lag", c_i32), ("krylov_pc_strip_only", c_i32)]


class ufe_ice_inputs(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in (
        "Hi", "Hs", "Hib", "SL", "fraction_gr", "fraction_gr_b", "effective_pressure",
        "mask_grounded_ice", "mask_floating_ice", "mask_icefree_land", "Ti",
        "till_friction_angle", "alpha_sq", "beta_sq", "BC_prescr_mask_b", "BC_prescr_u_b",
        "BC_prescr_v_b")]


class ufe_diva_state(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in (
        "u_vav_b", "v_vav_b", "tau_bx_b", "tau_by_b", "eta_3D_b", "u_base_b", "v_base_b",
        "u_3D_b", "v_3D_b", "du_dx_a", "du_dy_a", "dv_dx_a", "dv_dy_a", "du_dz_3D_a",
        "dv_dz_3D_a", "eta_3D_a", "basal_friction_coefficient_a")]


class ufe_ssa_state(ct.Structure):
    _fields_ = [("u_b", ct.c_void_p), ("v_b", ct.c_void_p), ("basal_friction_coefficient_a", ct.c_void_p)]


class ufe_solve_info(ct.Structure):
    _fields_ = [("n_visc_its", c_i32), ("n_Axb_its", c_i32), ("flags", c_i32), ("L2_uv", c_f64),
                ("visc_it_relax_applied", c_f64), ("Glens_flow_law_epsilon_sq_0_applied", c_f64),
                ("ms_total", c_f64), ("ms_closures", c_f64), ("ms_assembly", c_f64),
                ("ms_krylov", c_f64), ("ms_h2d", c_f64), ("ms_d2h", c_f64), ("gpu_launches", ct.c_int64),
                ("krylov_pc_used", c_i32), ("reserved", c_i32)]


SECONDARY_FIELDS = ("u_surf_b", "v_surf_b", "uabs_surf_b", "u_base_b", "v_base_b", "uabs_base_b", "u_vav_b", "v_vav_b",
                    "uabs_vav_b", "u_3D", "v_3D", "u_surf", "v_surf", "uabs_surf", "u_base", "v_base", "uabs_base", "u_vav",
                    "v_vav", "uabs_vav", "R_shear")


class ufe_secondary_velocities(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in SECONDARY_FIELDS]


class ufe_mesh_edges(ct.Structure):
    _fields_ = [("nE", c_i32), ("VE", ct.c_void_p), ("ETri", ct.c_void_p), ("A", ct.c_void_p),
                ("Cw", ct.c_void_p), ("D_x", ct.c_void_p), ("D_y", ct.c_void_p), ("D", ct.c_void_p)]


class ufe_thickness_config(ct.Structure):
    _fields_ = [("dHi_semiimplicit_fs", c_f64), ("dHi_PETSc_rtol", c_f64), ("dHi_PETSc_abstol", c_f64),
                ("BC_H", c_i32 * 4), ("dt_ice_max", c_f64), ("dt_ice_min", c_f64), ("Hi_min", c_f64),
                ("krylov_method", c_i32), ("krylov_maxits", c_i32)]


THICKNESS_IN = ("Hi", "Hb", "SL", "u_vav_b", "v_vav_b", "SMB", "BMB", "LMB", "fraction_margin", "dHi_dt_target")
THICKNESS_OUT = ("AMB", "dHi_dt", "Hi_tplusdt", "divQ")


class ufe_thickness_fields(ct.Structure):
    _fields_ = ([(n, ct.c_void_p) for n in THICKNESS_IN] +
                [("mask_noice", ct.c_void_p), ("BC_prescr_mask", ct.c_void_p), ("BC_prescr_Hi", ct.c_void_p)] +
                [(n, ct.c_void_p) for n in THICKNESS_OUT])


VERTICAL_IN = ("Hi", "Hib", "dHb_dt", "dHi_dt", "BMB", "mask_grounded_ice", "mask_floating_ice",
               "dzeta_dx_ak", "dzeta_dy_ak", "dzeta_dz_ak")


class ufe_vertical_velocity_inputs(ct.Structure):
    _fields_ = [(n, ct.c_void_p) for n in VERTICAL_IN]


class ufe_comm(ct.Structure):
    _fields_ = [("rank", c_i32), ("nranks", c_i32), ("device", c_i32), ("nccl_unique_id", ct.c_char_p)]


EXPORTS = [
    "ufe_last_error_string", "ufe_comm_get_unique_id", "ufe_version", "ufe_sizeof_solve_info", "ufe_partition_list",
    "ufe_krylov_solve", "ufe_spmv", "ufe_diva_create", "ufe_diva_destroy", "ufe_diva_set_config",
    "ufe_diva_solve", "ufe_ssa_solve", "ufe_diva_upload", "ufe_diva_solve_resident",
    "ufe_diva_download", "ufe_diva_reset_state", "ufe_calc_secondary_velocities", "ufe_ssa_diva_linearised", "ufe_mesh_get_operator",
    "ufe_mesh_apply_operator", "ufe_get_stiffness_csr", "ufe_bench_spmv", "ufe_get_ownership",
    "ufe_mesh_set_edges", "ufe_calc_dHi_dt", "ufe_calc_dHi_dt_explicit", "ufe_calc_dHi_dt_semiimplicit", "ufe_get_thickness_csr",
    "ufe_get_thickness_timing", "ufe_calc_vertical_velocities", "ufe_mesh_get_operator_a_a",
    "ufe_nd_analyse", "ufe_nd_tree_info", "ufe_nd_tree_node", "ufe_nd_tree_entry_map", "ufe_nd_tree_owners", "ufe_nd_tree_free",
    "ufe_solve_matrix_equation_CSR", "ufe_last_l0_preconditioner",
    "ufe_nd_solver_create", "ufe_nd_solver_factor", "ufe_nd_solver_solve", "ufe_nd_solver_info", "ufe_nd_solver_free",
]

_lib = None


def _preload_nccl():
    """libufe_diva.so needs libnccl.so.2.  In a process that also imports torch, the NCCL
    bundled with torch (nvidia/nccl/lib) must be the one that is loaded, so load it first
    when it exists; otherwise the system libnccl.so.2 is used."""
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            ct.CDLL(cand, mode=ct.RTLD_GLOBAL)
            return cand
    return None


def lib():
    """Load libufe_diva.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UfeError(2, f"{LIB_PATH} not found: build it with `make -C ufemism2.0_b200/csrc` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        _preload_nccl()
        _lib = ct.CDLL(LIB_PATH)
        _lib.ufe_last_error_string.restype = ct.c_char_p
        for name in EXPORTS:
            getattr(_lib, name)
    return _lib


def check(rc):
    if rc != 0:
        raise UfeError(rc, lib().ufe_last_error_string().decode())


def vp(a):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


def csr_struct(m, n, i1, i2, ptr, ind, val, keep):
    """Build a ufe_csr from numpy arrays (kept alive through ``keep``)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    keep.extend([ptr, ind, val])
    return ufe_csr(m, n, i2 - i1 + 1, n, i1, i2, 1, n, int(ind.size), vp(ptr), vp(ind), vp(val))
