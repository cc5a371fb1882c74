"""Config keys read by the DIVA/SSA velocity-solve path, under the reference's names.

Defaults and names follow ``src/UFEMISM/basic/model_configuration.f90`` (lines cited per
key).  ``Config.from_namelist`` reads a reference ``&CONFIG ... /`` file (keys carry the
``_config`` suffix there, copied to ``C%<name>`` at ``model_configuration.f90:3248``);
keys the path does not read are ignored.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, fields


@dataclass
class Config:
    # zeta grid (:266-268)
    choice_zeta_grid: str = "regular"
    nz: int = 12
    zeta_irregular_log_R: float = 10.0
    # general (:274-276)
    choice_stress_balance_approximation: str = "DIVA"
    do_include_SSADIVA_crossterms: bool = True
    # initialisation (:290-293)
    choice_initial_velocity: str = "zero"
    # viscosity iteration + linear solver (:307-313)
    visc_it_norm_dUV_tol: float = 5e-5
    visc_it_nit: int = 50
    visc_it_relax: float = 0.2
    visc_eff_min: float = 1e4
    vel_max: float = 5000.0
    stress_balance_PETSc_rtol: float = 1e-7
    stress_balance_PETSc_abstol: float = 1e-5
    # boundary conditions (:316-323)
    BC_u_west: str = "infinite"
    BC_u_east: str = "infinite"
    BC_u_south: str = "infinite"
    BC_u_north: str = "infinite"
    BC_v_west: str = "infinite"
    BC_v_east: str = "infinite"
    BC_v_south: str = "infinite"
    BC_v_north: str = "infinite"
    # sliding (:329-349)
    choice_sliding_law: str = "Zoet-Iverson"
    choice_idealised_sliding_law: str = ""
    slid_Weertman_m: float = 3.0
    slid_Budd_q_plastic: float = 0.3
    slid_Budd_u_threshold: float = 100.0
    slid_ZI_p: float = 5.0
    slid_ZI_ut: float = 200.0
    do_GL_subgrid_friction: bool = True
    do_subgrid_friction_on_A_grid: bool = False
    subgrid_friction_exponent_on_B_grid: float = 2.0
    slid_beta_max: float = 1e20
    slid_delta_v: float = 1e-3
    Hi_min: float = 0.0                                   # :437
    # ice thickness integration (:355-365, :391-392) -- read by calc_dHi_dt_explicit / _semiimplicit
    choice_ice_integration_method: str = "semi-implicit"
    dHi_semiimplicit_fs: float = 1.5
    dHi_PETSc_rtol: float = 1e-8
    dHi_PETSc_abstol: float = 1e-6
    BC_H_west: str = "zero"
    BC_H_east: str = "zero"
    BC_H_south: str = "zero"
    BC_H_north: str = "zero"
    dt_ice_max: float = 10.0
    dt_ice_min: float = 0.1
    # rheology (:590-601)
    choice_flow_law: str = "Glen"
    Glens_flow_law_exponent: float = 3.0
    Glens_flow_law_epsilon_sq_0: float = 1e-8
    choice_ice_rheology_Glen: str = "Huybrechts1992"
    uniform_Glens_flow_factor: float = 1e-16
    choice_enhancement_factor_transition: str = "separate"
    m_enh_sheet: float = 1.0
    m_enh_shelf: float = 1.0
    # idealised geometry parameters read by BCs / idealised sliding (:190-195)
    refgeo_idealised_SSA_icestream_Hi: float = -1.0
    refgeo_idealised_SSA_icestream_dhdx: float = 1.0
    refgeo_idealised_SSA_icestream_L: float = 0.0
    refgeo_idealised_SSA_icestream_m: float = 0.0
    refgeo_idealised_ISMIP_HOM_L: float = 0.0

    # ---- extensions of the B200 build (not reference keys; defaults keep the
    # reference's behaviour).  The PETSc options database is the reference's only other
    # knob and is unused there.
    b200_krylov_method: str = "bicgstab"      # 'bicgstab' | 'gmres'
    # 'jacobi' | 'bjacobi2' (2x2 u-v blocks) | 'bjacobi_lu' (one block per GPU, solved exactly by block
    # cyclic reduction; needs a banded = x-sorted, narrow mesh) | 'auto' (bjacobi_lu when it fits, else nd_lu on one GPU when it fits, else bjacobi2)
    # | 'nd_lu' (exact multifrontal nested-dissection factorisation of the whole matrix: wide meshes, one GPU)
    b200_krylov_pc: str = "auto"
    b200_krylov_pc_strip_only: bool = False   # several ranks: True = one strip block per rank even when the replicated exact solve fits
    b200_krylov_pc_lag: int = 0               # bjacobi_lu: >0 reuses a factorisation until a solve needs more Krylov its than this (0: factorise every Picard it)
    b200_krylov_maxits: int = 10000           # PETSc default maxits
    b200_krylov_guess_nonzero: bool = False   # False = KSP default (zero initial guess)

    @classmethod
    def from_namelist(cls, path: str, region: str = "ANT") -> "Config":
        """Parse a reference config file (Fortran NAMELIST ``&CONFIG``)."""
        names = {f.name: f for f in fields(cls)}
        cfg = cls()
        with open(path, "r", encoding="utf-8", errors="replace") as fh:
            text = fh.read()
        for line in text.splitlines():
            line = line.split("!")[0].strip()
            m = re.match(r"^([A-Za-z0-9_+\-]+)_config\s*=\s*(.+?)\s*,?$", line)
            if not m:
                continue
            key, raw = m.group(1), m.group(2).strip()
            if key == f"choice_initial_velocity_{region}":
                key = "choice_initial_velocity"
            if key not in names:
                continue
            typ = names[key].type
            if typ in ("bool", bool):
                val = raw.strip(".").upper().startswith("T")
            elif typ in ("int", int):
                val = int(raw)
            elif typ in ("float", float):
                val = float(re.sub(r"_dp$", "", raw).replace("D", "E").replace("d", "e"))
            else:
                val = raw.strip("'\"").strip()
            setattr(cfg, key, val)
        return cfg


BC_CODES = {"infinite": 1, "zero": 2, "periodic_ISMIP-HOM": 3, "infinite_SSA_icestream": 4}
SLIDING_CODES = {"no_sliding": 0, "idealised": 1, "Weertman": 2, "Coulomb": 3, "Budd": 4,
                 "Tsai2015": 5, "Schoof2005": 6, "Zoet-Iverson": 7}
IDEALISED_SLIDING_CODES = {"": 0, "SSA_icestream": 1, "ISMIP-HOM_C": 2, "ISMIP-HOM_D": 3,
                           "ISMIP-HOM_F": 5}
RHEOLOGY_CODES = {"uniform": 0, "Huybrechts1992": 1}
ENH_CODES = {"separate": 0, "interp": 1}
BC_H_CODES = {"infinite": 1, "zero": 2}
ICE_INTEGRATION_CODES = {"none": 0, "explicit": 1, "semi-implicit": 2}
