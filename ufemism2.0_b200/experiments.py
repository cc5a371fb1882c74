"""Named workloads (BASELINE.json ``configs``): synthetic mesh + geometry + the config keys
of the reference's own ``.cfg`` for that experiment (values cited; the ``.cfg`` files
themselves are not read at run time -- /root/reference does not exist on the GPU box).

All meshes are Delaunay triangulations of the jittered lattice points (``delaunay=True``), like
the reference's own Delaunay-refined meshes.  The alternating-diagonal ("union jack") lattice
triangulation of round 1 has vertices with 4 and 8 neighbours alternately; on it the reference's
least-squares operators carry checkerboard-like near-null modes, the Picard iteration of the
reference algorithm does not converge on any case with an ice shelf or ice-free part (oracle and
GPU alike; profiles/r2_notes.md) and GMRES(30)/ILU(0) runs into its 10 000 cap.  On the Delaunay
meshes the same algorithm converges monotonically and the oracle's GMRES needs ~100-400
iterations per solve -- the reference's own operating point (BASELINE.md scoreboard).
"""
from __future__ import annotations

import numpy as np

from . import synthetic
from .config import Config


def _periodic(C: Config):
    for c in "uv":
        for s in ("west", "east", "south", "north"):
            setattr(C, f"BC_{c}_{s}", "periodic_ISMIP-HOM")
    return C


def ISMIP_HOM(exp: str, L: float, n: int = 41, seed=synthetic.SEED, delaunay=True):
    """automated_testing/integrated_tests/idealised/ISMIP-HOM/config_ISMIP_HOM_{A,C}_<L>_DIVA.cfg:
    domain [-L,L]^2, resolution L/20, periodic BCs, uniform A = 1e-16, eps0^2 = 1e-12,
    Picard tol 5e-7 / <= 5000 its / relax 0.4, Krylov 1e-6 / 1e-4."""
    mesh = synthetic.lattice_mesh(-L, L, -L, L, n, n, seed=seed, delaunay=delaunay)
    C = _periodic(Config(visc_it_norm_dUV_tol=5e-7, visc_it_nit=5000, visc_it_relax=0.4,
                         stress_balance_PETSc_rtol=1e-6, stress_balance_PETSc_abstol=1e-4,
                         choice_ice_rheology_Glen="uniform", uniform_Glens_flow_factor=1e-16,
                         Glens_flow_law_epsilon_sq_0=1e-12, refgeo_idealised_ISMIP_HOM_L=L))
    if exp == "A":
        C.choice_sliding_law = "no_sliding"
        ice = synthetic.geometry_ISMIP_HOM_A(mesh, L)
    elif exp == "C":
        C.choice_sliding_law = "idealised"
        C.choice_idealised_sliding_law = "ISMIP-HOM_C"
        ice = synthetic.geometry_ISMIP_HOM_C(mesh, L)
    else:
        raise ValueError(exp)
    return mesh, C, ice


def SSA_icestream(nx=41, ny=41, delaunay=True):
    """automated_testing/integrated_tests/idealised/SSA_icestream/config_04_4km.cfg:65-68,150-153,
    253-254,468: domain +-400 km, A = 1e-18, H = 2000, dh/dx = -3e-4, L = 150 km, m = 1,
    Krylov 1e-7 / 1e-5; BC_u west/east 'infinite_SSA_icestream', v 'zero' (SSA solve)."""
    mesh = synthetic.lattice_mesh(-400e3, 400e3, -400e3, 400e3, nx, ny, delaunay=delaunay)
    C = Config(choice_stress_balance_approximation="SSA", choice_sliding_law="idealised",
               choice_idealised_sliding_law="SSA_icestream", choice_ice_rheology_Glen="uniform",
               uniform_Glens_flow_factor=1e-18, refgeo_idealised_SSA_icestream_Hi=2000.0,
               refgeo_idealised_SSA_icestream_dhdx=-3e-4, refgeo_idealised_SSA_icestream_L=150e3,
               refgeo_idealised_SSA_icestream_m=1.0, visc_it_norm_dUV_tol=5e-8, visc_it_nit=5000,
               Glens_flow_law_epsilon_sq_0=1e-10,
               visc_it_relax=0.3, stress_balance_PETSc_rtol=1e-7, stress_balance_PETSc_abstol=1e-5,
               BC_u_west="infinite_SSA_icestream", BC_u_east="infinite_SSA_icestream",
               BC_u_south="zero", BC_u_north="zero", BC_v_west="zero", BC_v_east="zero",
               BC_v_south="zero", BC_v_north="zero")
    ice = synthetic.geometry_SSA_icestream(mesh, 2000.0, -3e-4)
    return mesh, C, ice


def MISMIP_8km(h=8e3, seed=synthetic.SEED, delaunay=True):
    """config-files/config_MISMIP_8km_spinup_for_scaling.cfg: 2000x2000 km, Zoet-Iverson
    phi = 10 deg, Martin2011 hydrology, A = 1e-16, Picard 5e-5 / 50 / relax 0.4,
    Krylov 1e-4 / 1e-3.  Geometry: MISMIP_mod bed (idealised_geometries.f90:214-241)
    with a synthetic dome (the reference spins up from 100 m of ice)."""
    n = int(round(2000e3 / h)) + 1
    mesh = synthetic.lattice_mesh(-1000e3, 1000e3, -1000e3, 1000e3, n, n, seed=seed, delaunay=delaunay)
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    r = np.sqrt(x * x + y * y)
    Hb = 150.0 - 400.0 * r / 750000.0
    s = np.clip(r / 850e3, 0.0, 1.0)
    Hi = np.where(r < 850e3, np.maximum(2800.0 * (1.0 - s ** (4.0 / 3.0)) ** (3.0 / 8.0), 120.0), 0.0)
    ice = synthetic._finish_inputs(mesh, Hi, Hb, np.zeros(mesh.nV), phi=10.0)
    C = Config(visc_it_norm_dUV_tol=5e-5, visc_it_nit=50, visc_it_relax=0.4,
               stress_balance_PETSc_rtol=1e-4, stress_balance_PETSc_abstol=1e-3,
               choice_sliding_law="Zoet-Iverson", choice_ice_rheology_Glen="uniform",
               uniform_Glens_flow_factor=1e-16, Hi_min=0.1,
               dHi_semiimplicit_fs=1.5, dHi_PETSc_rtol=1e-8, dHi_PETSc_abstol=1e-4, dt_ice_max=10.0, dt_ice_min=0.2)
    return mesh, C, ice


def MISMIPplus(h=2e3, seed=synthetic.SEED, calving_front=None, delaunay=True, profile="surface"):
    """config-files/benchmarks/MISMIP+/config_MISMIPplus_2km_spinup.cfg: 800x80 km,
    Schoof2005 (alpha^2 = 0.5, beta^2 = 1e4), A = 1.428e-17, Picard 5e-5 / 50 / relax 0.2,
    Krylov 1e-7 / 1e-5, BC u: west zero, others infinite; v: zero everywhere.  Uniform
    resolution h (the reference refines to 2 km only near the grounding line)."""
    nx = int(round(800e3 / h)) + 1
    ny = int(round(80e3 / h)) + 1
    mesh = synthetic.lattice_mesh(0.0, 800e3, -40e3, 40e3, nx, ny, seed=seed, delaunay=delaunay)
    ice = synthetic.geometry_MISMIPplus(mesh, calving_front=calving_front, profile=profile)
    C = Config(visc_it_norm_dUV_tol=5e-5, visc_it_nit=50, visc_it_relax=0.2,
               stress_balance_PETSc_rtol=1e-7, stress_balance_PETSc_abstol=1e-5,
               choice_sliding_law="Schoof2005", choice_ice_rheology_Glen="uniform",
               uniform_Glens_flow_factor=1.4280330398280316e-17,
               BC_u_west="zero", BC_u_east="infinite", BC_u_south="infinite", BC_u_north="infinite",
               BC_v_west="zero", BC_v_east="zero", BC_v_south="zero", BC_v_north="zero",
               dHi_semiimplicit_fs=1.5, dHi_PETSc_rtol=1e-8, dHi_PETSc_abstol=1e-6, dt_ice_max=10.0, dt_ice_min=0.5,
               BC_H_west="infinite", BC_H_east="zero", BC_H_south="infinite", BC_H_north="infinite")
    return mesh, C, ice


def antarctic(n_vertices=1_000_000, seed=synthetic.SEED, delaunay=True):
    """Synthetic Antarctic-scale case: config-files/config_ant_template.cfg domain
    (6080x6080 km) and solver keys (Zoet-Iverson phi = 45 deg, Huybrechts1992 rheology with
    m_enh_shelf = 0.58, Picard 1e-3 / 500 / relax 0.2, Krylov 1e-5 / 1e-5, vel_max 4000,
    beta_max 1e9, delta_v 1e-2, eps0^2 = 1e-12, subgrid exponent 0), synthetic dome."""
    n = int(round(np.sqrt(n_vertices)))
    mesh = synthetic.lattice_mesh(-3040e3, 3040e3, -3040e3, 3040e3, n, n, seed=seed, delaunay=delaunay)
    ice = synthetic.geometry_antarctic_dome(mesh, seed=seed)
    ice.till_friction_angle[:] = 45.0
    # synthetic temperature: cold surface, warmer base
    z = np.linspace(0.0, 1.0, mesh.nz)
    ice.Ti[:, :] = np.asfortranarray(243.15 + 25.0 * z[None, :] ** 2 + 0.0 * ice.Hi[:, None])
    C = Config(visc_it_norm_dUV_tol=1e-3, visc_it_nit=500, visc_it_relax=0.2, vel_max=4000.0,
               stress_balance_PETSc_rtol=1e-5, stress_balance_PETSc_abstol=1e-5,
               choice_sliding_law="Zoet-Iverson", subgrid_friction_exponent_on_B_grid=0.0,
               slid_beta_max=1e9, slid_delta_v=1e-2, Hi_min=0.1, Glens_flow_law_epsilon_sq_0=1e-12,
               choice_ice_rheology_Glen="Huybrechts1992", uniform_Glens_flow_factor=2.9377e-18,
               m_enh_sheet=1.0, m_enh_shelf=0.58)
    return mesh, C, ice
