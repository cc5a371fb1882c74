"""GPU parity on the NAMED workloads (BASELINE.json ``configs``), run to convergence (``-m gpu``).

Every case runs through the C ABI with the library's default Krylov set-up (``krylov_pc = auto`` ->
the multifrontal nested-dissection factorisation, ``ufe_solve_info.krylov_pc_used == 4``) and the
workload's own tolerances, exactly as ``bench.py`` runs it, and is compared with the oracle's
direct-solve Picard loop on the same mesh / inputs / config: same Picard count, u and v within the
north_star's 1e-6 relative L2 and max-norm.

At sizes where the oracle's sparse LU takes minutes per Picard iteration (MISMIP 8 km: 125 k
triangles; Antarctic-shaped 1e5 vertices: 198 k triangles) the GPU solve still runs to convergence,
and parity is checked through two size-independent properties instead of a full oracle loop:
(i) the first cold-start Picard iterate and (ii) ONE Picard step taken from the GPU's converged
state, by the GPU and by the oracle (closures, assembly, exact linear solve, relaxation at the
physically relevant state).
"""
import copy

import numpy as np
import pytest

from ufemism2_0_b200 import diva, experiments

pytestmark = pytest.mark.gpu

TOL_UV = 1e-6
ND_LU = 4


def _rel(a, b, ref):
    return (np.linalg.norm(a - b) / max(np.linalg.norm(ref), 1e-300), np.abs(a - b).max() / max(np.abs(ref).max(), 1e-300))


def _check_uv(S, D, names=("u_vav_b", "v_vav_b"), tol=TOL_UV):
    ref = np.concatenate([D[names[0]], D[names[1]]])
    worst = 0.0
    for k in names:
        r = _rel(getattr(S, k), D[k], ref)
        assert r[0] < tol and r[1] < tol, (k, r)
        worst = max(worst, *r)
    return worst


STATE = ("u_vav_b", "v_vav_b", "u_base_b", "v_base_b", "tau_bx_b", "tau_by_b", "eta_3D_b")


def _one_step_parity(S, oracle, mesh, C, ice):
    """One Picard step from the solver's current (converged) host state, by the GPU and by the oracle."""
    D = oracle.new_DIVA_state(mesh)
    for k in STATE:
        D[k] = np.array(getattr(S, k), copy=True, order="F")
    C1 = copy.copy(C)
    C1.visc_it_nit = 0                     # the loop leaves after one iteration (it > visc_it_nit)
    S.set_config(C1)
    info = S.solve_DIVA(ice)
    assert info.n_visc_its == 1
    nv, _ = oracle.solve_DIVA(mesh, ice, C1, D, "direct")
    assert nv == 1
    S.set_config(C)
    return _check_uv(S, D)


def test_bench_workload_MISMIPplus_2km_to_convergence(oracle):
    """The exact workload bench.py times (MISMIP+ 800 x 80 km, uniform 2 km Delaunay mesh, 32 000 triangles,
    config_MISMIPplus_2km_spinup.cfg keys), all Picard iterations, default Krylov set-up."""
    mesh, C, ice = experiments.MISMIPplus(2e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        assert info.flags == 0, "the Picard loop must converge within visc_it_nit"
        assert info.krylov_pc_used == ND_LU                      # auto -> nd_lu
        assert info.n_Axb_its <= 2 * info.n_visc_its             # exact preconditioner: one Krylov iteration per solve
        D = oracle.new_DIVA_state(mesh)
        tr = []
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct", trace=tr)
        assert tr[-1][1] < C.visc_it_norm_dUV_tol and info.n_visc_its == nv
        _check_uv(S, D)
        assert np.abs(S.u_3D_b - D["u_3D_b"]).max() <= TOL_UV * np.abs(D["u_3D_b"]).max()
        # a warm second call (0.1 % thicker ice) converges in a handful of iterations to the oracle's answer
        ice2 = copy.copy(ice)
        ice2.Hi = ice.Hi * 1.001
        from ufemism2_0_b200 import synthetic
        ice2.Hs = synthetic.ice_surface_elevation(ice2.Hi, ice.Hb, ice.SL)
        info2 = S.solve_DIVA(ice2)
        nv2, _ = oracle.solve_DIVA(mesh, ice2, C, D, "direct")
        assert info2.flags == 0 and info2.n_visc_its == nv2
        _check_uv(S, D)
    finally:
        S.close()


def test_MISMIP_8km_config(oracle):
    """config_MISMIP_8km_spinup_for_scaling.cfg on a 2000 x 2000 km, 8 km Delaunay mesh (125 000 triangles, ice-free
    ocean beyond the dome): converged on the GPU; first cold iterate and one step at the converged state vs the oracle."""
    mesh, C, ice = experiments.MISMIP_8km()
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        C1 = copy.copy(C)
        C1.visc_it_nit = 0
        S.set_config(C1)
        S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        oracle.solve_DIVA(mesh, ice, C1, D, "direct")
        _check_uv(S, D)                                      # first cold-start iterate
        S.set_config(C)
        for k in STATE:
            getattr(S, k)[...] = 0.0
        info = S.solve_DIVA(ice)
        assert info.flags == 0 and info.krylov_pc_used == ND_LU and info.n_visc_its <= C.visc_it_nit
        _one_step_parity(S, oracle, mesh, C, ice)
    finally:
        S.close()


@pytest.mark.parametrize("exp, L_km", [("A", 10), ("A", 20), ("A", 40), ("A", 80), ("C", 40)])
def test_ISMIP_HOM_domain_sizes(oracle, exp, L_km):
    """ISMIP-HOM A / C (hybrid DIVA, periodic BCs) on the remaining domain sizes of the reference's test set
    (config_ISMIP_HOM_{A,C}_<L>_DIVA.cfg; L = 160 is covered by test_gpu_parity.py), to convergence."""
    mesh, C, ice = experiments.ISMIP_HOM(exp, L_km * 1e3, 31)
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.flags == 0 and abs(info.n_visc_its - nv) <= 1 and info.krylov_pc_used == ND_LU
        _check_uv(S, D)
    finally:
        S.close()


def test_SSA_icestream_to_convergence(oracle):
    """SSA_icestream (Schoof 2006 ice stream, 'infinite_SSA_icestream' copy BCs, A = 1e-18, Picard tol 5e-8): all
    ~500 Picard iterations on an isotropic 41 x 41 Delaunay mesh, and the converged flow against Schoof's closed form."""
    mesh, C, ice = experiments.SSA_icestream(41, 41)
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_SSA(ice)
        R = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri))
        nv, _ = oracle.solve_SSA(mesh, ice, C, R, "direct")
        assert info.flags == 0 and info.n_visc_its == nv and info.krylov_pc_used == ND_LU
        ref = np.concatenate([R["u_b"], R["v_b"]])
        for k in ("u_b", "v_b"):
            r = _rel(getattr(S, k), R[k], ref)
            assert r[0] < TOL_UV and r[1] < TOL_UV, (k, r)
        # Schoof_SSA_solution.f90:36-61 at the triangle centroids (coarse mesh: a loose physical sanity band)
        y = mesh.TriGC[:, 1]
        ua, _ = oracle.Schoof2006_icestream(C.uniform_Glens_flow_factor, C.Glens_flow_law_exponent, 2000.0, 3e-4,
                                         C.refgeo_idealised_SSA_icestream_L, C.refgeo_idealised_SSA_icestream_m, y)
        mid = np.abs(mesh.TriGC[:, 0]) < 100e3
        assert np.abs(np.abs(S.u_b[mid]) - np.abs(ua[mid])).max() < 0.3 * np.abs(ua).max()      # 20 km cells across a 150 km half-width
    finally:
        S.close()


def test_antarctic_1e5_vertices_converged(oracle):
    """Synthetic Antarctic-shaped mesh, 1e5 vertices (198 450 triangles, 396 900 unknowns), config_ant_template.cfg keys:
    the wide-mesh case the banded block solve cannot hold.  Converged on the GPU with the default set-up (nd_lu);
    one Picard step at the converged state against the oracle's exact solve."""
    mesh, C, ice = experiments.antarctic(100_000)
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        assert info.flags == 0 and info.krylov_pc_used == ND_LU
        assert info.n_Axb_its <= 2 * info.n_visc_its
        _one_step_parity(S, oracle, mesh, C, ice)
    finally:
        S.close()


def test_auto_falls_back_when_nd_is_disabled(oracle, monkeypatch):
    """krylov_pc = auto with UFE_AUTO_ND=0: the banded exact block solve on a narrow mesh (krylov_pc_used == 2), same answer."""
    monkeypatch.setenv("UFE_AUTO_ND", "0")
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.visc_it_nit = 5
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        assert info.krylov_pc_used == 2
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.n_visc_its == nv
        _check_uv(S, D)
    finally:
        S.close()
