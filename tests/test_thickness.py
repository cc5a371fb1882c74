"""ice-thickness rates of change (calc_dHi_dt_explicit / _semiimplicit).

CPU part: pins the oracle restatement against the known answers of the reference's own component
test (src/UFEMISM/validation/component_tests/ct_mass_conservation.f90: 'linear' and 'periodic' test
ice sheets, :289-365) and checks the vectorised edge / Voronoi construction of the synthetic meshes
against the loop-for-loop restatement.  GPU part (``-m gpu``): the CUDA path behind the C ABI
against the oracle on the same seeded inputs.
"""
import copy

import numpy as np
import pytest

from ufemism2_0_b200 import config, diva, experiments, mesh_types, synthetic

PI = 3.141592653589793


def _mesh(nx=25, ny=21, jitter=0.2, delaunay=True, L=400e3):
    return synthetic.lattice_mesh(-L, L, -L * 0.75, L * 0.75, nx, ny, jitter=jitter, delaunay=delaunay)


def _ct_fields(mesh, E, kind):
    """setup_test_ice_sheet_linear / _periodic (ct_mass_conservation.f90:289-365) + the constant inputs of
    run_mass_cons_test_on_mesh_with_ice_sheet (:171-181)."""
    nV, nTri = mesh.nV, mesh.nTri
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    xc, yc = E.Tricc[:, 0], E.Tricc[:, 1]
    if kind == "linear":
        u0, H0 = 1.0 / 2000.0, 1000.0
        Hi = np.full(nV, H0)
        exact = np.full(nV, -2.0 * u0 * H0)
        u, v = u0 * xc, u0 * yc
    else:
        u0, H0 = 1000.0, 1000.0
        lam = 4.0 * (mesh.xmax - mesh.xmin) / (2 * PI)
        Hi = H0 * (2.0 + np.sin(3 * PI * x / lam) * np.sin(3 * PI * y / lam))
        dHdx = 3 * PI * H0 / lam * np.cos(3 * PI * x / lam) * np.sin(3 * PI * y / lam)
        dHdy = 3 * PI * H0 / lam * np.sin(3 * PI * x / lam) * np.cos(3 * PI * y / lam)
        uu, vv = u0 * np.sin(2 * PI * x / lam), u0 * np.sin(2 * PI * y / lam)
        dudx, dvdy = 2 * PI * u0 / lam * np.cos(2 * PI * x / lam), 2 * PI * u0 / lam * np.cos(2 * PI * y / lam)
        exact = -1.0 * (Hi * dudx + uu * dHdx + Hi * dvdy + vv * dHdy)
        u, v = u0 * np.sin(2 * PI * xc / lam), u0 * np.sin(2 * PI * yc / lam)
    z = np.zeros(nV)
    f = dict(Hi=Hi, Hb=z.copy(), SL=np.full(nV, -100.0), u_vav_b=u, v_vav_b=v, SMB=z.copy(), BMB=z.copy(), LMB=z.copy(),
             fraction_margin=np.ones(nV), mask_noice=np.zeros(nV, dtype=np.int32), dHi_dt_target=z.copy())
    return f, exact


def _edges_dict(E):
    return {k: getattr(E, k) for k in ("nE", "VE", "EV", "ETri", "EBI", "Tricc", "A", "Cw", "D_x", "D_y", "D")}


def _interior(mesh):
    """vertices that are not on the border and have no border neighbour"""
    ok = mesh.VBI == 0
    for vi in np.nonzero(ok)[0]:
        nb = mesh.C[vi, : mesh.nC[vi]] - 1
        if (mesh.VBI[nb] > 0).any():
            ok[vi] = False
    return ok


# ------------------------------------------------------------------------------------------
# CPU: mesh edges / Voronoi data
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("jitter, delaunay", [(0.0, False), (0.0, True), (0.2, True), (0.3, True)])
def test_mesh_edges_vectorised_equal_the_reference_loops(oracle, jitter, delaunay):
    mesh = _mesh(13, 9, jitter, delaunay)
    E = mesh_types.calc_mesh_edges(mesh)
    O = oracle.calc_mesh_edges_oracle(mesh)
    for k in ("VE", "EV", "ETri", "EBI"):
        assert np.array_equal(getattr(E, k), O[k]), k
    for k in ("Tricc", "A", "Cw", "D_x", "D_y", "D"):
        assert np.abs(getattr(E, k) - O[k]).max() <= 1e-13 * np.abs(O[k]).max(), k
    assert E.nE == int(mesh.nC.sum()) // 2
    # every edge is listed by exactly its two end vertices (mesh_edges.f90:27-31)
    cnt = np.bincount(E.VE[E.VE > 0], minlength=E.nE + 1)[1:]
    assert (cnt == 2).all()
    # Voronoi cells tile the domain (calc_Voronoi_cell_areas' own check, mesh_secondary.f90:176-180)
    assert abs(E.A.sum() / ((mesh.xmax - mesh.xmin) * (mesh.ymax - mesh.ymin)) - 1.0) < 1e-12


def test_dummy_mesh_5_edges(oracle):
    """The reference's 5-vertex seed mesh: 8 edges, the four border edges have one triangle."""
    mesh = mesh_types.dummy_mesh_5(0.0, 2.0, 0.0, 2.0)
    E = mesh_types.calc_mesh_edges(mesh)
    O = oracle.construct_mesh_edges(mesh)
    assert E.nE == 8 and np.array_equal(E.ETri, O["ETri"]) and np.array_equal(E.EV, O["EV"])
    assert sorted(E.EBI.tolist()) == [0, 0, 0, 0, 1, 3, 5, 7]
    assert ((E.ETri > 0).sum(axis=1) == np.where(E.EBI > 0, 1, 2)).all()
    assert abs(E.A.sum() - 4.0) < 1e-14 and abs(E.A[4] - 2.0) < 1e-14


# ------------------------------------------------------------------------------------------
# CPU: oracle pinned on the reference's component-test ice sheets
# ------------------------------------------------------------------------------------------
def test_oracle_mass_cons_linear_known_answer(oracle):
    """ct_mass_conservation.f90 'linear': H = H0, u = u0 x, v = u0 y  =>  dH/dt = -2 u0 H0 exactly (the upwind
    finite-volume flux integrates a linear velocity exactly), for every scheme the test runs."""
    mesh = _mesh()
    E = mesh_types.calc_mesh_edges(mesh)
    f, exact = _ct_fields(mesh, E, "linear")
    C = config.Config()
    inner = _interior(mesh)
    ex = oracle.calc_dHi_dt_explicit(mesh, _edges_dict(E), C, f, 0.1)
    assert np.abs(ex["dHi_dt"][inner] - exact[inner]).max() < 1e-12
    assert np.abs(ex["divQ"][inner] + exact[inner]).max() < 1e-12
    assert ex["dt"] == 0.1
    # semi-implicit / implicit / over-implicit (:195-211): uniform thinning rate u0-damped, -2 u0 H0 / (1 + 2 u0 fs dt)
    # away from the border (where the explicit value is imposed)
    deep = inner.copy()
    for _ in range(3):
        deep = np.array([deep[vi] and deep[mesh.C[vi, : mesh.nC[vi]] - 1].all() for vi in range(mesh.nV)])
    for fs in (0.5, 1.0, 1.5):
        C.dHi_semiimplicit_fs = fs
        si = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, f, 0.1)
        want = -2.0 * (1 / 2000.0) * 1000.0 / (1.0 + 2.0 * (1 / 2000.0) * fs * 0.1)
        assert np.abs(si["dHi_dt"][deep] - want).max() < 2e-4
        assert np.abs(si["AMB"]).max() == 0.0
        # the PETSc-defaults restatement reaches the same solution
        C.dHi_PETSc_rtol, C.dHi_PETSc_abstol = 1e-12, 1e-12
        si2 = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, f, 0.1, linear_solver="ksp")
        assert np.abs(si2["Hi_tplusdt"] - si["Hi_tplusdt"]).max() < 1e-8


def test_oracle_mass_cons_periodic_converges_to_the_analytical_rate(oracle):
    """ct_mass_conservation.f90 'periodic': first-order upwind scheme -> the error against the analytical
    dH/dt shrinks with the resolution."""
    errs = []
    for n in (21, 41, 81):
        mesh = _mesh(n, n, 0.15, True)
        E = mesh_types.calc_mesh_edges(mesh)
        f, exact = _ct_fields(mesh, E, "periodic")
        ex = oracle.calc_dHi_dt_explicit(mesh, _edges_dict(E), config.Config(), f, 0.1)
        inner = _interior(mesh)
        errs.append(np.sqrt(np.mean((-ex["divQ"][inner] - exact[inner]) ** 2)) / np.abs(exact).max())
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.05, errs


def _random_case(mesh, E, seed=11, with_prescribed=True):
    """A rough synthetic state exercising every branch: margins (fraction_margin < 1), no-ice mask, prescribed
    thickness, floating and grounded border vertices."""
    rng = np.random.default_rng(seed)
    nV = mesh.nV
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    L = mesh.xmax
    Hb = -300.0 + 400.0 * np.cos(2.5 * x / L) + 100.0 * rng.standard_normal(nV)
    Hi = np.maximum(0.0, 1500.0 * (1.0 - (np.hypot(x, 1.3 * y) / (0.95 * L)) ** 2)) + 30.0 * rng.random(nV)
    Hi[rng.random(nV) < 0.05] = 0.0
    xc, yc = E.Tricc[:, 0], E.Tricc[:, 1]
    u = 300.0 * xc / L + 40.0 * rng.standard_normal(mesh.nTri)
    v = 200.0 * yc / L + 40.0 * rng.standard_normal(mesh.nTri)
    fm = np.where(rng.random(nV) < 0.15, rng.random(nV), 1.0)
    f = dict(Hi=Hi, Hb=Hb, SL=np.full(nV, 20.0), u_vav_b=u, v_vav_b=v, SMB=0.3 * rng.standard_normal(nV),
             BMB=-0.5 * rng.random(nV), LMB=-0.2 * rng.random(nV), fraction_margin=fm,
             mask_noice=(rng.random(nV) < 0.04).astype(np.int32), dHi_dt_target=0.05 * rng.standard_normal(nV))
    if with_prescribed:
        f["BC_prescr_mask"] = (rng.random(nV) < 0.05).astype(np.int32)
        f["BC_prescr_Hi"] = 500.0 * rng.random(nV) - 50.0
    return f


def test_oracle_explicit_branches(oracle):
    mesh = _mesh(17, 13)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E)
    C = config.Config(BC_H_west="infinite", BC_H_south="infinite", BC_H_north="infinite", BC_H_east="zero", dt_ice_min=0.5)
    ex = oracle.calc_dHi_dt_explicit(mesh, _edges_dict(E), C, f, 2.0)
    east = np.isin(mesh.VBI, (3, 4)) & (f["mask_noice"] == 0) & (f["BC_prescr_mask"] == 0)
    assert (ex["Hi_tplusdt"][east] == 0.0).all()
    assert (ex["Hi_tplusdt"][f["mask_noice"] == 1] == 0.0).all()
    pm = (f["BC_prescr_mask"] == 1) & (f["mask_noice"] == 0)
    assert np.array_equal(ex["Hi_tplusdt"][pm], np.maximum(0.0, f["BC_prescr_Hi"][pm]))
    assert (ex["Hi_tplusdt"] >= 0.0).all() and ex["dt"] <= 2.0
    # dH/dt is consistent with the thickness change, and AMB holds what the limits and masks removed
    assert np.allclose(ex["dHi_dt"], (ex["Hi_tplusdt"] - f["Hi"]) / ex["dt"])
    # the upwind matrix: non-negative diagonal, non-positive off-diagonals, pattern [vi, C(vi,:)]
    M = ex["M_divQ"]
    for vi in (0, mesh.nV // 2, mesh.nV - 1):
        k0, k1 = M.ptr[vi] - 1, M.ptr[vi + 1] - 1
        assert M.ind[k0] == vi + 1 and np.array_equal(M.ind[k0 + 1:k1], mesh.C[vi, : mesh.nC[vi]])
        assert M.val[k0] >= 0.0 and (M.val[k0 + 1:k1] <= 0.0).all()


def test_flux_limited_timestep_as_written(oracle):
    """calc_flux_limited_timestep divides by max(dHi_dt, 1e-9) for thinning ice (utilities :184-190), i.e. by 1e-9."""
    C = config.Config(dt_ice_max=10.0, dt_ice_min=0.1)
    assert oracle.calc_flux_limited_timestep(C, np.array([100.0, 5.0]), np.array([-3.0, 1.0])) == 10.0
    assert oracle.calc_flux_limited_timestep(C, np.array([2e-9, 5.0]), np.array([-3.0, 1.0])) == 2.0
    assert oracle.calc_flux_limited_timestep(C, np.array([1e-12, 5.0]), np.array([-3.0, 1.0])) == 0.1


# ------------------------------------------------------------------------------------------
# GPU parity
# ------------------------------------------------------------------------------------------
TOL_VAL = 1e-12       # matrix entries / elementwise fields: same formula order in fp64
TOL_SOLVE = 1e-6      # Krylov solution against the oracle's direct solve (north_star tolerance)


def _solver(mesh, C):
    C = copy.deepcopy(C)
    C.choice_sliding_law, C.choice_ice_rheology_Glen = "Weertman", "uniform"
    return diva.initialise_DIVA_solver(mesh, C)


def _relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("bc, prescribed", [("zero", False), ("infinite", True), ("mixed", True)])
def test_gpu_explicit_matches_oracle(oracle, bc, prescribed):
    mesh = _mesh(41, 33)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E, seed=5, with_prescribed=prescribed)
    C = config.Config(dt_ice_min=0.5)
    if bc != "zero":
        C.BC_H_west = C.BC_H_south = C.BC_H_north = "infinite"
        C.BC_H_east = "zero" if bc == "mixed" else "infinite"
    want = oracle.calc_dHi_dt_explicit(mesh, _edges_dict(E), C, f, 2.0)
    S = _solver(mesh, C)
    S.set_mesh_edges(E)
    got = S.calc_dHi_dt_explicit(f, 2.0)
    M = S.get_thickness_matrix("M_divQ")
    assert np.array_equal(M.ptr, want["M_divQ"].ptr) and np.array_equal(M.ind, want["M_divQ"].ind)   # bit-exact pattern
    assert _relmax(M.val, want["M_divQ"].val) < TOL_VAL
    assert got["dt"] == want["dt"]
    for k in ("divQ", "dHi_dt", "Hi_tplusdt", "AMB"):
        assert _relmax(got[k], want[k]) < TOL_VAL, k
    S.close()


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["bicgstab", "gmres"])
@pytest.mark.parametrize("fs", [0.5, 1.5])
def test_gpu_semiimplicit_matches_oracle(oracle, method, fs):
    mesh = _mesh(41, 33)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E, seed=7)
    C = config.Config(dt_ice_min=0.5, dHi_semiimplicit_fs=fs, BC_H_west="infinite", BC_H_north="infinite",
                      b200_krylov_method=method)
    want = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, f, 1.0)
    S = _solver(mesh, C)
    S.set_mesh_edges(E)
    got = S.calc_dHi_dt_semiimplicit(f, 1.0)
    AA, bb = S.get_thickness_matrix("AA")
    assert np.array_equal(AA.ptr, want["AA"].ptr) and np.array_equal(AA.ind, want["AA"].ind)
    assert _relmax(AA.val, want["AA"].val) < TOL_VAL and _relmax(bb, want["bb"]) < TOL_VAL
    assert got["flags"] == 0 and 0 < got["n_Axb_its"] < 200
    assert np.linalg.norm(got["Hi_tplusdt"] - want["Hi_tplusdt"]) / np.linalg.norm(want["Hi_tplusdt"]) < TOL_SOLVE
    assert _relmax(got["Hi_tplusdt"], want["Hi_tplusdt"]) < TOL_SOLVE
    assert _relmax(got["divQ"], want["divQ"]) < TOL_VAL
    assert np.abs(got["AMB"]).max() == 0.0
    assert np.allclose(got["dHi_dt"], (got["Hi_tplusdt"] - f["Hi"]) / 1.0, rtol=0, atol=1e-9)
    S.close()


@pytest.mark.gpu
def test_gpu_mass_cons_component_test_ice_sheets(oracle):
    """ct_mass_conservation.f90 through the C ABI: the 'linear' ice sheet's exact answer and the 'periodic' one
    against the oracle, explicit and over-implicit."""
    mesh = _mesh(61, 49)
    E = mesh_types.calc_mesh_edges(mesh)
    C = config.Config()
    S = _solver(mesh, C)
    S.set_mesh_edges(E)
    inner = _interior(mesh)
    f, exact = _ct_fields(mesh, E, "linear")
    got = S.calc_dHi_dt_explicit(f, 0.1)
    assert np.abs(got["dHi_dt"][inner] - exact[inner]).max() < 1e-12 and got["dt"] == 0.1
    f, exact = _ct_fields(mesh, E, "periodic")
    want = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, f, 0.1)
    got = S.calc_dHi_dt_semiimplicit(f, 0.1)
    assert _relmax(got["Hi_tplusdt"], want["Hi_tplusdt"]) < TOL_SOLVE
    assert _relmax(got["dHi_dt"], want["dHi_dt"]) < 1e-4      # (H' - H)/dt amplifies the Krylov tolerance by 1/dt
    S.close()


@pytest.mark.gpu
def test_gpu_thickness_uses_resident_velocities(oracle):
    """The predictor-corrector step: solve_DIVA, then the thickness update reads u_vav_b / v_vav_b on the device."""
    mesh, C, ice = experiments.MISMIPplus(8e3)
    E = mesh_types.calc_mesh_edges(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    S.solve_DIVA(ice)
    S.set_mesh_edges(E)
    nV = mesh.nV
    f = dict(Hi=ice.Hi, Hb=ice.Hb, SL=ice.SL, SMB=np.full(nV, 0.3), BMB=np.zeros(nV),
             LMB=np.zeros(nV), fraction_margin=np.ones(nV), mask_noice=np.zeros(nV, dtype=np.int32), dHi_dt_target=np.zeros(nV))
    a = S.calc_dHi_dt_semiimplicit(f, 1.0)
    g = dict(f, u_vav_b=S.u_vav_b, v_vav_b=S.v_vav_b)
    b = S.calc_dHi_dt_semiimplicit(g, 1.0)
    assert np.array_equal(a["Hi_tplusdt"], b["Hi_tplusdt"]) and np.array_equal(a["divQ"], b["divQ"])
    want = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, g, 1.0)
    assert _relmax(a["Hi_tplusdt"], want["Hi_tplusdt"]) < TOL_SOLVE
    S.close()


@pytest.mark.gpu
def test_gpu_thickness_large_mesh_properties():
    """~250 k vertices (beyond what the oracle's loops finish in seconds): the 'linear' known answer holds at every
    interior vertex (round-off grows with |x| u0 / h ~ 1e3 cancelling terms), and the semi-implicit solve converges."""
    mesh = synthetic.lattice_mesh(-500e3, 500e3, -500e3, 500e3, 501, 501, jitter=0.2, delaunay=True)
    E = mesh_types.calc_mesh_edges(mesh)
    S = _solver(mesh, config.Config())
    S.set_mesh_edges(E)
    f, exact = _ct_fields(mesh, E, "linear")
    got = S.calc_dHi_dt_explicit(f, 0.1)
    inner = mesh.VBI == 0
    assert np.abs(got["divQ"][inner] + exact[inner]).max() < 1e-10
    si = S.calc_dHi_dt_semiimplicit(f, 0.1)
    assert si["flags"] == 0 and np.isfinite(si["Hi_tplusdt"]).all()
    AA, bb = S.get_thickness_matrix("AA")
    import scipy.sparse as sp
    A = sp.csr_matrix((AA.val, AA.ind.astype(np.int64) - 1, AA.ptr.astype(np.int64) - 1), shape=(mesh.nV, mesh.nV))
    r = A @ si["Hi_tplusdt"] - bb
    assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(bb) + 1e-6 * np.sqrt(mesh.nV)   # dHi_PETSc_rtol / abstol (unpreconditioned norm, slack 1)
    S.close()


@pytest.mark.gpu
def test_gpu_thickness_error_paths():
    mesh = _mesh(13, 9)
    E = mesh_types.calc_mesh_edges(mesh)
    S = _solver(mesh, config.Config())
    f = _random_case(mesh, E, with_prescribed=False)
    with pytest.raises(diva.UfeError, match="ufe_mesh_set_edges has not been called"):
        S.calc_dHi_dt_explicit(f, 1.0)
    S.set_mesh_edges(E)
    with pytest.raises(diva.UfeError, match="need to provide prescribed both Hi and mask"):
        S.calc_dHi_dt_explicit(dict(f, BC_prescr_mask=np.zeros(mesh.nV, dtype=np.int32)), 1.0)
    bad = copy.deepcopy(S.C)
    bad.BC_H_east = "periodic"
    S.C = bad
    with pytest.raises(diva.UfeError, match="unknown BC_H"):
        S.calc_dHi_dt_explicit(f, 1.0)
    S.close()


# ------------------------------------------------------------------------------------------
# calc_vertical_velocities (last part)
# ------------------------------------------------------------------------------------------
def _zeta_gradients(oracle, mesh, Hi, Hs):
    """dzeta_dx_ak, dzeta_dy_ak, dzeta_dz_ak as calc_zeta_gradients defines them (zeta_gradients.f90:91-131)."""
    Mx, My = oracle.calc_matrix_operators_mesh_a_a(mesh)
    H = np.maximum(0.1, Hi)
    dHi_dx, dHi_dy = oracle.spmv(Mx, Hi), oracle.spmv(My, Hi)
    dHs_dx, dHs_dy = oracle.spmv(Mx, Hs), oracle.spmv(My, Hs)
    z = mesh.zeta[None, :]
    zx = (1.0 / H)[:, None] * (dHs_dx[:, None] - z * dHi_dx[:, None])
    zy = (1.0 / H)[:, None] * (dHs_dy[:, None] - z * dHi_dy[:, None])
    zz = np.repeat((-1.0 / H)[:, None], mesh.nz, axis=1)
    return np.asfortranarray(zx), np.asfortranarray(zy), np.asfortranarray(zz)


def test_oracle_a_a_operators_are_exact_on_linear_functions(oracle):
    """ct_discretisation_mapping_derivatives.f90: first-order operators reproduce the gradient of a linear function."""
    mesh = _mesh(15, 11)
    Mx, My = oracle.calc_matrix_operators_mesh_a_a(mesh)
    f = 3.0 + 2e-4 * mesh.V[:, 0] - 5e-5 * mesh.V[:, 1]
    assert np.abs(oracle.spmv(Mx, f) - 2e-4).max() < 1e-15 and np.abs(oracle.spmv(My, f) + 5e-5).max() < 1e-15
    # row vi starts with vi, then its neighbours in C order (one flood-fill sweep is enough: n_neighbours_min = 2)
    for vi in (0, mesh.nV // 3, mesh.nV - 1):
        k0, k1 = Mx.ptr[vi] - 1, Mx.ptr[vi + 1] - 1
        assert Mx.ind[k0] == vi + 1 and np.array_equal(Mx.ind[k0 + 1:k1], mesh.C[vi, : mesh.nC[vi]])


def test_oracle_vertical_velocities_uniform_slab_known_answer(oracle):
    """Incompressibility on a flat slab (derivation in vertical_velocities.f90:23-63): u = u0 x, v = u0 y at every depth,
    H uniform, flat fixed base  =>  w(zeta) = -2 u0 H (1 - zeta)."""
    mesh = _mesh(17, 13)
    E = mesh_types.calc_mesh_edges(mesh)
    nV, nT, nz = mesh.nV, mesh.nTri, mesh.nz
    u0, H = 1e-3, 800.0
    u3b = np.asfortranarray(np.repeat((u0 * E.Tricc[:, 0])[:, None], nz, axis=1))
    v3b = np.asfortranarray(np.repeat((u0 * E.Tricc[:, 1])[:, None], nz, axis=1))
    u3 = np.asfortranarray(np.repeat((u0 * mesh.V[:, 0])[:, None], nz, axis=1))
    v3 = np.asfortranarray(np.repeat((u0 * mesh.V[:, 1])[:, None], nz, axis=1))
    ice = dict(Hi=np.full(nV, H), Hib=np.full(nV, -50.0), dHb_dt=np.zeros(nV), dHi_dt=np.zeros(nV),
               mask_grounded_ice=np.ones(nV, dtype=np.int32), mask_floating_ice=np.zeros(nV, dtype=np.int32),
               dzeta_dx_ak=np.zeros((nV, nz), order="F"), dzeta_dy_ak=np.zeros((nV, nz), order="F"),
               dzeta_dz_ak=np.full((nV, nz), -1.0 / H, order="F"))
    w = oracle.calc_vertical_velocities(mesh, _edges_dict(E), ice, u3b, v3b, u3, v3, np.zeros(nV))
    inner = _interior(mesh)
    want = -2.0 * u0 * H * (1.0 - mesh.zeta)[None, :]
    assert np.abs(w[inner] - want).max() < 1e-11
    # thin ice: no stretching (:149-152); no ice: no velocity (:131-134)
    ice["Hi"][5] = 5.0
    ice["mask_grounded_ice"][7] = 0
    w = oracle.calc_vertical_velocities(mesh, _edges_dict(E), ice, u3b, v3b, u3, v3, np.full(nV, -0.3))
    assert (w[5] == w[5, -1]).all() and (w[7] == 0.0).all() and abs(w[9, -1] + 0.3) < 1e-12


@pytest.mark.gpu
def test_gpu_a_a_operators_match_oracle(oracle):
    mesh = _mesh(31, 23, jitter=0.3)
    S = _solver(mesh, config.Config())
    S.set_mesh_edges(mesh_types.calc_mesh_edges(mesh))
    Mx, My = oracle.calc_matrix_operators_mesh_a_a(mesh)
    for which, want in (("ddx", Mx), ("ddy", My)):
        got = S.get_operator_a_a(which)
        assert np.array_equal(got.ptr, want.ptr) and np.array_equal(got.ind, want.ind)      # bit-exact pattern
        assert _relmax(got.val, want.val) < 1e-11
    S.close()


@pytest.mark.gpu
def test_gpu_vertical_velocities_match_oracle(oracle):
    """solve_DIVA -> calc_secondary_velocities -> calc_vertical_velocities, all on the resident fields."""
    mesh, C, ice = experiments.MISMIPplus(8e3)
    E = mesh_types.calc_mesh_edges(mesh)
    S = diva.initialise_DIVA_solver(mesh, C)
    S.solve_DIVA(ice)
    S.set_mesh_edges(E)
    nV, nz = mesh.nV, mesh.nz
    rng = np.random.default_rng(2)
    zx, zy, zz = _zeta_gradients(oracle, mesh, ice.Hi, ice.Hs)
    vin = dict(Hi=ice.Hi, Hib=ice.Hib, dHb_dt=0.01 * rng.standard_normal(nV), dHi_dt=0.5 * rng.standard_normal(nV),
               BMB=rng.standard_normal(nV), mask_grounded_ice=ice.mask_grounded_ice, mask_floating_ice=ice.mask_floating_ice,
               dzeta_dx_ak=zx, dzeta_dy_ak=zy, dzeta_dz_ak=zz)
    with pytest.raises(diva.UfeError, match="ufe_calc_secondary_velocities"):
        S.calc_vertical_velocities(vin)
    sec = S.calc_secondary_velocities()
    w = S.calc_vertical_velocities(vin)
    want = oracle.calc_vertical_velocities(mesh, _edges_dict(E), vin, S.u_3D_b, S.v_3D_b, sec["u_3D"], sec["v_3D"], vin["BMB"])
    assert np.isfinite(w).all() and np.abs(want).max() > 0.0
    assert _relmax(w, want) < 1e-10
    # thin / ice-free vertices follow the special cases
    noice = (np.asarray(ice.mask_grounded_ice) == 0) & (np.asarray(ice.mask_floating_ice) == 0)
    assert (w[noice] == 0.0).all()
    S.close()


# ------------------------------------------------------------------------------------------
# calc_dHi_dt: the dispatcher the predictor-corrector scheme calls (conservation_of_mass_main.f90:22-109)
# ------------------------------------------------------------------------------------------
def test_oracle_calc_dHi_dt_dispatch_and_clipping(oracle):
    mesh = _mesh(15, 11)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E, seed=3)
    f["SMB"] = f["SMB"] - 400.0 * (np.arange(mesh.nV) % 7 == 0)          # strong melt: the linear solve goes negative there
    C = config.Config(choice_ice_integration_method="none")
    r = oracle.calc_dHi_dt(mesh, _edges_dict(E), C, f, 2.0)
    assert np.array_equal(r["Hi_tplusdt"], f["Hi"]) and not r["dHi_dt"].any()
    C.choice_ice_integration_method = "semi-implicit"
    r = oracle.calc_dHi_dt(mesh, _edges_dict(E), C, f, 5.0)
    raw = oracle.calc_dHi_dt_semiimplicit(mesh, _edges_dict(E), C, f, 5.0)
    assert (raw["Hi_tplusdt"] < -0.1).any() and r["found_negative_vals"] and (r["Hi_tplusdt"] >= 0.0).all()
    assert np.allclose(r["AMB"], (r["Hi_tplusdt"] - raw["Hi_tplusdt"]) / 5.0)    # what the clipping added
    C.choice_ice_integration_method = "bogus"
    with pytest.raises(ValueError, match="unknown choice_ice_integration_method"):
        oracle.calc_dHi_dt(mesh, _edges_dict(E), C, f, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["none", "explicit", "semi-implicit"])
def test_gpu_calc_dHi_dt_matches_oracle(oracle, method):
    mesh = _mesh(33, 25)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E, seed=9)
    f["SMB"] = f["SMB"] - 400.0 * (np.arange(mesh.nV) % 7 == 0)
    C = config.Config(choice_ice_integration_method=method, BC_H_west="infinite", dt_ice_min=0.5)
    want = oracle.calc_dHi_dt(mesh, _edges_dict(E), C, f, 5.0)
    S = _solver(mesh, C)
    S.set_mesh_edges(E)
    got = S.calc_dHi_dt(f, 5.0)
    assert got["dt"] == want["dt"]
    assert bool(got["flags"] & 8) == want["found_negative_vals"]
    tol = TOL_SOLVE if method == "semi-implicit" else TOL_VAL
    assert _relmax(got["Hi_tplusdt"], want["Hi_tplusdt"]) < tol
    assert np.abs(got["dHi_dt"] - want["dHi_dt"]).max() <= tol * max(np.abs(want["dHi_dt"]).max(), 1.0) * 10
    assert np.abs(got["AMB"] - want["AMB"]).max() <= tol * max(np.abs(want["dHi_dt"]).max(), 1.0) * 10
    if method != "none":
        assert (got["Hi_tplusdt"] >= 0.0).all() and _relmax(got["divQ"], want["divQ"]) < TOL_VAL
    if method == "semi-implicit":
        assert want["found_negative_vals"]
    S.C = config.Config(choice_ice_integration_method="implicit")
    with pytest.raises(diva.UfeError, match="unknown choice_ice_integration_method"):
        S.calc_dHi_dt(f, 1.0)
    S.close()


@pytest.mark.gpu
def test_gpu_flux_limited_timestep(oracle):
    """calc_flux_limited_timestep actually limiting dt: one thinning vertex with 1e-9 m of ice => dt_lim = Hi / 1e-9 = 1 yr
    (the reference's formula as written), which the device finds with its block-min + atomicMin reduction."""
    mesh = synthetic.lattice_mesh(-400e3, 400e3, -300e3, 300e3, 25, 21, jitter=0.2, delaunay=True, nz=8)
    E = mesh_types.calc_mesh_edges(mesh)
    f = _random_case(mesh, E, seed=21, with_prescribed=False)
    k = int(np.nonzero((mesh.VBI == 0) & (f["mask_noice"] == 0))[0][7])
    f["Hi"][k], f["SMB"][k], f["fraction_margin"][k] = 1e-9, -50.0, 1.0
    C = config.Config(dt_ice_min=0.5, dt_ice_max=10.0, nz=8)
    want = oracle.calc_dHi_dt_explicit(mesh, _edges_dict(E), C, f, 2.0)
    assert want["dt"] == 1.0
    S = _solver(mesh, C)
    S.set_mesh_edges(E)
    got = S.calc_dHi_dt_explicit(f, 2.0)
    assert got["dt"] == 1.0 and _relmax(got["Hi_tplusdt"], want["Hi_tplusdt"]) < TOL_VAL
    S.close()


@pytest.mark.gpu
def test_gpu_vertical_velocities_nz8(oracle):
    mesh, C, ice = experiments.ISMIP_HOM("C", 80e3, 17)
    mesh8 = synthetic.lattice_mesh(mesh.xmin, mesh.xmax, mesh.ymin, mesh.ymax, 17, 17, jitter=0.2, nz=8)
    ice8 = synthetic.geometry_ISMIP_HOM_C(mesh8, 80e3)
    C.nz = 8
    C.visc_it_nit = 6
    S = diva.initialise_DIVA_solver(mesh8, C)
    S.solve_DIVA(ice8)
    E = mesh_types.calc_mesh_edges(mesh8)
    S.set_mesh_edges(E)
    sec = S.calc_secondary_velocities()
    nV, nz = mesh8.nV, 8
    zx, zy, zz = _zeta_gradients(oracle, mesh8, ice8.Hi, ice8.Hs)
    rng = np.random.default_rng(5)
    vin = dict(Hi=ice8.Hi, Hib=ice8.Hib, dHb_dt=np.zeros(nV), dHi_dt=0.1 * rng.standard_normal(nV), BMB=-rng.random(nV),
               mask_grounded_ice=ice8.mask_grounded_ice, mask_floating_ice=ice8.mask_floating_ice,
               dzeta_dx_ak=zx, dzeta_dy_ak=zy, dzeta_dz_ak=zz)
    w = S.calc_vertical_velocities(vin)
    want = oracle.calc_vertical_velocities(mesh8, _edges_dict(E), vin, S.u_3D_b, S.v_3D_b, sec["u_3D"], sec["v_3D"], vin["BMB"])
    assert w.shape == (nV, 8) and np.abs(want).max() > 0.0 and _relmax(w, want) < 1e-10
    S.close()
