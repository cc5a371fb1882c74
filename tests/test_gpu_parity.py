"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path behind the C ABI against
the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): CSR sparsity patterns and index maps bit-exact; converged
u/v velocity fields within 1e-6 relative L2 and max-norm.  Floating-point operator values:
1e-11 relative (the GPU evaluates the same formulas; only libm pow/exp/hypot differ).
"""
import copy

import numpy as np
import pytest

from ufemism2_0_b200 import diva, experiments, synthetic
from ufemism2_0_b200.capi import UfeError

pytestmark = pytest.mark.gpu

TOL_UV = 1e-6          # north_star: relative L2 and max-norm of the converged u, v
TOL_OPVAL = 1e-11      # operator / matrix values (same formula order, fp64)


def rel(a, b, ref=None):
    ref = b if ref is None else ref
    return (np.linalg.norm(a - b) / max(np.linalg.norm(ref), 1e-300),
            np.abs(a - b).max() / max(np.abs(ref).max(), 1e-300))


# ------------------------------------------------------------------------------------------
# L0: SpMV and linear solve, the reference's own known answers
# ------------------------------------------------------------------------------------------
def _csr(dense):
    dense = np.asarray(dense, dtype=float)
    m, n = dense.shape
    ptr, ind, val = [1], [], []
    for i in range(m):
        for j in range(n):
            if dense[i, j] != 0.0:
                ind.append(j + 1); val.append(dense[i, j])
        ptr.append(len(ind) + 1)
    return diva.CSRMatrix(m, n, 1, m, np.array(ptr, np.int32), np.array(ind, np.int32), np.array(val))


def _csr_rows(rows, n):
    ptr, ind, val = [1], [], []
    for r in rows:
        for c, v in r:
            ind.append(c); val.append(v)
        ptr.append(len(ind) + 1)
    return diva.CSRMatrix(len(rows), n, 1, len(rows), np.array(ptr, np.int32), np.array(ind, np.int32),
                          np.array(val, np.float64))


def test_spmv_known_answers():
    """ut_mpi_CSR_matrix_vector_multiplication.f90:188-321 and ut_petsc.f90:85-149 (unsorted
    columns, in add_entry order, exactly as in the reference tests)."""
    import test_oracle_golden as G
    assert np.array_equal(diva.multiply_CSR_matrix_with_vector(_csr_rows(G.EQ1, 7), G.X7), G.Y1)
    assert np.array_equal(diva.multiply_CSR_matrix_with_vector(_csr_rows(G.EQ2, 7), G.X7), G.Y2)
    assert np.array_equal(diva.multiply_CSR_matrix_with_vector(_csr_rows(G.PETSC, 6), np.arange(1., 7.)),
                          np.array([1., 8., 28., 59., 40., 105.]))
    # rank-local row blocks (rows 1-2 | 3-6), ptr local 1-based, as the two-rank PETSc test holds them
    A1 = _csr_rows(G.PETSC[2:], 6)
    A1 = diva.CSRMatrix(6, 6, 3, 6, A1.ptr, A1.ind, A1.val)
    assert np.array_equal(diva.multiply_CSR_matrix_with_vector(A1, np.arange(1., 7.)), np.array([28., 59., 40., 105.]))


def test_spmv_2D_random_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    m, n, nz = 1000, 700, 12
    dense = (rng.random((m, n)) < 0.01) * rng.standard_normal((m, n))
    A = _csr(dense)
    X = np.asfortranarray(rng.standard_normal((n, nz)))
    got = diva.multiply_CSR_matrix_with_vector(A, X)
    want = oracle.spmv_2D(oracle.CSR(m, n, 1, m, A.ptr, A.ind, A.val), X)
    assert rel(got, want)[1] < 1e-13
    # empty matrix rows and a single layer
    got1 = diva.multiply_CSR_matrix_with_vector(A, X[:, 0].copy())
    assert rel(got1, want[:, 0])[1] < 1e-13


@pytest.mark.parametrize("method", ["bicgstab", "gmres"])
def test_tridiagonal_solve_known_answer(method):
    """ut_mpi_CSR_matrix_solving.f90:217-270: x = [1,3.5,5,5.5,5,3.5,1]."""
    n = 7
    D = np.zeros((n, n))
    for i in range(n):
        if i in (0, n - 1):
            D[i, i] = 1.0
        else:
            D[i, i - 1:i + 2] = [-1.0, 2.0, -1.0]
    x, its, fl = diva.solve_matrix_equation_CSR(_csr(D), np.ones(n), np.zeros(n), 1e-12, 1e-14, method=method)
    assert fl == 0 and its <= 7
    assert np.abs(x - np.array([1, 3.5, 5, 5.5, 5, 3.5, 1.0])).max() < 1e-10


# ------------------------------------------------------------------------------------------
# operators: patterns bit-exact, values to round-off, application == oracle SpMV
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def homA(oracle):
    mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 41)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-11, 1e-10
    S = diva.initialise_DIVA_solver(mesh, C)
    yield mesh, C, ice, S
    S.close()


def test_operator_patterns_bit_exact(homA):
    mesh, C, ice, S = homA
    for fam, names in diva.FAMILIES.items():
        for w in names[1]:
            nm = ("M2_%s_b_b" % w) if fam == "b_b" else ("M_%s_%s" % (w, fam))
            G, R = S.get_operator(fam, w), mesh.ops[nm]
            assert np.array_equal(G.ptr, R.ptr), nm
            assert np.array_equal(G.ind, R.ind), nm
            assert rel(G.val, R.val)[1] < TOL_OPVAL, nm


def test_operator_patterns_bit_exact_irregular_mesh(oracle):
    """Strongly jittered, anisotropic lattice: exercises the BFS growth and border rows."""
    mesh = synthetic.lattice_mesh(0.0, 300e3, -20e3, 20e3, 61, 9, jitter=0.35, seed=11)
    oracle.calc_all_matrix_operators_mesh(mesh)
    S = diva.initialise_DIVA_solver(mesh, experiments.MISMIPplus(8e3)[1])
    try:
        for fam, names in diva.FAMILIES.items():
            for w in names[1]:
                nm = ("M2_%s_b_b" % w) if fam == "b_b" else ("M_%s_%s" % (w, fam))
                G, R = S.get_operator(fam, w), mesh.ops[nm]
                assert np.array_equal(G.ptr, R.ptr) and np.array_equal(G.ind, R.ind), nm
                assert rel(G.val, R.val)[1] < 1e-9, nm
    finally:
        S.close()


def test_apply_operators_match_oracle(homA, oracle):
    mesh, C, ice, S = homA
    rng = np.random.default_rng(5)
    da = rng.standard_normal(mesh.nV)
    db3 = np.asfortranarray(rng.standard_normal((mesh.nTri, mesh.nz)))
    assert rel(S.ddx_a_b_2D(da), oracle.spmv(mesh.ops["M_ddx_a_b"], da))[1] < 1e-12
    assert rel(S.map_a_b_2D(da), oracle.spmv(mesh.ops["M_map_a_b"], da))[1] < 1e-12
    assert rel(S.map_b_a_3D(db3), oracle.spmv_2D(mesh.ops["M_map_b_a"], db3))[1] < 1e-12
    # operator exactness on a linear function (ct_discretisation_mapping_derivatives.f90:533-575)
    f = 3.0 + 2e-5 * mesh.V[:, 0] - 1e-5 * mesh.V[:, 1]
    assert np.abs(S.ddx_a_b_2D(f) - 2e-5).max() < 1e-14
    assert np.abs(S.ddy_a_b_2D(f) + 1e-5).max() < 1e-14


# ------------------------------------------------------------------------------------------
# L1: assembly (pattern bit-exact) + linear solve against the direct solve
# ------------------------------------------------------------------------------------------
def _random_linearised_inputs(mesh, seed):
    rng = np.random.default_rng(seed)
    nT = mesh.nTri
    N_b = 1e9 * (1.0 + 0.3 * rng.random(nT))
    return dict(u_b=rng.standard_normal(nT), v_b=rng.standard_normal(nT), N_b=N_b,
                dN_dx_b=1e3 * rng.standard_normal(nT), dN_dy_b=1e3 * rng.standard_normal(nT),
                basal_friction_coefficient_b=1e3 * rng.random(nT), tau_dx_b=1e4 * rng.standard_normal(nT),
                tau_dy_b=1e4 * rng.standard_normal(nT))


@pytest.mark.parametrize("bc", ["infinite", "zero", "periodic_ISMIP-HOM", "mixed", "prescribed"])
def test_linearised_solve_and_stiffness_pattern(oracle, bc):
    L = 80e3
    mesh = synthetic.lattice_mesh(-L, L, -L, L, 25, 25)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C = experiments.ISMIP_HOM("A", L, 25)[1]
    kw = {}
    if bc == "mixed":
        C.BC_u_west, C.BC_u_east, C.BC_u_south, C.BC_u_north = "zero", "infinite", "infinite", "zero"
        C.BC_v_west, C.BC_v_east, C.BC_v_south, C.BC_v_north = "infinite", "zero", "zero", "infinite"
    elif bc == "prescribed":
        for c in "uv":
            for s in ("west", "east", "south", "north"):
                setattr(C, f"BC_{c}_{s}", "zero")
        rng = np.random.default_rng(9)
        mask = (rng.random(mesh.nTri) < 0.1).astype(np.int32)
        kw = dict(BC_prescr_mask_b=mask, BC_prescr_u_b=rng.standard_normal(mesh.nTri),
                  BC_prescr_v_b=rng.standard_normal(mesh.nTri))
    else:
        for c in "uv":
            for s in ("west", "east", "south", "north"):
                setattr(C, f"BC_{c}_{s}", bc)
    inp = _random_linearised_inputs(mesh, 21)
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        u, v, up, vp_, its = S.solve_SSA_DIVA_linearised(PETSc_rtol=1e-12, PETSc_abstol=1e-13, **inp, **kw)
        okw = {}
        if kw:
            okw = dict(bc_mask=kw["BC_prescr_mask_b"], bc_u=kw["BC_prescr_u_b"], bc_v=kw["BC_prescr_v_b"])
        ur, vr, upr, vpr, _, A, bb = oracle.solve_SSA_DIVA_linearised(
            mesh, C, inp["u_b"], inp["v_b"], inp["N_b"], inp["dN_dx_b"], inp["dN_dy_b"], inp["basal_friction_coefficient_b"],
            inp["tau_dx_b"], inp["tau_dy_b"], 0, 0, "direct", return_system=True, **okw)
        Ag, bg = S.get_stiffness_matrix()
        assert np.array_equal(Ag.ptr, A.ptr) and np.array_equal(Ag.ind, A.ind)      # bit-exact pattern
        assert rel(Ag.val, A.val)[1] < 1e-13 and rel(bg, bb)[1] < 1e-13
        assert np.array_equal(up, upr) and np.array_equal(vp_, vpr)
        scale = np.concatenate([ur, vr])
        assert rel(u, ur, scale)[0] < TOL_UV and rel(u, ur, scale)[1] < TOL_UV
        assert rel(v, vr, scale)[0] < TOL_UV and rel(v, vr, scale)[1] < TOL_UV
        assert its > 0
    finally:
        S.close()


# ------------------------------------------------------------------------------------------
# L2: full Picard solves
# ------------------------------------------------------------------------------------------
def _check_uv(S, D, names=("u_vav_b", "v_vav_b")):
    ref = np.concatenate([D["u_vav_b"], D["v_vav_b"]])
    for k in names:
        r = rel(getattr(S, k), D[k], ref)
        assert r[0] < TOL_UV and r[1] < TOL_UV, (k, r)


@pytest.mark.parametrize("method", ["bicgstab", "gmres"])
def test_ISMIP_HOM_A_DIVA(homA, oracle, method):
    mesh, C, ice, S = homA
    C2 = copy.copy(C)
    C2.b200_krylov_method = method
    S.set_config(C2)
    for k in S.STATE_FIELDS_B:
        getattr(S, k)[:] = 0
    S.eta_3D_b[:] = 0
    info = S.solve_DIVA(ice)
    D = oracle.new_DIVA_state(mesh)
    nv, _ = oracle.solve_DIVA(mesh, ice, C2, D, "direct")
    assert info.flags == 0 and info.gpu_launches > 0
    assert abs(info.n_visc_its - nv) <= 1
    _check_uv(S, D)
    ref3 = np.abs(D["u_3D_b"]).max()
    assert np.abs(S.u_3D_b - D["u_3D_b"]).max() / ref3 < TOL_UV
    assert np.abs(S.eta_3D_a - D["eta_3D_a"]).max() / np.abs(D["eta_3D_a"]).max() < 1e-5


def test_ISMIP_HOM_C_DIVA(oracle):
    mesh, C, ice = experiments.ISMIP_HOM("C", 160e3, 31)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-12
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert (info.flags & diva.KRYLOV_MAXIT) == 0
        assert abs(info.n_visc_its - nv) <= 1
        _check_uv(S, D)
    finally:
        S.close()


@pytest.mark.parametrize("lag", [0, 20])
def test_SSA_icestream_vs_oracle(oracle, lag):
    """SSA solve with the 'infinite_SSA_icestream' copy BC.  A = 1e-18 makes these systems far
    too stiff for point Jacobi (10 000-iteration cap), so this runs with the strip-LU block-Jacobi
    preconditioner (one GPU: the block is the whole matrix, solved exactly by block cyclic reduction).
    The first Picard iterates are compared (the reference needs ~500 iterations to converge here
    and the early, non-contractive part of that sequence amplifies round-off)."""
    mesh, C, ice = experiments.SSA_icestream(15, 61)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-12
    C.visc_it_nit = 4
    C.b200_krylov_pc, C.b200_krylov_pc_lag = "bjacobi_lu", lag
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_SSA(ice)
        R = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri))
        nv, _ = oracle.solve_SSA(mesh, ice, C, R, "direct")
        assert info.n_visc_its == nv and (info.flags & (diva.KRYLOV_MAXIT | diva.KRYLOV_DIVERGED)) == 0
        if lag == 0:
            assert info.n_Axb_its <= 2 * nv          # exact preconditioner: one or two Krylov steps per solve
        ref = np.concatenate([R["u_b"], R["v_b"]])
        for k in ("u_b", "v_b"):
            r = rel(getattr(S, k), R[k], ref)
            assert r[0] < TOL_UV and r[1] < TOL_UV, (k, r)
    finally:
        S.close()


@pytest.mark.parametrize("method", ["bicgstab", "gmres"])
def test_MISMIPplus_bjacobi_lu(oracle, method, lag=20):
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    C.visc_it_nit = 6
    C.b200_krylov_pc, C.b200_krylov_pc_lag, C.b200_krylov_method = "bjacobi_lu", lag, method
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.n_visc_its == nv and (info.flags & (diva.KRYLOV_MAXIT | diva.KRYLOV_DIVERGED)) == 0
        _check_uv(S, D)
        # the reference-layout matrix is unaffected by the preconditioner choice
        A, bb = S.get_stiffness_matrix()
        assert A.ptr[-1] - 1 == A.ind.size
    finally:
        S.close()


def test_MISMIPplus_8km_two_solves_warm_start(oracle):
    """Cold start then a warm second call with perturbed thickness (time-stepping pattern)."""
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    C.visc_it_nit = 8
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.n_visc_its == nv
        _check_uv(S, D)
        ice2 = copy.copy(ice)
        ice2.Hi = ice.Hi * 1.001
        ice2.Hs = synthetic.ice_surface_elevation(ice2.Hi, ice.Hb, ice.SL)
        info2 = S.solve_DIVA(ice2)
        nv2, _ = oracle.solve_DIVA(mesh, ice2, C, D, "direct")
        assert info2.n_visc_its == nv2
        _check_uv(S, D)
    finally:
        S.close()


def test_no_grounded_ice_returns_zero(homA):
    mesh, C, ice, S = homA
    ice2 = copy.copy(ice)
    ice2.mask_grounded_ice = np.zeros(mesh.nV, dtype=np.int32)
    S.u_vav_b[:] = 3.0
    info = S.solve_DIVA(ice2)
    assert info.n_visc_its == 0 and np.all(S.u_vav_b == 0.0) and np.all(S.u_3D_b == 0.0)


def test_resident_path_equals_host_path(homA):
    mesh, C, ice, S = homA
    for k in S.STATE_FIELDS_B:
        getattr(S, k)[:] = 0
    S.eta_3D_b[:] = 0
    a = S.solve_DIVA(ice)
    u_host = S.u_vav_b.copy()
    S.upload(ice, state=False)
    S.reset_state_resident()
    b = S.solve_DIVA_resident()
    S.download()
    assert a.n_visc_its == b.n_visc_its and a.n_Axb_its == b.n_Axb_its
    assert np.array_equal(u_host, S.u_vav_b)          # deterministic reductions: bitwise repeatable


def test_error_behaviour(homA):
    mesh, C, ice, S = homA
    with pytest.raises(UfeError):
        S.solve_DIVA(ice, BC_prescr_mask_b=np.zeros(mesh.nTri, np.int32))     # DIVA_main.f90:139-141
    bad = copy.copy(ice)
    bad.Hi = None
    with pytest.raises(UfeError):
        S.solve_DIVA(bad)


# ------------------------------------------------------------------------------------------
# size-independent properties at a larger size (no oracle): linearity of the operators,
# residual of the returned linear solve, idempotence of a converged warm restart
# ------------------------------------------------------------------------------------------
def test_properties_at_scale():
    mesh, C, ice = experiments.MISMIP_8km(16e3)
    C.visc_it_nit = 3
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        rng = np.random.default_rng(1)
        a, b = rng.standard_normal(mesh.nV), rng.standard_normal(mesh.nV)
        lin = S.ddx_a_b_2D(2.0 * a + 3.0 * b) - (2.0 * S.ddx_a_b_2D(a) + 3.0 * S.ddx_a_b_2D(b))
        assert np.abs(lin).max() < 1e-12 * np.abs(S.ddx_a_b_2D(a)).max() * 10
        S.solve_DIVA(ice)
        A, bb = S.get_stiffness_matrix()
        # row sums of free rows: derivative stencils annihilate constants => A*1 = -beta_eff on the diagonal block
        import scipy.sparse as sp
        M = sp.csr_matrix((A.val, A.ind - 1, A.ptr - 1), shape=(A.m, A.n))
        assert M.shape[0] == 2 * mesh.nTri and np.isfinite(A.val).all() and np.isfinite(bb).all()
    finally:
        S.close()


# ------------------------------------------------------------------------------------------
# committed golden fixture (tests/golden/oracle_lattice_9x7.npz): no oracle code at test time
# ------------------------------------------------------------------------------------------
def test_against_committed_golden_fixture():
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_lattice_9x7.npz"))
    L = float(G["L"])
    mesh = synthetic.lattice_mesh(-L, L, -L, L, 9, 7, jitter=0.25, seed=424242)
    assert np.array_equal(mesh.V, G["V"]) and np.array_equal(mesh.Tri, G["Tri"])
    _, C, _ = experiments.ISMIP_HOM("C", L, 9)
    ice = synthetic.geometry_ISMIP_HOM_C(mesh, L)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-12
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        for fam, names in diva.FAMILIES.items():
            for w in names[1]:
                nm = ("M2_%s_b_b" % w) if fam == "b_b" else ("M_%s_%s" % (w, fam))
                A = S.get_operator(fam, w)
                assert np.array_equal(A.ptr, G[nm + "_ptr"]) and np.array_equal(A.ind, G[nm + "_ind"]), nm
                assert rel(A.val, G[nm + "_val"])[1] < TOL_OPVAL, nm
        info = S.solve_DIVA(ice)
        assert info.n_visc_its == int(G["n_visc_its"])
        ref = np.concatenate([G["u_vav_b"], G["v_vav_b"]])
        for k in ("u_vav_b", "v_vav_b"):
            r = rel(getattr(S, k), G[k], ref)
            assert r[0] < TOL_UV and r[1] < TOL_UV, (k, r)
        A, bb = S.get_stiffness_matrix()
        assert np.array_equal(A.ptr, G["A_ptr"]) and np.array_equal(A.ind, G["A_ind"])
    finally:
        S.close()


# ------------------------------------------------------------------------------------------
# every sliding law / rheology / option of the path against the oracle (few Picard iterations)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("law", ["Weertman", "Coulomb", "Budd", "Tsai2015", "Schoof2005", "Zoet-Iverson", "no_sliding"])
def test_sliding_laws(oracle, law):
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.choice_sliding_law = law
    C.slid_beta_max = 1e9
    C.visc_it_nit = 3
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    ice.till_friction_angle[:] = 15.0
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.n_visc_its == nv
        _check_uv(S, D)
        r = rel(S.basal_friction_coefficient_a, D["basal_friction_coefficient_a"])
        assert r[1] < 1e-6, r          # beta ~ |u|^(1/m-1) magnifies the 1e-8 differences of u
    finally:
        S.close()


@pytest.mark.parametrize("variant", ["Huybrechts1992", "enh_interp", "no_crossterms", "no_GL_subgrid", "SSA"])
def test_rheology_and_options(oracle, variant):
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.visc_it_nit = 3
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    if variant in ("Huybrechts1992", "enh_interp"):
        C.choice_ice_rheology_Glen = "Huybrechts1992"
        C.m_enh_sheet, C.m_enh_shelf = 1.3, 0.6
        z = np.linspace(0.0, 1.0, mesh.nz)
        ice.Ti[:, :] = np.asfortranarray(250.0 + 20.0 * z[None, :] ** 2 + 1e-5 * mesh.V[:, 0:1])
        if variant == "enh_interp":
            C.choice_enhancement_factor_transition = "interp"
    elif variant == "no_crossterms":
        C.do_include_SSADIVA_crossterms = False
    elif variant == "no_GL_subgrid":
        C.do_GL_subgrid_friction = False
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        if variant == "SSA":
            info = S.solve_SSA(ice)
            R = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri))
            nv, _ = oracle.solve_SSA(mesh, ice, C, R, "direct")
            assert info.n_visc_its == nv
            ref = np.concatenate([R["u_b"], R["v_b"]])
            for k in ("u_b", "v_b"):
                r = rel(getattr(S, k), R[k], ref)
                assert r[0] < TOL_UV and r[1] < TOL_UV, (k, r)
        else:
            info = S.solve_DIVA(ice)
            D = oracle.new_DIVA_state(mesh)
            nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
            assert info.n_visc_its == nv
            _check_uv(S, D)
    finally:
        S.close()


def test_calc_secondary_velocities(homA, oracle):
    """the step right after the solve, on the device-resident 3-D velocities."""
    mesh, C, ice, S = homA
    S.set_config(C)
    S.solve_DIVA(ice)
    got = S.calc_secondary_velocities()
    want = oracle.calc_secondary_velocities(mesh, S.u_3D_b, S.v_3D_b)
    assert set(got) == set(want)
    for k, w in want.items():
        scale = max(np.abs(w).max(), 1e-300)
        assert np.abs(got[k] - w).max() / scale < 1e-12, k
    assert np.array_equal(got["u_surf"], got["u_3D"][:, 0])      # same sums in the same order


def test_two_rank_nccl_solve_if_two_gpus():
    """Strip partition over two GPUs (halo exchange + all-reduces over NCCL, replicated exact
    preconditioner); runs tests/tools/multi_gpu_check.py under torchrun.  Skipped on one-GPU boxes."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_prescribed_velocity_BC_through_solve_DIVA(oracle):
    """solve_DIVA with the optional BC_prescr_* arguments (DIVA_main.f90:137-152): prescribed rows get
    a unit diagonal and the prescribed value; the pattern cache is rebuilt when the mask changes."""
    mesh, C, ice = experiments.MISMIPplus(8e3)
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.visc_it_nit = 3
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    rng = np.random.default_rng(4)
    mask = (mesh.TriGC[:, 0] < 100e3).astype(np.int32)
    bu, bv = 5.0 + rng.random(mesh.nTri), rng.random(mesh.nTri) - 0.5
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        for use_bc in (True, False, True):
            for k in S.STATE_FIELDS_B:
                getattr(S, k)[:] = 0
            S.eta_3D_b[:] = 0
            D = oracle.new_DIVA_state(mesh)
            if use_bc:
                info = S.solve_DIVA(ice, BC_prescr_mask_b=mask, BC_prescr_u_b=bu, BC_prescr_v_b=bv)
                nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct", bc_mask=mask, bc_u=bu, bc_v=bv)
            else:
                info = S.solve_DIVA(ice)
                nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
            assert info.n_visc_its == nv
            _check_uv(S, D)
    finally:
        S.close()


def test_error_paths_on_device():
    mesh, C, ice = experiments.ISMIP_HOM("A", 80e3, 9)
    C2 = copy.copy(C)
    C2.choice_ice_rheology_Glen = "Huybrechts1992"
    S = diva.initialise_DIVA_solver(mesh, C2)
    try:
        bad = copy.copy(ice)
        bad.Ti = None
        with pytest.raises(UfeError):               # Huybrechts1992 needs Ti
            S.solve_DIVA(bad)
        with pytest.raises(UfeError):               # no assembled matrix before the first solve
            S.get_stiffness_matrix()
    finally:
        S.close()
    with pytest.raises(UfeError):                   # nz outside [2, 32]
        m2 = copy.copy(mesh)
        m2.nz = 40
        m2.zeta = np.linspace(0.0, 1.0, 40)
        diva.initialise_DIVA_solver(m2, C)


def test_remap_DIVA_solver_lifecycle(oracle):
    """remap_DIVA_solver (DIVA_main.f90:264-373): b -> a (old mesh), a(old) -> a(new) by the caller's remapping, a -> b
    (new mesh); checked against the same chain of oracle SpMVs, and the remapped solver then solves."""
    from scipy.spatial import cKDTree
    mesh_old, C, ice_old = experiments.MISMIPplus(16e3)
    mesh_new, _, ice_new = experiments.MISMIPplus(10e3)
    S = diva.initialise_DIVA_solver(mesh_old, C)
    S.solve_DIVA(ice_old)
    u_old, eta_old = S.u_vav_b.copy(), S.eta_3D_b.copy()
    nearest = cKDTree(mesh_old.V).query(mesh_new.V)[1]           # stand-in for the conservative mesh-to-mesh remapping
    S2 = diva.remap_DIVA_solver(S, mesh_new, lambda d: d[nearest])
    mesh_old.ops = oracle.calc_all_matrix_operators_mesh(mesh_old)
    mesh_new.ops = oracle.calc_all_matrix_operators_mesh(mesh_new)
    want_u = oracle.spmv(mesh_new.ops["M_map_a_b"], oracle.spmv(mesh_old.ops["M_map_b_a"], u_old)[nearest])
    want_eta = oracle.spmv_2D(mesh_new.ops["M_map_a_b"], np.asfortranarray(oracle.spmv_2D(mesh_old.ops["M_map_b_a"], eta_old)[nearest]))
    assert S2.u_vav_b.shape == (mesh_new.nTri,) and rel(S2.u_vav_b, want_u)[1] < 1e-12
    assert rel(S2.eta_3D_b, want_eta)[1] < 1e-12
    assert not S2.u_base_b.any()                                  # everything else is reallocated
    info = S2.solve_DIVA(ice_new)
    assert info.n_visc_its > 0 and np.isfinite(S2.u_vav_b).all()
    with pytest.raises(UfeError, match="remapped"):
        diva.remap_DIVA_solver(S2, mesh_old, lambda d: d)


# ------------------------------------------------------------------------------------------
# multifrontal nested-dissection exact solver (ufe_nd_solver_*): the stiffness system of a linearised
# solve, taken from the device in the reference layout, against a CPU sparse LU of the same system
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("workload, leaf", [("mismipplus_8km", 48), ("ismip_hom_c", 24), ("mismip_16km", 96)])
def test_nd_multifrontal_solver_matches_sparse_LU(workload, leaf):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from ufemism2_0_b200 import nd
    if workload == "mismipplus_8km":
        mesh, C, ice = experiments.MISMIPplus(8e3)
    elif workload == "ismip_hom_c":
        mesh, C, ice = experiments.ISMIP_HOM("C", 80e3, 31)
    else:
        mesh, C, ice = experiments.MISMIP_8km(16e3)
    C.visc_it_nit = 2
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        S.solve_DIVA(ice)
        A, bb = S.get_stiffness_matrix()
    finally:
        S.close()
    N = 2 * mesh.nTri
    ptr, ind = (A.ptr - 1).astype(np.int32), (A.ind - 1).astype(np.int32)
    M = sp.csr_matrix((A.val, ind, ptr), shape=(N, N))
    xr = spla.splu(M.tocsc()).solve(bb)
    sol = nd.Solver(np.asarray(mesh.TriGC), A.ptr, A.ind, leaf)     # the reference's 1-based arrays, as the C ABI takes them
    try:
        sol.factor(A.val)
        x0, r0 = sol.solve(bb, n_refine=0)
        x1, r1 = sol.solve(bb, n_refine=2)
        assert r0 < 1e-8 and r1 < 1e-13, (r0, r1)
        assert np.abs(x1 - xr).max() <= 1e-9 * np.abs(xr).max()
        # linearity and a second factorisation with new values on the same analysis
        rng = np.random.default_rng(3)
        b2 = rng.standard_normal(N)
        x2, _ = sol.solve(b2, n_refine=2)
        x3, _ = sol.solve(2.0 * bb - 3.0 * b2, n_refine=2)
        assert np.abs(x3 - (2.0 * x1 - 3.0 * x2)).max() <= 1e-9 * max(np.abs(x1).max(), np.abs(x2).max()) * 5
        M2 = M + sp.diags(np.abs(M.diagonal()) * 0.5)
        M2 = M2.tocsr(); M2.sort_indices()
        Mo = M.copy(); Mo.sort_indices()
        if np.array_equal(M2.indices, Mo.indices) and np.array_equal(ind, Mo.indices):
            sol.factor(M2.data)
            x4, r4 = sol.solve(bb, n_refine=2)
            assert r4 < 1e-13 and np.abs(M2 @ x4 - bb).max() <= 1e-10 * np.abs(bb).max()
        info = sol.info()
        assert info["factor_ms"] > 0 and info["front_bytes"] > 0
    finally:
        sol.close()


@pytest.mark.parametrize("late_schur_from, tensor_cores, lazy_zero", [("256", "1", "1"), ("64", "1", "0"), ("64", "0", "1"), ("0", "1", "0")])
def test_nd_factorisation_variants_agree(monkeypatch, late_schur_from, tensor_cores, lazy_zero):
    """The organisation of the trailing updates is a performance choice, not a numerical one: right-looking K = 64 passes,
    the late Schur-complement pass with K = p (UFE_ND_SCHUR_MIN_P), fp64 SIMT or fp64 tensor cores (UFE_ND_UPD_MMA), one memset
    of all fronts or level-by-level zeroing through a write-only extend-add (UFE_ND_LAZY_ZERO) give
    the sparse-LU solution of a wide-mesh stiffness system (fronts of up to ~900 unknowns, pivot counts above 256)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from ufemism2_0_b200 import nd
    mesh, C, ice = experiments.antarctic(20000)
    C.visc_it_nit, C.b200_krylov_pc = 1, "nd_lu"
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        S.solve_DIVA(ice)
        A, bb = S.get_stiffness_matrix()
    finally:
        S.close()
    N = 2 * mesh.nTri
    M = sp.csr_matrix((A.val, (A.ind - 1).astype(np.int32), (A.ptr - 1).astype(np.int32)), shape=(N, N))
    xr = spla.splu(M.tocsc()).solve(bb)
    monkeypatch.setenv("UFE_ND_SCHUR_MIN_P", late_schur_from)
    monkeypatch.setenv("UFE_ND_UPD_MMA", tensor_cores)
    monkeypatch.setenv("UFE_ND_LAZY_ZERO", lazy_zero)
    sol = nd.Solver(np.asarray(mesh.TriGC), A.ptr, A.ind, 64)
    try:
        assert sol.info()["max_front"] > 512
        sol.factor(A.val)
        sol.factor(A.val)                 # a second factorisation over the used fronts (nothing may survive from the first)
        x0, r0 = sol.solve(bb, n_refine=0)
        x1, r1 = sol.solve(bb, n_refine=2)
        assert r0 < 1e-10 and r1 < 1e-13, (r0, r1)
        assert np.abs(x1 - xr).max() <= 1e-9 * np.abs(xr).max()
    finally:
        sol.close()


@pytest.mark.parametrize("workload, method, lag", [("mismipplus_8km", "bicgstab", 0), ("mismipplus_8km", "gmres", 20),
                                                   ("ismip_hom_c", "bicgstab", 0)])
def test_DIVA_with_nd_lu_preconditioner(oracle, workload, method, lag):
    """krylov_pc = 'nd_lu': the multifrontal solver as the exact preconditioner of the Picard loop's linear solves;
    same Picard count and velocities as the oracle's direct solve, one or two Krylov iterations per fresh factorisation."""
    if workload == "mismipplus_8km":
        mesh, C, ice = experiments.MISMIPplus(8e3)
        C.visc_it_nit = 6
    else:
        mesh, C, ice = experiments.ISMIP_HOM("C", 80e3, 21)
        C.visc_it_nit = 8
    oracle.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    C.b200_krylov_pc, C.b200_krylov_pc_lag, C.b200_krylov_method = "nd_lu", lag, method
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        info = S.solve_DIVA(ice)
        D = oracle.new_DIVA_state(mesh)
        nv, _ = oracle.solve_DIVA(mesh, ice, C, D, "direct")
        assert info.n_visc_its == nv and (info.flags & (diva.KRYLOV_MAXIT | diva.KRYLOV_DIVERGED)) == 0
        assert info.krylov_pc_used == 4
        if lag == 0:
            assert info.n_Axb_its <= 3 * info.n_visc_its, (info.n_Axb_its, info.n_visc_its)
        _check_uv(S, D)
    finally:
        S.close()


# ------------------------------------------------------------------------------------------
# L0 as the reference calls it: solve_matrix_equation_CSR on the ranks of a handle (one rank here; the two-rank row
# blocks are exercised by tests/tools/multi_gpu_check.py)
# ------------------------------------------------------------------------------------------
def _scipy_csr(A):
    import scipy.sparse as sp
    return sp.csr_matrix((A.val, A.ind - 1, A.ptr - 1), shape=(A.i2 - A.i1 + 1, A.n))


@pytest.mark.parametrize("pc", ["auto", "nd_lu", "bjacobi_lu", "jacobi"])
def test_L0_handle_solves_the_DIVA_stiffness_system(pc):
    """solve_matrix_equation_CSR_PETSc( A_CSR, bb, xx, ...) (petsc_basic.f90:32-64, call site
    solve_linearised_SSA_DIVA.f90:159) with the stiffness matrix the reference would pass: every preconditioner is
    reachable from the handle's config; x equals the sparse-LU solution of the same system."""
    import scipy.sparse.linalg as spla
    mesh, C, ice = experiments.MISMIPplus(8e3)
    C.visc_it_nit = 3
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        S.solve_DIVA(ice)
        A, bb = S.get_stiffness_matrix()
        want = spla.splu(_scipy_csr(A).tocsc()).solve(bb)
        C2 = copy.deepcopy(C)
        C2.b200_krylov_pc = pc
        if pc == "jacobi":
            C2.b200_krylov_maxits = 200          # point Jacobi does not converge on this operator: the cap must be reported
        S.set_config(C2)
        x, its, fl, used = S.solve_matrix_equation_CSR(A, bb, np.zeros_like(bb), 1e-12, 1e-11)
        if pc == "jacobi":
            assert used == 0 and its <= 200 and (fl != 0 or rel(x, want)[1] < 1e-6), (its, fl)
            return
        assert used == (2 if pc == "bjacobi_lu" else 4), used
        assert fl == 0 and its <= 3, (its, fl)
        assert rel(x, want)[1] < 1e-9
    finally:
        S.close()


@pytest.mark.parametrize("method", ["bicgstab", "gmres"])
def test_L0_handle_generic_systems(method):
    """Systems that are not the handle's stiffness matrix (the thickness system of conservation_of_mass_semiimplicit.f90:155,
    the reference's tridiagonal known answer ut_mpi_CSR_matrix_solving.f90:217-270): banded block solve or point Jacobi."""
    mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 11)
    C.b200_krylov_method = method
    S = diva.initialise_DIVA_solver(mesh, C)
    try:
        n = 7
        D = np.zeros((n, n))
        for i in range(n):
            if i in (0, n - 1):
                D[i, i] = 1.0
            else:
                D[i, i - 1:i + 2] = [-1.0, 2.0, -1.0]
        x, its, fl, used = S.solve_matrix_equation_CSR(_csr(D), np.ones(n), np.zeros(n), 1e-12, 1e-14)
        assert fl == 0 and its <= 7 and used in (0, 2)
        assert np.abs(x - np.array([1, 3.5, 5, 5.5, 5, 3.5, 1.0])).max() < 1e-10
        # a random diagonally dominant banded system, unsorted columns
        rng = np.random.default_rng(5)
        n = 3000
        rows = []
        for i in range(n):
            cols = sorted(set(int(c) for c in np.clip(i + rng.integers(-40, 41, 8), 0, n - 1)) - {i})
            r = [(c + 1, float(rng.standard_normal())) for c in cols]
            rng.shuffle(r)
            r.insert(len(r) // 2, (i + 1, 10.0 + float(rng.random())))
            rows.append(r)
        A = _csr_rows(rows, n)
        b = rng.standard_normal(n)
        import scipy.sparse.linalg as spla
        want = spla.splu(_scipy_csr(A).tocsc()).solve(b)
        x, its, fl, used = S.solve_matrix_equation_CSR(A, b, np.zeros(n), 1e-13, 1e-14)
        assert fl == 0 and rel(x, want)[1] < 1e-10, (its, fl, used)
        # the reference's size check (petsc_basic.f90:91)
        with pytest.raises(UfeError, match="sub-sizes"):
            S.solve_matrix_equation_CSR(A, b[:-1], np.zeros(n - 1), 1e-13, 1e-14)
        half = diva.CSRMatrix(n, n, 1, n // 2, A.ptr[:n // 2 + 1], A.ind, A.val)
        with pytest.raises(UfeError, match="all rows"):
            S.solve_matrix_equation_CSR(half, b[:n // 2], np.zeros(n // 2), 1e-13, 1e-14)
    finally:
        S.close()
