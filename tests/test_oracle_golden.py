"""Pins the oracle (oracle/) against the known-answer vectors the reference's own tests
hold for this path.  CPU only."""
import numpy as np
import pytest

from ufemism2_0_b200 import mesh_types, synthetic


def _csr(O, rows, n):
    """rows: list of [(col, val), ...] in insertion order (add_entry_CSR_dist order)."""
    ptr, ind, val = [1], [], []
    for r in rows:
        for c, v in r:
            ind.append(c)
            val.append(v)
        ptr.append(len(ind) + 1)
    return O.CSR(len(rows), n, 1, len(rows), np.array(ptr, np.int32), np.array(ind, np.int32),
                 np.array(val, np.float64))


# src/UPSY/validation/unit_tests/ut_mpi_CSR_matrix_vector_multiplication.f90:188-321
EQ1 = [[(1, 1.)], [(2, 2.), (3, 3.)], [(2, 4.), (5, 1.)], [(3, 2.), (4, 3.), (6, 4.)], [(4, 1.), (5, 2.)],
       [(5, 3.), (7, 4.)], [(6, 1.), (7, 2.)]]
EQ2 = [[(1, 1.), (5, 5.)], [(2, 2.), (3, 3.), (6, 6.)], [(2, 4.), (5, 1.)],
       [(1, 5.), (3, 2.), (4, 3.), (6, 4.), (7, 5.)], [(2, 6.), (4, 1.), (5, 2.)],
       [(1, 5.), (5, 3.), (7, 4.)], [(2, 6.), (6, 1.), (7, 2.)]]
X7 = np.array([1., 2., 3., 4., 1., 2., 3.])
Y1 = np.array([1., 13., 9., 26., 6., 15., 8.])
Y2 = np.array([6., 25., 9., 46., 18., 20., 20.])
# ut_petsc.f90:85-149 (two ranks: rows 1-2 and 3-6)
PETSC = [[(1, 1.)], [(1, 2.), (2, 3.)], [(2, 4.), (4, 5.)], [(4, 6.), (5, 7.)], [(5, 8.)], [(5, 9.), (6, 10.)]]
# ut_mpi_CSR_matrix_solving.f90:217-270
TRI = [[(1, 1.)]] + [[(i - 1, -1.), (i, 2.), (i + 1, -1.)] for i in range(2, 7)] + [[(7, 1.)]]
X_TRI = np.array([1., 3.5, 5., 5.5, 5., 3.5, 1.])


def test_spmv_known_answers(oracle):
    assert np.array_equal(oracle.spmv(_csr(oracle, EQ1, 7), X7), Y1)
    assert np.array_equal(oracle.spmv(_csr(oracle, EQ2, 7), X7), Y2)
    assert np.array_equal(oracle.spmv(_csr(oracle, PETSC, 6), np.arange(1., 7.)),
                          np.array([1., 8., 28., 59., 40., 105.]))


def test_spmv_distributed_rows_match_reference_ranks(oracle):
    # the PETSc test splits rows 1-2 | 3-6 over two ranks; ptr is rank-local 1-based
    A0 = _csr(oracle, PETSC[:2], 6)
    A1 = _csr(oracle, PETSC[2:], 6)
    assert list(A0.ptr) == [1, 2, 4] and list(A1.ptr) == [1, 3, 5, 6, 8]
    x = np.arange(1., 7.)
    assert np.array_equal(np.concatenate([oracle.spmv(A0, x), oracle.spmv(A1, x)]),
                          np.array([1., 8., 28., 59., 40., 105.]))


def test_tridiagonal_solve_known_answer(oracle):
    import ctypes as ct
    A = _csr(oracle, TRI, 7)
    b = np.ones(7)
    x = np.zeros(7)
    # Jacobi(500, 1e-7), checked to 1e-5 like ut_mpi_CSR_matrix_solving.f90:100-112
    oracle.lib().ora_jacobi(7, oracle._p(A.ptr), oracle._p(A.ind), oracle._p(A.val), oracle._p(b), oracle._p(x),
                            500, ct.c_double(1e-7))
    assert np.abs(x - X_TRI).max() < 1e-5
    for nranks in (1, 2):
        xk, its, reason, _ = oracle.ksp_solve(A, b, 1e-10, 1e-12, nranks=nranks)
        assert reason in (2, 3) and np.abs(xk - X_TRI).max() < 1e-8
    assert np.abs(oracle.direct_solve(A, b) - X_TRI).max() < 1e-12


def test_partition_list(oracle):
    # mpi_distributed_memory.f90:42-68, incl. the ntot <= 2n branch
    for ntot in (0, 1, 5, 7, 8, 9, 100, 101, 1000003):
        for n in (1, 2, 3, 7, 8):
            ranges = [oracle.partition_list(ntot, i, n) for i in range(n)]
            if ntot > 2 * n:
                assert ranges[0][0] == 1 and ranges[-1][1] == ntot
                for (a1, a2), (b1, b2) in zip(ranges[:-1], ranges[1:]):
                    assert b1 == a2 + 1
                sizes = [b - a + 1 for a, b in ranges]
                assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            else:
                assert ranges[0] == (1, ntot) and all(r == (1, 0) for r in ranges[1:])
    assert oracle.partition_list(10, 0, 3) == (1, 4) and oracle.partition_list(10, 2, 3) == (8, 10)


def test_dummy_mesh_5_connectivity():
    # mesh_dummy_meshes.f90:61-113: arrays the reference hard-codes, rebuilt from (V, Tri, VBI)
    m = mesh_types.dummy_mesh_5(0., 1., 0., 1.)
    assert m.nC.tolist() == [3, 3, 3, 3, 4] and m.niTri.tolist() == [2, 2, 2, 2, 4]
    assert m.C[:4, :3].tolist() == [[2, 5, 4], [3, 5, 1], [4, 5, 2], [1, 5, 3]]
    assert m.iTri[:4, :2].tolist() == [[1, 4], [2, 1], [3, 2], [4, 3]]
    assert m.TriC.tolist() == [[2, 4, 0], [3, 1, 0], [4, 2, 0], [1, 3, 0]]
    # interior vertex: same cyclic order (start is arbitrary)
    c5 = m.C[4, :4].tolist()
    assert c5 in ([1, 2, 3, 4], [2, 3, 4, 1], [3, 4, 1, 2], [4, 1, 2, 3])
    assert m.TriBI.tolist() == [4, 4, 2, 8]      # border trace of calc_TriBI, mesh_secondary.f90:72-135


@pytest.fixture(scope="module")
def mesh41(oracle):
    m = synthetic.lattice_mesh(-400e3, 400e3, -400e3, 400e3, 41, 41)
    oracle.calc_all_matrix_operators_mesh(m)
    return m


def test_mesh_conventions(mesh41):
    m = mesh41
    T = m.Tri.astype(np.int64) - 1
    x, y = m.V[:, 0], m.V[:, 1]
    area2 = (x[T[:, 1]] - x[T[:, 0]]) * (y[T[:, 2]] - y[T[:, 0]]) - (x[T[:, 2]] - x[T[:, 0]]) * (y[T[:, 1]] - y[T[:, 0]])
    assert (area2 > 0).all()                              # counter-clockwise
    assert (np.diff(m.V[:, 0]) >= 0).all()                # x-sorted vertices
    assert (np.diff(m.TriGC[:, 0]) >= -1e-9).all()        # x-sorted triangles
    # TriC(ti,n) shares the edge opposite vertex n
    for ti in range(0, m.nTri, 97):
        for n in range(3):
            tj = m.TriC[ti, n]
            if tj == 0:
                continue
            e = {m.Tri[ti, (n + 1) % 3], m.Tri[ti, (n + 2) % 3]}
            assert e.issubset(set(m.Tri[tj - 1]))
    assert (m.TriBI > 0).sum() > 0 and set(np.unique(m.VBI)) == set(range(9))


def test_operator_exactness(oracle, mesh41):
    # ct_discretisation_mapping_derivatives.f90:533-575: map/ddx/ddy reproduce linear functions,
    # the 2nd-order b_b operators reproduce quadratics, to round-off
    m, ops = mesh41, mesh41.ops
    x, y, gx, gy = m.V[:, 0], m.V[:, 1], m.TriGC[:, 0], m.TriGC[:, 1]
    f, fb = 3 + 2e-5 * x - 1e-5 * y, 3 + 2e-5 * gx - 1e-5 * gy
    assert np.abs(oracle.spmv(ops["M_map_a_b"], f) - fb).max() < 1e-11
    assert np.abs(oracle.spmv(ops["M_ddx_a_b"], f) - 2e-5).max() < 1e-15
    assert np.abs(oracle.spmv(ops["M_ddy_a_b"], f) + 1e-5).max() < 1e-15
    assert np.abs(oracle.spmv(ops["M_map_b_a"], fb) - f).max() < 1e-11
    assert np.abs(oracle.spmv(ops["M_ddx_b_a"], fb) - 2e-5).max() < 1e-15
    q = 1 + 2e-5 * gx - 1e-5 * gy + 3e-10 * gx ** 2 - 2e-10 * gx * gy + 1e-10 * gy ** 2
    assert np.abs(oracle.spmv(ops["M2_d2dx2_b_b"], q) - 6e-10).max() < 1e-18
    assert np.abs(oracle.spmv(ops["M2_d2dxdy_b_b"], q) + 2e-10).max() < 1e-18
    assert np.abs(oracle.spmv(ops["M2_d2dy2_b_b"], q) - 2e-10).max() < 1e-18
    assert np.abs(oracle.spmv(ops["M2_ddx_b_b"], q) - (2e-5 + 6e-10 * gx - 2e-10 * gy)).max() < 1e-14
    # pattern conventions: a_b rows = the triangle's 3 vertices in Tri order; b_b diagonal first
    A = ops["M_map_a_b"]
    assert np.array_equal(A.ind.reshape(-1, 3), m.Tri)
    B = ops["M2_ddx_b_b"]
    assert np.array_equal(B.ind[B.ptr[:-1] - 1], np.arange(1, m.nTri + 1))
    for fam in (("M_map_a_b", "M_ddx_a_b", "M_ddy_a_b"), ("M_map_b_a", "M_ddx_b_a", "M_ddy_b_a")):
        assert all(ops[k].ind is ops[fam[0]].ind or np.array_equal(ops[k].ind, ops[fam[0]].ind) for k in fam)


def test_operator_row_ranges_are_rank_independent(oracle, mesh41):
    # rows built for a sub-range (one rank's partition_list range) equal the same rows of the full build
    m = mesh41
    i1, i2 = oracle.partition_list(m.nTri, 1, 3)
    part = oracle.calc_operator_rows(m, "b_b_2nd", i1, i2)[2]
    full = m.ops["M2_d2dx2_b_b"]
    k0, k1 = full.ptr[i1 - 1] - 1, full.ptr[i2] - 1
    assert np.array_equal(part.ind, full.ind[k0:k1]) and np.array_equal(part.val, full.val[k0:k1])
    assert part.ptr[0] == 1 and np.array_equal(np.diff(part.ptr), np.diff(full.ptr[i1 - 1:i2 + 1]))


def test_laplace_solve_on_b_grid(oracle, mesh41):
    # ct_discretisation_solve_Laplace_eq.f90:153-212
    m, ops = mesh41, mesh41.ops
    c, r0 = -1e-9, m.xmax * 0.8
    gx, gy = m.TriGC[:, 0], m.TriGC[:, 1]
    f_ex = -c / 4 * r0 ** 2 + c / 4 * (gx ** 2 + gy ** 2)
    xx, yy = ops["M2_d2dx2_b_b"], ops["M2_d2dy2_b_b"]
    rows = []
    bb = np.zeros(m.nTri)
    for ti in range(m.nTri):
        if np.hypot(gx[ti], gy[ti]) >= r0:
            rows.append([(ti + 1, 1.0)])
            bb[ti] = f_ex[ti]
        else:
            k0, k1 = xx.ptr[ti] - 1, xx.ptr[ti + 1] - 1
            rows.append([(int(xx.ind[k]), xx.val[k] + yy.val[k]) for k in range(k0, k1)])
            bb[ti] = c
    A = _csr(oracle, rows, m.nTri)
    f = oracle.direct_solve(A, bb)
    assert np.abs(f - f_ex).max() / np.abs(f_ex).max() < 1e-9     # quadratic solution is reproduced exactly
    xk, its, reason, _ = oracle.ksp_solve(A, bb, 1e-12, 1e-14, nranks=2)
    assert reason in (2, 3) and np.abs(xk - f_ex).max() / np.abs(f_ex).max() < 1e-8


def test_zeta_and_vertical_integrals(oracle):
    z = mesh_types.zeta_regular(12)
    assert z[0] == 0.0 and z[-1] == 1.0 and np.allclose(np.diff(z), 1 / 11)
    zl = mesh_types.zeta_irregular_log(12, 10.0)
    assert zl[0] == 0.0 and abs(zl[-1] - 1.0) < 1e-15 and (np.diff(zl) > 0).all()
    assert abs((zl[1] - zl[0]) / (zl[-1] - zl[-2]) - 10.0 ** (10 / 11)) < 1e-9   # constant spacing ratio
    F = np.stack([np.ones(12), z, z ** 2]).copy()
    avg = oracle.vertical_average(z, F)
    assert np.allclose(avg[:2], [1.0, 0.5], atol=1e-15) and abs(avg[2] - 1 / 3) < 2e-3   # trapezoid
    I = oracle.integrate_from_zeta_is_one_to_zeta_is_zetap(z, F)
    assert np.all(I[:, -1] == 0) and np.allclose(I[0], z - 1.0, atol=1e-15)
    assert np.allclose(I[1], 0.5 * (z ** 2 - 1.0), atol=1e-15)                  # trapezoid exact for linear


def test_schoof_2006_closed_form(oracle):
    # Schoof_SSA_solution.f90:36-61 with the SSA_icestream config values
    A, n, H, tt, L, m = 1e-18, 3.0, 2000.0, -3e-4, 150e3, 1.0
    y = np.array([0.0, 75e3, 150e3, 299e3, 301e3, -75e3])
    u, tau = oracle.Schoof2006_icestream(A, n, H, tt, L, m, y)
    f = 910.0 * 9.81 * 2000.0 * 3e-4
    assert np.allclose(tau, f * np.abs(y / L))
    assert u[4] == 0.0 and u[3] > 0 and u[0] > u[1] > u[2] > u[3] and u[1] == u[5]
    B = A ** (-1 / 3)
    u0 = -2 * f ** 3 * L ** 4 / (B ** 3 * H ** 3) * (-(2 ** 4) / 4 + 3 * 2 ** 5 / 10 - 3 * 2 ** 6 / 24 + 2 ** 7 / 56)
    assert abs(u[0] - u0) / u0 < 1e-12


def test_bc_copy_tables(oracle, mesh41):
    # find_ti_copy_ISMIP_HOM_periodic: weights normalised, copies lie near the displaced point
    import ctypes as ct
    m = mesh41
    L = 400e3
    cm = oracle._cmesh(m)
    tc, wc = np.zeros(m.nC_mem, np.int32), np.zeros(m.nC_mem)
    border = np.nonzero(m.TriBI > 0)[0]
    for ti in border[::7]:
        oracle.lib().ora_find_ti_copy(ct.byref(cm), 1, ct.c_double(L), int(ti) + 1, oracle._p(tc), oracle._p(wc))
        nz = tc > 0
        assert 3 <= nz.sum() <= 8 and abs(wc.sum() - 1.0) < 1e-14
        gx, gy = m.TriGC[ti]
        px = gx - L / 2 if gx > 0 else gx + L / 2
        py = gy - L / 2 if gy > 0 else gy + L / 2
        d = np.hypot(m.TriGC[tc[nz] - 1, 0] - px, m.TriGC[tc[nz] - 1, 1] - py)
        assert d.max() < 3 * 20e3


# ------------------------------------------------------------------------------------------
# committed fixtures (tests/golden/, generated by tests/golden/make_golden.py)
# ------------------------------------------------------------------------------------------
def _golden_dir():
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_against_committed_known_answers(oracle):
    import json, os
    K = json.load(open(os.path.join(_golden_dir(), "reference_known_answers.json")))
    for name, case in K.items():
        rows = [[(int(c), float(v)) for c, v in r] for r in case["rows"]]
        A = _csr(oracle, rows, case["n"])
        if "y" in case:
            assert np.array_equal(oracle.spmv(A, np.array(case["x"], float)), np.array(case["y"], float)), name
        else:
            x = oracle.direct_solve(A, np.array(case["b"], float))
            assert np.abs(x - np.array(case["x"])).max() < case["tol"], name


def test_oracle_reproduces_committed_lattice_fixture(oracle):
    """The npz was produced by the oracle; re-deriving it guards the fixture against silent drift
    of either the oracle or the synthetic-mesh generator."""
    import os
    from ufemism2_0_b200 import experiments
    G = np.load(os.path.join(_golden_dir(), "oracle_lattice_9x7.npz"))
    L = float(G["L"])
    mesh = synthetic.lattice_mesh(-L, L, -L, L, 9, 7, jitter=0.25, seed=424242)
    assert np.array_equal(mesh.V, G["V"]) and np.array_equal(mesh.Tri, G["Tri"])
    oracle.calc_all_matrix_operators_mesh(mesh)
    for nm, A in mesh.ops.items():
        assert np.array_equal(A.ptr, G[nm + "_ptr"]) and np.array_equal(A.ind, G[nm + "_ind"]), nm
        assert np.allclose(A.val, G[nm + "_val"], rtol=1e-12, atol=0), nm


def test_secondary_velocities_closed_forms(oracle, mesh41):
    """calc_secondary_velocities (conservation_of_momentum_main.f90:176-245) on a field with a known
    answer: u = (a + b x)(2 - zeta) is linear in x and zeta, so the b->a map and the trapezoidal
    vertical average are exact."""
    mesh = mesh41
    oracle.calc_all_matrix_operators_mesh(mesh)
    z = mesh.zeta
    gx = mesh.TriGC[:, 0]
    u3 = np.asfortranarray((3.0 + 1e-5 * gx)[:, None] * (2.0 - z)[None, :])
    v3 = np.asfortranarray(-0.5 * u3)
    o = oracle.calc_secondary_velocities(mesh, u3, v3)
    interior = mesh.VBI == 0
    xa = mesh.V[:, 0]
    assert np.allclose(o["u_surf"][interior], (3.0 + 1e-5 * xa[interior]) * 2.0, rtol=1e-12)
    assert np.allclose(o["u_base"][interior], (3.0 + 1e-5 * xa[interior]) * 1.0, rtol=1e-12)
    assert np.allclose(o["u_vav_b"], (3.0 + 1e-5 * gx) * 1.5, rtol=1e-13)
    assert np.allclose(o["uabs_vav_b"], np.sqrt(1.25) * np.abs(o["u_vav_b"]), rtol=1e-13)
    assert np.allclose(o["R_shear"], (o["uabs_base"] + 0.1) / (o["uabs_surf"] + 0.1))
    assert np.array_equal(o["u_3D"][:, 0], o["u_surf"])
