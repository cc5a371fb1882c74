"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes as ct
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from ufemism2_0_b200 import capi, config, diva, experiments

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ufe_diva.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ufe_[A-Za-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(capi.EXPORTS)
    assert lib.ufe_version() == 200


def test_partition_list_matches_oracle(oracle):
    for ntot in (0, 3, 16, 17, 1000, 2000003):
        for n in (1, 2, 4, 8):
            for i in range(n):
                assert diva.partition_list(ntot, i, n) == oracle.partition_list(ntot, i, n)


def test_struct_layouts_match_header():
    # compile a tiny C program against the header and compare sizeof() with the ctypes mirrors
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "ufe_diva.h"
        int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ufe_csr), sizeof(ufe_mesh),
          sizeof(ufe_config), sizeof(ufe_ice_inputs), sizeof(ufe_diva_state), sizeof(ufe_ssa_state),
          sizeof(ufe_solve_info), sizeof(ufe_comm), sizeof(ufe_mesh_edges), sizeof(ufe_thickness_config),
          sizeof(ufe_thickness_fields)); printf("%zu\\n", sizeof(ufe_vertical_velocity_inputs)); return 0; }""")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    mirrors = [capi.ufe_csr, capi.ufe_mesh, capi.ufe_config, capi.ufe_ice_inputs, capi.ufe_diva_state,
               capi.ufe_ssa_state, capi.ufe_solve_info, capi.ufe_comm, capi.ufe_mesh_edges, capi.ufe_thickness_config,
               capi.ufe_thickness_fields, capi.ufe_vertical_velocity_inputs]
    assert sizes == [ct.sizeof(m) for m in mirrors]


def test_config_namelist_parser(tmp_path):
    p = tmp_path / "config.cfg"
    p.write_text(textwrap.dedent("""
        &CONFIG
          ! comment
          start_time_of_run_config                     = 0.0      ! ignored key
          visc_it_norm_dUV_tol_config                  = 5E-7                             ! Stop criterion
          visc_it_nit_config                           = 5000
          visc_it_relax_config                         = 0.4_dp
          stress_balance_PETSc_rtol_config             = 1E-6
          BC_u_west_config                             = 'periodic_ISMIP-HOM'   ! Boundary conditions
          choice_sliding_law_config                    = 'no_sliding'
          do_GL_subgrid_friction_config                = .FALSE.
          uniform_Glens_flow_factor_config             = 1.4280330398280316E-017
          choice_initial_velocity_ANT_config           = 'zero'
        /
        """))
    C = config.Config.from_namelist(str(p))
    assert C.visc_it_norm_dUV_tol == 5e-7 and C.visc_it_nit == 5000 and C.visc_it_relax == 0.4
    assert C.BC_u_west == "periodic_ISMIP-HOM" and C.BC_u_east == "infinite"
    assert C.choice_sliding_law == "no_sliding" and C.do_GL_subgrid_friction is False
    assert C.uniform_Glens_flow_factor == 1.4280330398280316e-17
    # defaults of model_configuration.f90:307-313
    D = config.Config()
    assert (D.visc_it_norm_dUV_tol, D.visc_it_nit, D.visc_it_relax, D.stress_balance_PETSc_rtol,
            D.stress_balance_PETSc_abstol) == (5e-5, 50, 0.2, 1e-7, 1e-5)


def test_unknown_choices_raise_like_crash():
    C = config.Config(choice_sliding_law="Coulombic")
    with pytest.raises(capi.UfeError, match="unknown choice_sliding_law"):
        diva.config_struct(C)
    with pytest.raises(capi.UfeError, match="unknown choice_BC_u"):
        diva.config_struct(config.Config(BC_u_west="open"))
    with pytest.raises(capi.UfeError, match="unknown choice_flow_law"):
        diva.config_struct(config.Config(choice_flow_law="Newton"))


def test_no_cpu_fallback_without_gpu():
    """Without a usable sm_100 device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 9)
    with pytest.raises(capi.UfeError):
        diva.initialise_DIVA_solver(mesh, C)
    A = diva.CSRMatrix(2, 2, 1, 2, np.array([1, 2, 3], np.int32), np.array([1, 2], np.int32), np.ones(2))
    with pytest.raises(capi.UfeError):
        diva.multiply_CSR_matrix_with_vector(A, np.ones(2))
    with pytest.raises(capi.UfeError):
        diva.solve_matrix_equation_CSR(A, np.ones(2), np.zeros(2), 1e-8, 1e-10)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ufemism2.0_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "ufe_oracle" not in txt and "oracle/" not in txt, f
    # tools/ is oracle-free as well (the checker scripts that use the oracle live in tests/tools/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "tools", f), errors="replace").read()
            assert "import oracle" not in txt and "ufe_oracle" not in txt and "'oracle'" not in txt and '"oracle"' not in txt, f


GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["UFE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["UFE_ROOT"], "oracle"))
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import synthetic, diva
import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the unique-id broadcast bench.py performs before ufe_diva_create
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0: uid[:] = torch.arange(128, dtype=torch.uint8)
dist.broadcast(uid, 0)
assert uid.tolist() == list(range(128))
mesh = synthetic.lattice_mesh(-1e5, 1e5, -1e5, 1e5, 17, 13)
ti1, ti2 = diva.partition_list(mesh.nTri, rank, world)
vi1, vi2 = diva.partition_list(mesh.nV, rank, world)
# each rank builds and applies only its own rows (strip decomposition), then all-gathers
rows = O.calc_operator_rows(mesh, "a_b", ti1, ti2)
f = 1.0 + 3e-5 * mesh.V[:, 0] - 2e-5 * mesh.V[:, 1]
y_loc = O.spmv(rows[0], f)
sizes = [diva.partition_list(mesh.nTri, r, world) for r in range(world)]
parts = [torch.zeros(b - a + 1, dtype=torch.float64) for a, b in sizes]
dist.all_gather(parts, torch.from_numpy(y_loc))
y = torch.cat(parts).numpy()
full = O.calc_operator_rows(mesh, "a_b", 1, mesh.nTri)[0]
assert np.array_equal(y, O.spmv(full, f))
# halo range = column range touched by the owned rows (calc_j_node_range)
lo, hi = rows[0].ind.min(), rows[0].ind.max()
assert lo <= vi1 + 0 or vi1 > vi2
t = torch.tensor([float(lo), float(hi)], dtype=torch.float64)
allr = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allr, t)
if rank == 0:
    assert allr[0][0] == 1 and allr[-1][1] == mesh.nV
# Krylov-style dot product: local partial sums + all_reduce == global
s = torch.tensor([float(np.dot(y_loc, y_loc))], dtype=torch.float64)
dist.all_reduce(s)
assert abs(s.item() - float(np.dot(y, y))) <= 1e-9 * abs(s.item())
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_strip_decomposition(tmp_path, oracle):
    w = tmp_path / "worker.py"
    w.write_text(GLOO_WORKER)
    env = dict(os.environ, UFE_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(w)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU restatement of the reference path) prints one JSON line
    with the keys the driver reads; small workload so that it runs in seconds."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ismip_hom_a",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    # the whole solve, run to completion and converged: no extrapolation, no linear solve at the iteration cap
    cb = line["cpu_baseline"]
    assert cb["extrapolated"] is False and cb["picard_converged"] is True and cb["krylov_its_per_solve_max"] < 10000
    assert line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


REF = "/root/reference"
PATH_KEYS = [f.name for f in __import__("dataclasses").fields(config.Config) if not f.name.startswith("b200_")]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("cfg_file, make", [
    ("config-files/benchmarks/MISMIP+/config_MISMIPplus_2km_spinup.cfg", lambda: experiments.MISMIPplus(8e3)[1]),
    ("config-files/config_MISMIP_8km_spinup_for_scaling.cfg", lambda: experiments.MISMIP_8km(64e3)[1]),
])
def test_workload_configs_equal_the_reference_cfg_files(cfg_file, make):
    """The named workloads hard-code the solver keys of the reference's own .cfg files (the files are
    not available on the GPU box); here, where they are, every key the path reads must agree."""
    want = config.Config.from_namelist(os.path.join(REF, cfg_file))
    got = make()
    skip = {"choice_initial_velocity", "nz", "choice_zeta_grid", "zeta_irregular_log_R",
            "refgeo_idealised_SSA_icestream_Hi", "refgeo_idealised_SSA_icestream_dhdx", "refgeo_idealised_SSA_icestream_L",
            "refgeo_idealised_SSA_icestream_m", "refgeo_idealised_ISMIP_HOM_L", "choice_idealised_sliding_law"}
    diffs = {k: (getattr(got, k), getattr(want, k)) for k in PATH_KEYS if k not in skip and getattr(got, k) != getattr(want, k)}
    assert not diffs, diffs


def test_nested_dissection_prototype_solves_the_stiffness_system():
    """tests/tools/nd_prototype.py (round-2 planning prototype of the multifrontal solver): exact solve on a small MISMIP+ mesh."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "nd_prototype.py"), "mismipplus:16000", "24"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"residual ([0-9.e+-]+) \| x vs SuperLU ([0-9.e+-]+)", out.stdout)
    assert m and float(m.group(1)) < 1e-12 and float(m.group(2)) < 1e-7, out.stdout
    assert " level  fronts" in out.stdout


def _block_pattern(A, nT):
    """0-based block CSR over triangles of a scipy matrix with the 2x2 (u,v) unknown ordering."""
    import scipy.sparse as sp
    P = sp.csr_matrix((np.ones(A.nnz), A.indices // 2, A.indptr), shape=(2 * nT, nT))
    R = sp.csr_matrix((np.ones(2 * nT), (np.arange(2 * nT) // 2, np.arange(2 * nT))), shape=(nT, 2 * nT))
    B = (R @ P).tocsr()
    B.sort_indices()
    return B.indptr.astype(np.int32), B.indices.astype(np.int32)


@pytest.mark.parametrize("workload, leaf", [("mismipplus", 24), ("ismip_hom", 16)])
def test_nd_symbolic_analysis_drives_an_exact_multifrontal_solve(oracle, workload, leaf):
    """The staged host-side nested-dissection analysis (ufe_nd_analyse): a numpy multifrontal factorisation that uses
    ONLY its maps (assembly map of every block entry, extend-add positions, post-order) solves the oracle's stiffness
    system exactly; the tree invariants hold."""
    import scipy.sparse.linalg as spla
    from ufemism2_0_b200 import nd
    mesh, C, ice = experiments.MISMIPplus(16e3) if workload == "mismipplus" else experiments.ISMIP_HOM("C", 80e3, 21)
    mesh.ops = oracle.calc_all_matrix_operators_mesh(mesh)
    cap = {}
    orig = oracle.direct_solve
    def grab(A, b):
        cap.setdefault("A", A.to_scipy().tocsr()); cap.setdefault("b", np.asarray(b).copy())
        return orig(A, b)
    oracle.direct_solve = grab
    try:
        C.visc_it_nit = 0
        oracle.solve_DIVA(mesh, ice, C, oracle.new_DIVA_state(mesh))
    finally:
        oracle.direct_solve = orig
    A, b, nT = cap["A"], cap["b"], mesh.nTri
    A.sort_indices()
    bptr, bind = _block_pattern(A, nT)
    T = nd.analyse(np.asarray(mesh.TriGC), bptr, bind, leaf)
    # invariants: every triangle eliminated exactly once; children precede parents; root has no boundary
    allsep = np.concatenate([n.sep for n in T.nodes])
    assert np.array_equal(np.sort(allsep), np.arange(nT))
    assert all(n.parent > i for i, n in enumerate(T.nodes[:-1])) and T.nodes[-1].parent == -1 and T.nodes[-1].bnd.size == 0
    assert T.max_front == max(2 * (n.sep.size + n.bnd.size) for n in T.nodes) and T.n_levels == 1 + max(n.level for n in T.nodes)
    # numeric phase from the maps alone
    F = [np.zeros((2 * (n.sep.size + n.bnd.size),) * 2) for n in T.nodes]
    Ac = A.tocoo()
    # block entry index of every scalar entry: position of (i//2, j//2) in the block CSR
    rowb, colb = Ac.row // 2, Ac.col // 2
    e = np.array([bptr[r] + np.searchsorted(bind[bptr[r]:bptr[r + 1]], c) for r, c in zip(rowb, colb)])
    for v, r, c, k in zip(Ac.data, Ac.row, Ac.col, e):
        F[T.entry_node[k]][2 * T.entry_row[k] + r % 2, 2 * T.entry_col[k] + c % 2] += v
    dof = lambda t: np.stack([2 * t, 2 * t + 1], axis=1).ravel()
    X, F12, F21 = {}, {}, {}
    for i, n in enumerate(T.nodes):
        ns = 2 * n.sep.size
        X[i] = np.linalg.inv(F[i][:ns, :ns])
        F12[i], F21[i] = F[i][:ns, ns:], F[i][ns:, :ns]
        if n.parent >= 0:
            S = F[i][ns:, ns:] - F21[i] @ (X[i] @ F12[i])
            m = dof(n.up.astype(np.int64))
            F[n.parent][np.ix_(m, m)] += S
    x, z = b.copy(), {}
    for i, n in enumerate(T.nodes):
        z[i] = X[i] @ x[dof(n.sep)]
        x[dof(n.bnd)] -= F21[i] @ z[i]
    for i in range(len(T.nodes) - 1, -1, -1):
        n = T.nodes[i]
        x[dof(n.sep)] = z[i] - X[i] @ (F12[i] @ x[dof(n.bnd)])
    xr = spla.splu(A.tocsc()).solve(b)
    assert np.linalg.norm(A @ x - b) <= 1e-12 * np.linalg.norm(b)
    assert np.abs(x - xr).max() <= 1e-7 * np.abs(xr).max()


def test_nd_solver_fails_loudly_without_a_device():
    """The numeric multifrontal phase has no CPU fallback: on a machine without a GPU ufe_nd_solver_create returns
    UFE_ERR_CUDA with a message; a pattern that was not analysed is rejected before anything is allocated."""
    import torch
    from ufemism2_0_b200 import nd
    from ufemism2_0_b200.capi import UfeError
    nT = 64
    gc = np.stack([np.arange(nT, dtype=float), np.zeros(nT)], axis=1)
    rows = np.repeat(np.arange(2 * nT), 2)
    cols = np.stack([np.arange(2 * nT), np.arange(2 * nT) ^ 1], axis=1).ravel()      # 2x2 diagonal blocks only
    ptr = np.arange(0, 4 * nT + 1, 2, dtype=np.int32)
    if not torch.cuda.is_available():
        with pytest.raises(UfeError) as e:
            nd.Solver(gc, ptr + 1, cols.astype(np.int32) + 1, 8)      # 1-based, like type_sparse_matrix_CSR_dp
        assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    bptr, bind = nd.block_pattern(ptr, cols, nT)
    assert np.array_equal(bptr, np.arange(nT + 1)) and np.array_equal(bind, np.arange(nT))
    assert rows.size == cols.size


def _multifrontal_solve_from_maps(T, A, b, bptr, bind):
    """numpy numeric phase driven only by the maps of the symbolic analysis (same organisation as the device code:
    assembly map, partial elimination of sep, extend-add through ``up``, forward / backward sweeps over the post-order)."""
    A = A.tocsr(); A.sort_indices()
    Ac = A.tocoo()
    F = [np.zeros((2 * (n.sep.size + n.bnd.size),) * 2) for n in T.nodes]
    e = np.array([bptr[r] + np.searchsorted(bind[bptr[r]:bptr[r + 1]], c) for r, c in zip(Ac.row // 2, Ac.col // 2)], dtype=np.int64)
    for v, r, c, k in zip(Ac.data, Ac.row, Ac.col, e):
        F[T.entry_node[k]][2 * T.entry_row[k] + r % 2, 2 * T.entry_col[k] + c % 2] += v
    dof = lambda t: np.stack([2 * t, 2 * t + 1], axis=1).ravel().astype(np.int64)
    X, F12, F21 = {}, {}, {}
    for i, n in enumerate(T.nodes):
        ns = 2 * n.sep.size
        X[i] = np.linalg.inv(F[i][:ns, :ns])
        F12[i], F21[i] = F[i][:ns, ns:], F[i][ns:, :ns]
        if n.parent >= 0:
            m = dof(n.up.astype(np.int64))
            F[n.parent][np.ix_(m, m)] += F[i][ns:, ns:] - F21[i] @ (X[i] @ F12[i])
    x, z = b.astype(float).copy(), {}
    for i, n in enumerate(T.nodes):
        z[i] = X[i] @ x[dof(n.sep)]
        x[dof(n.bnd)] -= F21[i] @ z[i]
    for i in range(len(T.nodes) - 1, -1, -1):
        n = T.nodes[i]
        x[dof(n.sep)] = z[i] - X[i] @ (F12[i] @ x[dof(n.bnd)])
    return x


@pytest.mark.parametrize("case", ["single_triangle", "one_leaf", "two_components", "long_range_couplings", "unsymmetric_pattern",
                                  "coincident_centroids"])
def test_nd_symbolic_analysis_edge_cases(case):
    """Patterns that are not meshes: a 1-triangle system, a tree that is a single leaf, a disconnected graph, random
    far couplings (what periodic BC rows produce), a structurally unsymmetric pattern (BC rows reference neighbours that
    do not reference them back) and coincident centroids (degenerate bisection) -- every one must give maps that solve
    the system exactly."""
    import scipy.sparse as sp
    from ufemism2_0_b200 import nd
    rng = np.random.default_rng(11)
    nT, leaf = {"single_triangle": (1, 4), "one_leaf": (7, 8), "two_components": (60, 6), "long_range_couplings": (150, 8),
                "unsymmetric_pattern": (120, 8), "coincident_centroids": (40, 4)}[case]
    gc = rng.uniform(0, 1, (nT, 2))
    if case == "coincident_centroids":
        gc[:] = 0.5
    # block pattern: k nearest neighbours in the plane (+ extras), always with the diagonal
    d = ((gc[:, None, :] - gc[None, :, :]) ** 2).sum(-1)
    k = min(nT, 6)
    nb = np.argsort(d, axis=1, kind="stable")[:, :k]
    rows, cols = np.repeat(np.arange(nT), k), nb.ravel()
    if case == "two_components":
        half = nT // 2
        keep = (rows < half) == (cols < half)
        rows, cols = rows[keep], cols[keep]
    if case == "long_range_couplings":
        extra = rng.integers(0, nT, (30, 2))
        rows, cols = np.concatenate([rows, extra[:, 0]]), np.concatenate([cols, extra[:, 1]])
    if case != "unsymmetric_pattern":
        rows, cols = np.concatenate([rows, cols]), np.concatenate([cols, rows])
    rows, cols = np.concatenate([rows, np.arange(nT)]), np.concatenate([cols, np.arange(nT)])
    B = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(nT, nT)); B.sum_duplicates(); B.sort_indices()
    bptr, bind = B.indptr.astype(np.int32), B.indices.astype(np.int32)
    # scalar matrix: random 2x2 blocks on the pattern, diagonally dominant
    Bc = B.tocoo()
    r2 = (2 * Bc.row[:, None] + np.array([0, 0, 1, 1])[None, :]).ravel()
    c2 = (2 * Bc.col[:, None] + np.array([0, 1, 0, 1])[None, :]).ravel()
    A = sp.csr_matrix((rng.standard_normal(r2.size), (r2, c2)), shape=(2 * nT, 2 * nT))
    A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsr()
    b = rng.standard_normal(2 * nT)
    T = nd.analyse(gc, bptr, bind, leaf)
    assert np.array_equal(np.sort(np.concatenate([n.sep for n in T.nodes])), np.arange(nT))
    assert T.nodes[-1].parent == -1 and T.nodes[-1].bnd.size == 0
    if case in ("single_triangle", "one_leaf", "coincident_centroids"):
        assert len(T.nodes) == 1 and T.n_levels == 1
    # the scalar pattern of A has exactly the analysed block pattern
    b2ptr, b2ind = nd.block_pattern(A.indptr, A.indices, nT)
    assert np.array_equal(b2ptr, bptr) and np.array_equal(b2ind, bind)
    x = _multifrontal_solve_from_maps(T, A, b, bptr, bind)
    assert np.linalg.norm(A @ x - b) <= 1e-12 * np.linalg.norm(b)


def test_nd_analyse_rejects_bad_arguments():
    from ufemism2_0_b200 import nd
    from ufemism2_0_b200.capi import UfeError
    gc = np.zeros((3, 2))
    with pytest.raises(UfeError):
        nd.analyse(gc, np.array([0, 1, 2, 3], dtype=np.int32), np.array([0, 1, 7], dtype=np.int32), 4)     # column out of range
    with pytest.raises(UfeError):
        nd.analyse(gc, np.array([0, 1, 2, 3], dtype=np.int32), np.array([0, 1, 2], dtype=np.int32), 0)     # leaf size < 1


def test_nd_tree_distribution_over_ranks():
    """Host logic of the multi-rank multifrontal solver (ufe_nd_tree_owners): rank r owns the sub-tree below the r-th node
    of level log2(P) and the left spine above it, every rank gets about 1/P of the triangles, and every parent-child pair
    on different ranks is a (left spine, child 1) pair -- the only place a Schur complement crosses ranks."""
    import scipy.sparse as sp
    from ufemism2_0_b200 import nd
    mesh, C, ice = experiments.MISMIPplus(4e3)
    nT = mesh.nTri
    Tri = np.asarray(mesh.Tri) - 1
    Pm = sp.csr_matrix((np.ones(3 * nT), (np.repeat(np.arange(nT), 3), Tri.ravel())), shape=(nT, mesh.nV))
    B = (Pm @ Pm.T).tocsr(); B.sort_indices()
    T = nd.analyse(np.asarray(mesh.TriGC), B.indptr.astype(np.int32), B.indices.astype(np.int32), 48, nranks=(1, 2, 4, 8))
    nn = len(T.nodes)
    children = [[] for _ in range(nn)]
    for i, q in enumerate(T.nodes):
        if q.parent >= 0:
            children[q.parent].append(i)
    assert np.all(T.owners[1] == 0)
    for P in (2, 4, 8):
        own = T.owners[P]
        Ld = P.bit_length() - 1
        assert own.min() == 0 and own.max() == P - 1
        assert own[nn - 1] == 0                                     # the root (last in post-order) is rank 0's
        cut = [i for i, q in enumerate(T.nodes) if q.level == Ld]
        assert sorted(own[cut]) == list(range(P))                    # one sub-tree root per rank at level log2(P)
        crossings = 0
        for i, q in enumerate(T.nodes):
            if q.level > Ld:
                assert own[i] == own[q.parent]                       # below the cut a sub-tree stays on its rank
            elif q.parent >= 0 and own[i] != own[q.parent]:
                crossings += 1
                assert children[q.parent].index(i) == 1 and own[i] > own[q.parent]
        assert crossings == P - 1                                    # log2(P) exchange rounds, P - 1 links in all
        tri_per_rank = np.zeros(P)
        for i, q in enumerate(T.nodes):
            tri_per_rank[own[i]] += q.sep.size
        assert tri_per_rank.sum() == nT
        assert tri_per_rank.max() < 1.35 * nT / P, tri_per_rank      # coordinate bisection balances the sub-trees
    with pytest.raises(capi.UfeError, match="power of two"):
        nd.analyse(np.asarray(mesh.TriGC), B.indptr.astype(np.int32), B.indices.astype(np.int32), 48, nranks=(3,))


def test_nd_analysis_is_independent_of_the_host_thread_count(tmp_path):
    """The once-per-mesh symbolic analysis runs the two halves of the top cuts on different host threads and appends their
    sub-trees in post-order: the tree, the extend-add maps and the entry map must be exactly the single-threaded ones."""
    script = tmp_path / "nd_sum.py"
    script.write_text(textwrap.dedent("""
        import hashlib, sys
        sys.path.insert(0, %r)
        import ufe_pkg; ufe_pkg.load()
        import numpy as np, scipy.sparse as sp
        from ufemism2_0_b200 import experiments, nd
        mesh, C, ice = experiments.antarctic(30000)
        nT = mesh.nTri
        Tri = np.asarray(mesh.Tri) - 1
        P = sp.csr_matrix((np.ones(3 * nT), (np.repeat(np.arange(nT), 3), Tri.ravel())), shape=(nT, mesh.nV))
        B = (P @ P.T).tocsr(); B.sort_indices()
        T = nd.analyse(np.asarray(mesh.TriGC), B.indptr.astype(np.int32), B.indices.astype(np.int32), 32, nranks=(4,))
        h = hashlib.sha256()
        for q in T.nodes:
            h.update(np.int32([q.level, q.parent]).tobytes()); h.update(q.sep.tobytes()); h.update(q.bnd.tobytes()); h.update(q.up.tobytes())
        for a in (T.entry_node, T.entry_row, T.entry_col, T.owners[4]):
            h.update(np.ascontiguousarray(a).tobytes())
        print(len(T.nodes), T.n_levels, T.max_front, h.hexdigest())
        """ % ROOT))
    out = []
    for threads in ("1", "8"):
        env = dict(os.environ, UFE_ND_HOST_THREADS=threads)
        r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(r.stdout.strip().splitlines()[-1])
    assert out[0] == out[1], out


def test_pivot_block_inversion_scheme():
    """numpy restatement of mf_inv32_cta (csrc/ufe_nd_numeric.cu): in-place Gauss-Jordan where column p is replaced by the
    inverse's column of the pivot row, implicit threshold pivoting as a row permutation, lazy row scaling, the columns
    processed in groups of four (the owner group runs its four steps first, the other groups apply them afterwards), and
    the final scatter Ip[pinv[i]][perm[q]] = a_i[q] s_i.  Checks the scheme itself on well- and badly-pivoted blocks."""
    rng = np.random.default_rng(7)

    def invert(A, thresh=0.05):
        n = A.shape[0]
        W = A.copy()
        perm = list(range(n))
        s = np.ones(n)
        for g in range(n // 4):
            cols = list(range(4 * g, 4 * g + 4))
            steps = []
            for t, p in enumerate(cols):                    # the owner of the four columns, on its own columns only
                c = W[:, p].copy()
                if abs(c[perm[p]]) < thresh:
                    cand = [(abs(c[perm[k]]), -k) for k in range(p, n)]
                    k = -max(cand)[1]
                    perm[p], perm[k] = perm[k], perm[p]
                P = perm[p]
                piv = c[P]
                if abs(piv) < 1e-13:
                    piv = -1e-13 if piv < 0 else 1e-13
                d = 1.0 / piv
                m = c * d
                m[P] = 0.0
                s[P] = d
                steps.append((m, P))
                for q in cols:
                    W[:, q] -= m * W[P, q]
                W[:, p] = -m
                W[P, p] = 1.0
            for m, P in steps:                              # every other group applies the four steps to its columns
                for q in range(n):
                    if q not in cols:
                        W[:, q] -= m * W[P, q]
        pinv = [0] * n
        for p in range(n):
            pinv[perm[p]] = p
        out = np.zeros_like(A)
        for i in range(n):
            for q in range(n):
                out[pinv[i], perm[q]] = W[i, q] * s[i]
        return out

    for case in range(6):
        A = rng.standard_normal((32, 32)) * 0.3
        if case < 2:
            A += np.eye(32)                                  # Jacobi-scaled fronts: unit diagonal, no pivot search
        if case == 5:
            A[:, 3] *= 1e-3; A[7, 7] = 0.0                   # small pivots force the search
        X = invert(A)
        assert np.abs(X @ A - np.eye(32)).max() < 1e-9 * max(1.0, np.abs(X).max()), case
