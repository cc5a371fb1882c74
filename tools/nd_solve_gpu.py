"""Multifrontal nested-dissection solver on the device (ufe_nd_solver_*): factor / solve timings on the stiffness
system of synthetic Antarctic-shaped meshes of growing size (the 'wide' meshes the banded exact preconditioner cannot
hold).  One truncated Picard iteration on the device assembles the system; prints one JSON line per size.
usage: python tools/nd_solve_gpu.py [leaf_triangles] nV ..."""
import copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
import numpy as np
from ufemism2_0_b200 import experiments, diva, nd

leaf = int(sys.argv[1]) if len(sys.argv) > 1 else 96
sizes = [int(float(a)) for a in sys.argv[2:]] or [10_000, 30_000, 100_000]
for nV in sizes:
    mesh, C, ice = experiments.antarctic(nV)
    C = copy.copy(C)
    C.visc_it_nit, C.b200_krylov_maxits, C.b200_krylov_pc = 0, 5, "bjacobi2"
    S = diva.initialise_DIVA_solver(mesh, C)
    S.solve_DIVA(ice, outputs=False)
    A, bb = S.get_stiffness_matrix()
    S.close()
    N = 2 * mesh.nTri
    t = time.time(); sol = nd.Solver(np.asarray(mesh.TriGC), A.ptr, A.ind, leaf); t_sym = time.time() - t
    sol.factor(A.val); sol.factor(A.val)
    f_ms = sol.info()["factor_ms"]
    out = {"nV": mesh.nV, "unknowns": N, "leaf_triangles": leaf, "symbolic_host_s": t_sym}
    for k in (0, 1, 2):
        x, rr = sol.solve(bb, n_refine=k)
        out[f"relres_refine{k}"] = rr; out[f"solve_ms_refine{k}"] = sol.info()["solve_ms"]
    i = sol.info()
    out.update({"factor_ms": f_ms, "factor_TFLOPs": i["factor_flops"] / f_ms / 1e9, "front_GB": i["front_bytes"] / 1e9,
                "n_fronts": i["n_fronts"], "n_levels": i["n_levels"], "max_front": i["max_front"]})
    print(json.dumps(out), flush=True)
    sol.close()
