"""Full DIVA velocity solves on wide synthetic Antarctic-shaped meshes with krylov_pc = 'nd_lu' (multifrontal nested
dissection as the exact preconditioner), next to 'bjacobi2' on the same mesh with the same iteration caps.  Cold start.
usage: python tools/nd_diva_wide.py <max Picard its> <pc lag> nV ...      (one JSON line per mesh and preconditioner)"""
import copy, dataclasses, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
import numpy as np
from ufemism2_0_b200 import experiments, diva

nit, lag = int(sys.argv[1]), int(sys.argv[2])
pcs = os.environ.get("ND_PCS", "nd_lu,bjacobi2").split(",")
for nV in [int(float(a)) for a in sys.argv[3:]]:
    mesh, C0, ice = experiments.antarctic(nV)
    ref = None
    for pc in pcs:
        C = copy.copy(C0)
        C.visc_it_nit, C.b200_krylov_pc, C.b200_krylov_pc_lag = nit, pc, lag
        C.b200_krylov_maxits = 2000 if pc != "nd_lu" else 200
        S = diva.initialise_DIVA_solver(mesh, C)
        t = time.time(); info = S.solve_DIVA(ice); wall = time.time() - t
        u, v = np.array(S.u_vav_b), np.array(S.v_vav_b)
        out = {"nV": mesh.nV, "unknowns": 2 * mesh.nTri, "pc": pc, "pc_lag": lag, "wall_s": wall, **dataclasses.asdict(info),
               "picard_converged": bool(info.n_visc_its < nit), "max_speed": float(np.hypot(u, v).max())}
        if ref is None: ref = (u, v)
        else:
            out["rel_L2_vs_first_pc"] = float(np.sqrt(((u - ref[0]) ** 2 + (v - ref[1]) ** 2).sum() / max((ref[0] ** 2 + ref[1] ** 2).sum(), 1e-300)))
        print(json.dumps(out), flush=True)
        S.close()
