"""Small run of the multifrontal path for compute-sanitizer (memcheck / racecheck / synccheck): nd_lu through auto on a
narrow and a wide mesh (register Gauss-Jordan panels, trailing updates, extend-add, cluster and plain sweeps, Richardson
step), BiCGStab / GMRES on an aged factorisation, the gather-record closures, the L0 entry point on the handle."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva

os.environ.setdefault("UFE_ND_GRAPHS", "0")          # the sanitizer follows plain launches
for name, (mesh, C, ice), nit in (("MISMIP+ 8km", experiments.MISMIPplus(8e3), 2), ("Antarctic 4e3", experiments.antarctic(4000), 2)):
    for meth, lag in (("bicgstab", 0), ("gmres", 20)):
        C2 = copy.copy(C)
        C2.visc_it_nit, C2.b200_krylov_method, C2.b200_krylov_pc, C2.b200_krylov_pc_lag = nit, meth, "auto", lag
        S = diva.initialise_DIVA_solver(mesh, C2)
        info = S.solve_DIVA(ice)
        A, bb = S.get_stiffness_matrix()
        x, its, fl, used = S.solve_matrix_equation_CSR(A, bb, np.zeros_like(bb), 1e-10, 1e-9)
        print(name, meth, lag, info.n_visc_its, info.n_Axb_its, info.flags, info.krylov_pc_used, "L0", its, fl, used,
              float(np.abs(S.u_vav_b).max()), flush=True)
        S.close()
# the large-front kernels on a small mesh: late Schur pass from 64 pivots, 128-wide tiles (SIMT and tensor cores) for every level
os.environ["UFE_ND_SCHUR_MIN_P"] = "64"
os.environ["UFE_ND_BIG_MIN_CTAS"] = "1"
for mma, min_mode in (("1", "0"), ("0", "2")):
    os.environ["UFE_ND_UPD_MMA"], os.environ["UFE_ND_UPD_MMA_MIN_MODE"] = mma, min_mode
    mesh, C, ice = experiments.antarctic(6000)
    C2 = copy.copy(C)
    C2.visc_it_nit, C2.b200_krylov_pc = 2, "nd_lu"
    S = diva.initialise_DIVA_solver(mesh, C2)
    info = S.solve_DIVA(ice)
    print("Antarctic 6e3 large-front kernels, tensor cores", mma, info.n_visc_its, info.n_Axb_its, info.flags, info.krylov_pc_used,
          float(np.abs(S.u_vav_b).max()), flush=True)
    S.close()
print("SANITIZE_RUN_DONE")
