"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): operator construction,
closures, assembly, both Krylov methods, all three preconditioners, secondary velocities."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva
mesh, C, ice = experiments.ISMIP_HOM("C", 80e3, 13)
C.visc_it_nit = 2
for meth in ("bicgstab", "gmres"):
    for pc in ("jacobi", "bjacobi2", "bjacobi_lu"):
        C2 = copy.copy(C); C2.b200_krylov_method, C2.b200_krylov_pc, C2.b200_krylov_maxits = meth, pc, 60
        S = diva.initialise_DIVA_solver(mesh, C2)
        info = S.solve_DIVA(ice)
        sec = S.calc_secondary_velocities()
        print(meth, pc, info.n_visc_its, info.n_Axb_its, info.flags, float(np.abs(S.u_vav_b).max()), flush=True)
        S.close()
mesh, C, ice = experiments.SSA_icestream(9, 21)
C.visc_it_nit = 2
S = diva.initialise_DIVA_solver(mesh, C); print('ssa', S.solve_SSA(ice).n_Axb_its); S.close()
print("SANITIZE_RUN_DONE")
