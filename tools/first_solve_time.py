"""Time to the first solution on a named workload: handle creation (operators, pattern), the once-per-mesh symbolic analysis of the
multifrontal solver (host), graph capture, and one truncated solve; then the same call again (everything cached).
usage: python tools/first_solve_time.py [workload]"""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "antarctic_1m"
mesh, C, ice, label = bench.make_workload(name)
from ufemism2_0_b200 import diva
C = copy.copy(C)
C.visc_it_nit = 1
t = time.time(); S = diva.initialise_DIVA_solver(mesh, C); t_create = time.time() - t
t = time.time(); info = S.solve_DIVA(ice, outputs=False); t1 = time.time() - t
t = time.time(); info2 = S.solve_DIVA(ice, outputs=False); t2 = time.time() - t
print(f"{name}: create {t_create:.2f} s, first solve ({info.n_visc_its} Picard its) incl. analysis {t1:.2f} s, the same again {t2:.2f} s, pc {info.krylov_pc_used}")
S.close()
