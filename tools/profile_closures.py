"""Profile probe for the non-Krylov kernels on a GPU box: one truncated Picard iteration (closures, assembly)
and one thickness update on the synthetic Antarctic-scale mesh.  Used under ncu for profiles/."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva, mesh_types

nV = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
t = time.time(); mesh, C, ice = experiments.antarctic(nV); print('mesh s', time.time() - t, 'nV', mesh.nV, 'nTri', mesh.nTri, flush=True)
C.visc_it_nit = 1; C.b200_krylov_maxits = 4; C.b200_krylov_pc = "bjacobi2"
S = diva.initialise_DIVA_solver(mesh, C)
info = S.solve_DIVA(ice, outputs=False)
print('solve', info.n_visc_its, info.ms_closures, info.ms_assembly, info.ms_krylov, flush=True)
S.set_mesh_edges(mesh_types.calc_mesh_edges(mesh))
n = mesh.nV
f = dict(Hi=ice.Hi, Hb=ice.Hb, SL=ice.SL, SMB=np.full(n, 0.3), BMB=np.zeros(n), LMB=np.zeros(n), fraction_margin=np.ones(n),
         mask_noice=np.zeros(n, dtype=np.int32), dHi_dt_target=np.zeros(n))
for _ in range(2):
    out = S.calc_dHi_dt_semiimplicit(f, 1.0)
print('thickness', out["n_Axb_its"], S.thickness_timing(), flush=True)
S.close()
