import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva
import oracle as O
O.build()
mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 21)
C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-10, 1e-9
O.calc_all_matrix_operators_mesh(mesh)
C2 = copy.copy(C); C2.b200_krylov_pc = 'bjacobi_lu'
S = diva.initialise_DIVA_solver(mesh, C2)
for nit in (22, 26, 30, 34, 38, 42, 46, 50):
    C1 = copy.copy(C2); C1.visc_it_nit = nit
    D = O.new_DIVA_state(mesh); tr = []
    nv, _ = O.solve_DIVA(mesh, ice, C1, D, "direct", trace=tr)
    S.set_config(C1)
    for k in S.STATE_FIELDS_B: getattr(S, k)[:] = 0
    S.eta_3D_b[:] = 0
    info = S.solve_DIVA(ice)
    ref = np.abs(D['u_vav_b']).max()
    e = np.abs(S.u_vav_b - D['u_vav_b'])
    print('nit', nit, 'picard', info.n_visc_its, nv, 'relax', info.visc_it_relax_applied, 'eps', info.Glens_flow_law_epsilon_sq_0_applied, 'L2', info.L2_uv, tr[-1][1],
          'max err u', e.max() / ref, 'at tri', e.argmax(), 'TriBI', mesh.TriBI[e.argmax()], 'eta err', np.abs(S.eta_3D_a - D['eta_3D_a']).max() / np.abs(D['eta_3D_a']).max(), flush=True)
print([ (i, float('%.4g' % l)) for i, l, _ in tr][20:])
