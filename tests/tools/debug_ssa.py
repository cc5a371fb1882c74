import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva
import oracle as O
O.build()
mesh, C, ice = experiments.SSA_icestream(15, 61)
O.calc_all_matrix_operators_mesh(mesh)
C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-11, 1e-10
C.visc_it_nit = 60
R = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri)); tr = []
nv, _ = O.solve_SSA(mesh, ice, C, R, 'direct', trace=tr)
print('oracle', nv, [(i, float('%.3g' % l)) for i, l, _ in tr][:8], 'umax', np.abs(R['u_b']).max())
for meth in ('bicgstab', 'gmres'):
    for pc in ('jacobi', 'bjacobi2'):
        C2 = copy.copy(C); C2.b200_krylov_method = meth; C2.b200_krylov_pc = pc
        S = diva.initialise_DIVA_solver(mesh, C2)
        info = S.solve_SSA(ice)
        print(meth, pc, info.n_visc_its, info.n_Axb_its, info.flags, info.L2_uv, 'umax', np.abs(S.u_b).max(),
              'rel', np.linalg.norm(S.u_b - R['u_b']) / np.linalg.norm(R['u_b']))
        # one linearised solve from the oracle's first-iteration inputs
        S.close()
# first linearised system: compare matrices
C1 = copy.copy(C); C1.visc_it_nit = 0
S = diva.initialise_DIVA_solver(mesh, C1)
info = S.solve_SSA(ice)
R1 = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri))
O.solve_SSA(mesh, ice, C1, R1, 'direct')
A, bb = S.get_stiffness_matrix()
Ao, bo = O.assemble_stiffness(mesh, C1, R1['N_b'], R1['dN_dx_b'], R1['dN_dy_b'], R1['basal_friction_coefficient_b'], R1['tau_dx_b'], R1['tau_dy_b'], np.zeros(mesh.nTri), np.zeros(mesh.nTri))
print('pattern eq', np.array_equal(A.ptr, Ao.ptr), np.array_equal(A.ind, Ao.ind), 'val', np.abs(A.val - Ao.val).max() / np.abs(Ao.val).max(), 'bb', np.abs(bb - bo).max() / np.abs(bo).max())
print('first it: gpu its', info.n_Axb_its, info.flags, 'u rel', np.linalg.norm(S.u_b - R1['u_b']) / np.linalg.norm(R1['u_b']))
S.close()
