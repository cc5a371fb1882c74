"""bjacobi_lu preconditioner on a GPU box: convergence and parity (prints)."""
import sys, os, copy, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva
import oracle as O
O.build()
which = sys.argv[1:] or ['ssa', 'm8', 'm4', 'm2']

def rel(a, b, ref): return np.linalg.norm(a - b) / np.linalg.norm(ref), np.abs(a - b).max() / np.abs(ref).max()

if 'ssa' in which:
    mesh, C, ice = experiments.SSA_icestream(15, 61)
    O.calc_all_matrix_operators_mesh(mesh)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-11, 1e-10
    C.visc_it_nit = int(os.environ.get('SSA_NIT', '8'))
    R = dict(u_b=np.zeros(mesh.nTri), v_b=np.zeros(mesh.nTri))
    nv, _ = O.solve_SSA(mesh, ice, C, R, 'direct')
    for seg in (0,):
        for meth in ('bicgstab',):
            C2 = copy.copy(C); C2.b200_krylov_pc = 'bjacobi_lu'; C2.b200_krylov_pc_lag = seg; C2.b200_krylov_method = meth
            S = diva.initialise_DIVA_solver(mesh, C2)
            t = time.time(); info = S.solve_SSA(ice); w = time.time() - t
            ref = np.concatenate([R['u_b'], R['v_b']])
            print('SSA seg', seg, meth, 'picard', info.n_visc_its, '(oracle %d)' % nv, 'krylov', info.n_Axb_its, 'flags', info.flags,
                  'rel', rel(S.u_b, R['u_b'], ref), 'wall %.3f' % w, 'launches', info.gpu_launches, flush=True)
            S.close()
for tag, h, nit in (('m8', 8e3, 8), ('m4', 4e3, 6), ('m2', 2e3, 50)):
    if tag not in which: continue
    mesh, C, ice = experiments.MISMIPplus(h)
    C.visc_it_nit = int(os.environ.get('M_NIT', nit))
    D = None
    if tag != 'm2':
        O.calc_all_matrix_operators_mesh(mesh)
        C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-11, 1e-10
        D = O.new_DIVA_state(mesh); nv, _ = O.solve_DIVA(mesh, ice, C, D, 'direct')
    for seg in (0,):
        C2 = copy.copy(C); C2.b200_krylov_pc = 'bjacobi_lu'; C2.b200_krylov_pc_lag = seg
        S = diva.initialise_DIVA_solver(mesh, C2)
        t = time.time(); info = S.solve_DIVA(ice); w = time.time() - t
        msg = ''
        if D is not None:
            ref = np.concatenate([D['u_vav_b'], D['v_vav_b']])
            msg = 'rel u %s v %s (oracle picard %d)' % (rel(S.u_vav_b, D['u_vav_b'], ref), rel(S.v_vav_b, D['v_vav_b'], ref), nv)
        print('MISMIP+', h, 'nTri', mesh.nTri, 'seg', seg, 'picard', info.n_visc_its, 'krylov', info.n_Axb_its, 'flags', info.flags,
              'ms total %.1f asm %.1f krylov %.1f' % (info.ms_total, info.ms_assembly, info.ms_krylov), 'wall %.3f' % w,
              'launches', info.gpu_launches, msg, flush=True)
        S.close()
