"""First-light diagnostics on a GPU box (prints, no asserts)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva
import oracle as O

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300), np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

# L0: tridiagonal known answer
n=7; rows=[]; 
ptr=[1]; ind=[]; val=[]
for i in range(1,n+1):
    if i in (1,n): ind+= [i]; val+=[1.0]
    else: ind += [i-1,i,i+1]; val += [-1.0,2.0,-1.0]
    ptr.append(len(ind)+1)
A = diva.CSRMatrix(n,n,1,n,np.array(ptr,np.int32),np.array(ind,np.int32),np.array(val))
for meth in ('bicgstab','gmres'):
    x,its,fl = diva.solve_matrix_equation_CSR(A, np.ones(n), np.zeros(n), 1e-10, 1e-12, method=meth)
    print('L0', meth, x, its, fl)
print('spmv', diva.multiply_CSR_matrix_with_vector(A, np.arange(1.0,8.0)))

for exp in ('A','C'):
    mesh, C, ice = experiments.ISMIP_HOM(exp, 160e3, 41)
    t=time.time(); S = diva.initialise_DIVA_solver(mesh, C); print('create', time.time()-t)
    O.calc_all_matrix_operators_mesh(mesh)
    for fam, names in (('a_b',('map','ddx','ddy')),('b_a',('map','ddx','ddy')),('b_b',('ddx','ddy','d2dx2','d2dxdy','d2dy2'))):
        for w in names:
            G = S.get_operator(fam, w)
            nm = ('M2_%s_b_b'%w) if fam=='b_b' else ('M_%s_%s'%(w,fam))
            R = mesh.ops[nm]
            same = np.array_equal(G.ptr,R.ptr) and np.array_equal(G.ind,R.ind)
            print(nm, 'pattern equal', same, 'val rel', rel(G.val,R.val) if same else None, 'bit-equal', np.array_equal(G.val,R.val) if same else None)
    for meth in ('bicgstab','gmres'):
        C.b200_krylov_method = meth
        C.stress_balance_PETSc_rtol=1e-10; C.stress_balance_PETSc_abstol=1e-9
        S.set_config(C)
        for k in ('u_vav_b','v_vav_b','tau_bx_b','tau_by_b','u_base_b','v_base_b'): getattr(S,k)[:] = 0
        S.eta_3D_b[:] = 0
        t=time.time(); info = S.solve_DIVA(ice); print(exp, meth, 'gpu solve', time.time()-t, info)
    D = O.new_DIVA_state(mesh); 
    t=time.time(); nv,na = O.solve_DIVA(mesh, ice, C, D, 'direct'); print('oracle', time.time()-t, nv)
    print('u_vav', rel(S.u_vav_b, D['u_vav_b']), 'v_vav', rel(S.v_vav_b, D['v_vav_b']), 'u3D', rel(S.u_3D_b, D['u_3D_b']))
    Ag, bg = S.get_stiffness_matrix()
    print('stiffness nnz', Ag.ptr[-1]-1)
    ms, by = S.bench_spmv(20); print('spmv ms', ms, 'GB/s', by/ms/1e6)
    S.close()
