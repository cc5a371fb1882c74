"""Exploration on a GPU box: sizes, iteration counts, SpMV GB/s (prints, no asserts)."""
import sys, time, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva

def run(name, mesh, C, ice, nit, maxits, methods=('bicgstab',), pcs=('jacobi','bjacobi2')):
    print('==', name, 'nV', mesh.nV, 'nTri', mesh.nTri, flush=True)
    t = time.time(); S = diva.initialise_DIVA_solver(mesh, C); print('create s', time.time() - t, flush=True)
    for meth in methods:
        for pc in pcs:
            C2 = copy.copy(C); C2.b200_krylov_method = meth; C2.b200_krylov_pc = pc
            C2.visc_it_nit = nit; C2.b200_krylov_maxits = maxits
            S.set_config(C2)
            for k in S.STATE_FIELDS_B: getattr(S, k)[:] = 0
            S.eta_3D_b[:] = 0
            t = time.time(); info = S.solve_DIVA(ice); print(meth, pc, 'wall', time.time() - t, info, flush=True)
            for fl in (False, True):
                ms, by = S.bench_spmv(20, fl); print('  spmv flush', fl, 'ms', ms, 'GB/s', by / ms / 1e6, flush=True)
    S.close()

which = sys.argv[1:] or ['mismip', 'ant']
if 'mismip' in which:
    mesh, C, ice = experiments.MISMIPplus(2e3)
    run('MISMIP+ 2km', mesh, C, ice, 50, 10000)
if 'm8' in which:
    mesh, C, ice = experiments.MISMIP_8km()
    run('MISMIP 8km', mesh, C, ice, 50, 10000)
if 'ant' in which:
    t = time.time(); mesh, C, ice = experiments.antarctic(1_000_000); print('mesh s', time.time() - t)
    run('antarctic 1M', mesh, C, ice, 2, 3000, pcs=('bjacobi2',))
