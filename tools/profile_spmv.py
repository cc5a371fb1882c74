"""Roofline probe on a GPU box: builds the synthetic Antarctic-scale mesh, assembles the
stiffness matrix with one (truncated) Picard iteration and times the Krylov MatMult kernel.
Used under ncu for profiles/ (see profiles/README.md)."""
import sys, time, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva

nV = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t = time.time(); mesh, C, ice = experiments.antarctic(nV); print('mesh s', time.time() - t, 'nV', mesh.nV, 'nTri', mesh.nTri, flush=True)
C.visc_it_nit = 0; C.b200_krylov_maxits = 40; C.b200_krylov_pc = "bjacobi2"   # the timed kernel is the MatMult; no factorisation wanted here
t = time.time(); S = diva.initialise_DIVA_solver(mesh, C); print('create s', time.time() - t, flush=True)
t = time.time(); info = S.solve_DIVA(ice, outputs=False); print('solve wall', time.time() - t, info, flush=True)
for fl in (False, True):
    ms, by = S.bench_spmv(reps, fl)
    print('spmv flush', fl, 'ms', ms, 'bytes', by, 'GB/s', by / ms / 1e6, flush=True)
S.close()
