"""torchrun --nproc-per-node N tests/tools/multi_gpu_check.py : N-rank solves (strip partition by
partition_list, halo exchange + allreduce over NCCL) compared on rank 0 with the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import torch, torch.distributed as dist
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva, capi, mesh_types

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
def make_comm():
    """A fresh NCCL unique id per handle (an id can initialise one communicator only)."""
    uid = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0:
        buf = (capi.ct.c_char * 128)()
        capi.check(capi.lib().ufe_comm_get_unique_id(buf))
        uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device='cuda')
    dist.broadcast(uid, 0)
    return (rank, world, local, bytes(uid.cpu().tolist()))

def rel(a, b, ref):
    return np.linalg.norm(a - b) / np.linalg.norm(ref), np.abs(a - b).max() / np.abs(ref).max()

ok = True
# the cases below are small: keep their rows partitioned (the default would solve them redundantly on every rank)
os.environ['UFE_REDUNDANT_MAX_UNKNOWNS'] = '0'
cases = []
for pc_name, meth in (('auto', 'bicgstab'), ('bjacobi2', 'bicgstab'), ('bjacobi2', 'gmres'), ('auto', 'gmres')):
    cases.append(('ISMIP-HOM A', experiments.ISMIP_HOM('A', 160e3, 41), pc_name, meth))
cases.append(('ISMIP-HOM C', experiments.ISMIP_HOM('C', 160e3, 31), 'auto', 'bicgstab'))
cases.append(('MISMIP+ 8km', experiments.MISMIPplus(8e3), 'auto', 'bicgstab'))
cases.append(('MISMIP+ 8km strip-only', experiments.MISMIPplus(8e3), 'bjacobi_lu', 'bicgstab'))
# the multifrontal solver with its elimination sub-trees distributed over the ranks (explicitly, and on a wide mesh)
cases.append(('MISMIP+ 8km nd', experiments.MISMIPplus(8e3), 'nd_lu', 'gmres'))
cases.append(('Antarctic 2e4', experiments.antarctic(20000), 'nd_lu', 'bicgstab'))
if os.environ.get('UFE_CHECK_ONLY') == 'nd':       # short run: only the distributed multifrontal cases
    cases = [c for c in cases if c[2] == 'nd_lu']
oracle_cache = {}
for name, (mesh, C, ice), pc_name, meth in cases:
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
    C.b200_krylov_pc, C.b200_krylov_method = pc_name, meth
    C.b200_krylov_pc_strip_only = name.endswith('strip-only')
    if name.startswith('MISMIP'): C.visc_it_nit = 8
    if name.startswith('Antarctic'): C.visc_it_nit = 4
    S = diva.initialise_DIVA_solver(mesh, C, make_comm())
    t = time.time(); info = S.solve_DIVA(ice); wall = time.time() - t
    sec = S.calc_secondary_velocities()
    own = S.ownership()
    if rank == 0:
        import oracle as O
        O.build()
        key = name.split(' strip')[0].split(' nd')[0]
        if key not in oracle_cache:
            O.calc_all_matrix_operators_mesh(mesh)
            D = O.new_DIVA_state(mesh); nv, _ = O.solve_DIVA(mesh, ice, C, D, 'direct')
            oracle_cache[key] = (D, nv, mesh)
        D, nv, mesh0 = oracle_cache[key]
        ref = np.concatenate([D['u_vav_b'], D['v_vav_b']])
        ru, rv = rel(S.u_vav_b, D['u_vav_b'], ref), rel(S.v_vav_b, D['v_vav_b'], ref)
        r3 = np.abs(S.u_3D_b - D['u_3D_b']).max() / np.abs(D['u_3D_b']).max()
        want = O.calc_secondary_velocities(mesh0, S.u_3D_b, S.v_3D_b)
        rs = max(np.abs(sec[k] - w).max() / max(np.abs(w).max(), 1e-300) for k, w in want.items())
        good = max(ru + rv) < 1e-6 and abs(info.n_visc_its - nv) <= 1 and r3 < 1e-6 and rs < 1e-12
        ok &= good
        if pc_name == 'nd_lu' or (pc_name == 'auto' and world in (1, 2, 4, 8)): good = good and info.krylov_pc_used == 4
        print(f'{name} [{meth}+{pc_name} -> pc {info.krylov_pc_used}]: ranks {world} comm {"peer" if info.reserved else "nccl"} Picard {info.n_visc_its} (oracle {nv}) '
              f'Krylov {info.n_Axb_its} flags {info.flags} u {ru[1]:.2e} v {rv[1]:.2e} u3D {r3:.2e} secondary {rs:.1e} wall {wall:.3f}s '
              f'{"OK" if good else "MISMATCH"}', flush=True)
    thk = None
    if name == 'MISMIP+ 8km':      # the thickness update that follows the solve: replicated on every rank
        E = mesh_types.calc_mesh_edges(mesh)
        S.set_mesh_edges(E)
        n = mesh.nV
        f = dict(Hi=ice.Hi, Hb=ice.Hb, SL=ice.SL, SMB=np.full(n, 0.3), BMB=np.zeros(n), LMB=np.zeros(n), fraction_margin=np.ones(n),
                 mask_noice=np.zeros(n, dtype=np.int32), dHi_dt_target=np.zeros(n))
        S.solve_DIVA_resident()    # leaves the velocities distributed: the thickness call has to gather them itself
        thk = S.calc_dHi_dt(f, 1.0)
        t_all = [torch.zeros(n, dtype=torch.float64, device='cuda') for _ in range(world)]
        dist.all_gather(t_all, torch.from_numpy(thk['Hi_tplusdt']).cuda())
        same = all(torch.equal(t_all[0], t) for t in t_all)
        S.download()
        if rank == 0:
            Ed = {k: getattr(E, k) for k in ('nE', 'VE', 'ETri', 'A', 'Cw', 'D_x', 'D_y', 'D')}
            want = O.calc_dHi_dt(mesh, Ed, C, dict(f, u_vav_b=S.u_vav_b, v_vav_b=S.v_vav_b), 1.0)
            rt = np.abs(thk['Hi_tplusdt'] - want['Hi_tplusdt']).max() / np.abs(want['Hi_tplusdt']).max()
            good_t = same and rt < 1e-6
            ok &= good_t
            print(f'{name} thickness update (calc_dHi_dt, replicated on {world} ranks): identical on all ranks {same}, Hi_tplusdt vs oracle {rt:.2e} '
                  f'Krylov {thk["n_Axb_its"]} {"OK" if good_t else "MISMATCH"}', flush=True)
    S.close()
if os.environ.get('UFE_CHECK_ONLY') == 'nd':
    dist.barrier()
    if rank == 0: print('MULTI_GPU_CHECK', 'PASS' if ok else 'FAIL', '(nd cases only)', flush=True)
    dist.destroy_process_group()
    sys.exit(0)
# L0 as the reference calls it (solve_matrix_equation_CSR_PETSc, petsc_basic.f90:32-64): every rank passes its row block
import scipy.sparse as sp, scipy.sparse.linalg as spla
def gather_rows(local, n_total, i1):
    """all ranks' slices -> the full vector on every rank (slices tile 1..n in rank order)"""
    sizes = [torch.zeros(1, dtype=torch.int64, device='cuda') for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.size], dtype=torch.int64, device='cuda'))
    parts = [torch.zeros(int(s.item()), dtype=torch.float64, device='cuda') for s in sizes]
    full = torch.zeros(n_total, dtype=torch.float64, device='cuda')
    full[i1 - 1:i1 - 1 + local.size] = torch.from_numpy(np.ascontiguousarray(local)).cuda()
    dist.all_reduce(full)
    return full.cpu().numpy()
mesh, C, ice = experiments.MISMIPplus(8e3)
C.visc_it_nit = 3
C.b200_krylov_pc = 'auto'
S = diva.initialise_DIVA_solver(mesh, C, make_comm())
S.solve_DIVA(ice)
A, bb = S.get_stiffness_matrix()                       # this rank's rows, exactly what the reference holds
loc = sp.csr_matrix((A.val, A.ind - 1, A.ptr - 1), shape=(A.i2 - A.i1 + 1, A.n))
x, its, fl, used = S.solve_matrix_equation_CSR(A, bb, np.zeros_like(bb), 1e-12, 1e-11)
xf, bf = gather_rows(x, A.m, A.i1), gather_rows(bb, A.m, A.i1)
# residual of the rank's own rows against the gathered solution; the exact solution through a gathered matrix on rank 0
r_loc = np.abs(loc @ xf - bb).max() / np.abs(bf).max()
r_all = torch.tensor([r_loc], dtype=torch.float64, device='cuda'); dist.all_reduce(r_all, op=dist.ReduceOp.MAX)
good = fl == 0 and its <= 3 and r_all.item() < 1e-10 and (used == 4 if world in (1, 2, 4, 8) else used in (0, 2))
ok &= good
if rank == 0:
    print(f'L0 solve_matrix_equation_CSR, DIVA stiffness rows of {world} ranks [auto -> pc {used}]: Krylov {its} flags {fl} residual {r_all.item():.2e} '
          f'{"OK" if good else "MISMATCH"}', flush=True)
# a generic system (not the stiffness matrix): banded, diagonally dominant, row blocks by partition_list
n = 6000
i1, i2 = diva.partition_list(n, rank, world)
rng = np.random.default_rng(11)
rows_all = []
for i in range(n):
    cols = sorted(set(int(c) for c in np.clip(i + rng.integers(-30, 31, 6), 0, n - 1)) - {i})
    rows_all.append([(i + 1, 8.0 + float(rng.random()))] + [(c + 1, float(rng.standard_normal())) for c in cols])
b_all = rng.standard_normal(n)
ptr, ind, val = [1], [], []
for r in rows_all[i1 - 1:i2]:
    for c, v in r: ind.append(c); val.append(v)
    ptr.append(len(ind) + 1)
Ag = diva.CSRMatrix(n, n, i1, i2, np.array(ptr, np.int32), np.array(ind, np.int32), np.array(val))
for pc_name in ('jacobi', 'auto'):
    C2 = __import__('copy').deepcopy(C); C2.b200_krylov_pc = pc_name
    S.set_config(C2)
    x, its, fl, used = S.solve_matrix_equation_CSR(Ag, b_all[i1 - 1:i2], np.zeros(i2 - i1 + 1), 1e-13, 1e-14)
    xf = gather_rows(x, n, i1)
    if rank == 0:
        full = sp.lil_matrix((n, n))
        for i, r in enumerate(rows_all):
            for c, v in r: full[i, c - 1] = v
        want = spla.splu(full.tocsc()).solve(b_all)
        e = np.abs(xf - want).max() / np.abs(want).max()
        good = fl == 0 and e < 1e-10
        ok &= good
        print(f'L0 solve_matrix_equation_CSR, generic banded system n={n} on {world} ranks [{pc_name} -> pc {used}]: Krylov {its} flags {fl} '
              f'x vs sparse LU {e:.2e} {"OK" if good else "MISMATCH"}', flush=True)
S.close()
# the default for small systems: not partitioned, every rank solves the whole system, bit-identical results
del os.environ['UFE_REDUNDANT_MAX_UNKNOWNS']
mesh, C, ice = experiments.MISMIPplus(8e3)
C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-11
C.visc_it_nit = 8
S = diva.initialise_DIVA_solver(mesh, C, make_comm())
info = S.solve_DIVA(ice)
u_all = [torch.zeros(mesh.nTri, dtype=torch.float64, device='cuda') for _ in range(world)]
dist.all_gather(u_all, torch.from_numpy(S.u_vav_b).cuda())
same = all(torch.equal(u_all[0], t) for t in u_all)
if rank == 0:
    D, nv, _ = oracle_cache['MISMIP+ 8km']
    ref = np.concatenate([D['u_vav_b'], D['v_vav_b']])
    ru = rel(S.u_vav_b, D['u_vav_b'], ref)
    good = same and info.reserved == 2 and max(ru) < 1e-6 and S.ownership() == (1, mesh.nV, 1, mesh.nTri)
    ok &= good
    print(f'MISMIP+ 8km redundant default: ranks {world} mode {info.reserved} identical on all ranks {same} Picard {info.n_visc_its} (oracle {nv}) '
          f'u {ru[1]:.2e} {"OK" if good else "MISMATCH"}', flush=True)
S.close()
dist.barrier()
if rank == 0:
    print('MULTI_GPU_CHECK', 'PASS' if ok else 'FAIL', flush=True)
dist.destroy_process_group()
