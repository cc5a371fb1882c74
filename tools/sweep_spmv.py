"""BASELINE.json configs[4]: Krylov / SpMV sweep on synthetic Antarctic-shaped meshes of growing size.
Per size: one truncated Picard iteration assembles the stiffness matrix, then the Krylov MatMult kernel is timed with
CUDA events (L2 flushed between launches) and a capped BiCGStab run gives the time of a whole Krylov iteration.
Prints one JSON line per size (single GPU)."""
import copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments, diva

sizes = [int(float(a)) for a in sys.argv[1:]] or [10_000, 30_000, 100_000, 300_000, 1_000_000, 3_000_000]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6500.0)
for nV in sizes:
    t = time.time(); mesh, C, ice = experiments.antarctic(nV); t_mesh = time.time() - t
    C = copy.copy(C)
    C.visc_it_nit, C.b200_krylov_maxits, C.b200_krylov_pc = 0, 200, "bjacobi2"
    S = diva.initialise_DIVA_solver(mesh, C)
    info = S.solve_DIVA(ice, outputs=False)
    ms, nbytes = S.bench_spmv(50, flush_l2=True)
    ms_hot, _ = S.bench_spmv(50, flush_l2=False)
    n = 2 * mesh.nTri
    it_ms = info.ms_krylov / max(info.n_Axb_its, 1)
    it_bytes = 2.0 * nbytes + 16 * 8.0 * n
    print(json.dumps({"nV": mesh.nV, "nTri": mesh.nTri, "unknowns": n, "spmv_ms_l2_flushed": ms, "spmv_GBs": nbytes / ms / 1e6,
                      "spmv_frac_of_peak": nbytes / ms / 1e6 / peak, "spmv_ms_l2_warm": ms_hot, "spmv_GBs_l2_warm": nbytes / ms_hot / 1e6,
                      "bicgstab_its_timed": info.n_Axb_its, "ms_per_krylov_iteration": it_ms, "krylov_iteration_GBs": it_bytes / it_ms / 1e6,
                      "krylov_iteration_frac_of_peak": it_bytes / it_ms / 1e6 / peak, "ms_closures": info.ms_closures,
                      "ms_assembly": info.ms_assembly, "host_mesh_s": t_mesh}), flush=True)
    S.close()
    del S, mesh, ice
