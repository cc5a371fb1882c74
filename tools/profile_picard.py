"""A few Picard iterations of a named workload on the device (for `ncu` launch lists and timing A/B runs).
usage: python tools/profile_picard.py [workload] [n_picard] [repeats]   -- prints one JSON line with the device-time split"""
import copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "mismipplus_2km"
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mesh, C, ice, label = bench.make_workload(name)
from ufemism2_0_b200 import diva
C = copy.copy(C)
C.visc_it_nit = nit - 1
S = diva.initialise_DIVA_solver(mesh, C)
S.upload(ice, state=True)
S.solve_DIVA_resident()              # analysis, graph capture
out = []
for _ in range(reps):
    S.reset_state_resident()
    t0 = time.perf_counter()
    i = S.solve_DIVA_resident()
    out.append({"wall_ms": 1e3 * (time.perf_counter() - t0), "ms_total": i.ms_total, "ms_closures": i.ms_closures, "ms_assembly": i.ms_assembly,
                "ms_krylov": i.ms_krylov, "n_visc_its": i.n_visc_its, "n_Axb_its": i.n_Axb_its, "launches": i.gpu_launches})
print(json.dumps({"workload": name, "runs": out}))
S.close()
