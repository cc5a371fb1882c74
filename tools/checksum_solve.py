"""sha256 of the velocity fields after a fixed number of Picard iterations (bit-identity checks between two builds:
UFE_LIB_PATH=<other libufe_diva.so> python tools/checksum_solve.py ...).
usage: python tools/checksum_solve.py [workload] [n_picard]"""
import copy, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "mismipplus_8km"
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh, C, ice, label = bench.make_workload(name)
from ufemism2_0_b200 import diva
C = copy.copy(C)
C.visc_it_nit = nit - 1
S = diva.initialise_DIVA_solver(mesh, C)
info = S.solve_DIVA(ice)
out = {"workload": name, "n_visc_its": info.n_visc_its, "L2_uv": info.L2_uv}
for f in ("u_vav_b", "v_vav_b", "u_3D_b", "v_3D_b", "eta_3D_b", "tau_bx_b"):
    out[f] = hashlib.sha256(getattr(S, f).tobytes()).hexdigest()[:16]
print(json.dumps(out))
S.close()
