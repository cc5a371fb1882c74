"""Small run for compute-sanitizer (memcheck / racecheck) over the thickness path: edge upload, explicit and
semi-implicit schemes with every BC / mask branch, calc_dHi_dt, a_a operators, vertical velocities."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import config, diva, experiments, mesh_types
import test_thickness as T

mesh, C, ice = experiments.MISMIPplus(16e3)
C.visc_it_nit = 2
E = mesh_types.calc_mesh_edges(mesh)
S = diva.initialise_DIVA_solver(mesh, C)
S.solve_DIVA(ice)
S.set_mesh_edges(E)
f = T._random_case(mesh, E, seed=4)
for bc in ("zero", "infinite"):
    C2 = copy.copy(C); C2.BC_H_west = C2.BC_H_east = C2.BC_H_north = C2.BC_H_south = bc
    for m in ("explicit", "semi-implicit", "none"):
        C2.choice_ice_integration_method = m
        S.C = C2
        r = S.calc_dHi_dt(f, 2.0)
        print(bc, m, r["dt"], r["n_Axb_its"], r["flags"], float(np.abs(r["Hi_tplusdt"]).max()), flush=True)
g = {k: v for k, v in f.items() if k not in ("u_vav_b", "v_vav_b", "BC_prescr_mask", "BC_prescr_Hi")}
print("resident", S.calc_dHi_dt_semiimplicit(g, 1.0)["n_Axb_its"], flush=True)
S.calc_secondary_velocities()
nV, nz = mesh.nV, mesh.nz
vin = dict(Hi=ice.Hi, Hib=ice.Hib, dHb_dt=np.zeros(nV), dHi_dt=np.ones(nV), BMB=-np.ones(nV), mask_grounded_ice=ice.mask_grounded_ice,
           mask_floating_ice=ice.mask_floating_ice, dzeta_dx_ak=np.zeros((nV, nz), order="F"), dzeta_dy_ak=np.zeros((nV, nz), order="F"),
           dzeta_dz_ak=np.asfortranarray(np.repeat((-1.0 / np.maximum(0.1, ice.Hi))[:, None], nz, axis=1)))
w = S.calc_vertical_velocities(vin)
print("w", float(np.abs(w).max()), S.get_operator_a_a("ddx").val.size, flush=True)
S.close()
print("SANITIZE_RUN_DONE")
