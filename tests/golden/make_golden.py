"""Generates the committed golden fixtures of tests/golden/.

1. reference_known_answers.json -- the known-answer vectors the reference's own unit tests hold
   for this path, transcribed by hand from the cited files (the reference is Fortran + MPI + PETSc
   and cannot be executed in this image, so nothing here is produced by running it):
     src/UPSY/validation/unit_tests/ut_mpi_CSR_matrix_vector_multiplication.f90:188-321
     src/UPSY/validation/unit_tests/ut_petsc.f90:85-149
     src/UPSY/validation/unit_tests/ut_mpi_CSR_matrix_solving.f90:217-270
2. oracle_lattice_9x7.npz -- outputs of the oracle (oracle/, the CPU restatement pinned by (1) and
   by tests/test_oracle_golden.py) on a small seeded mesh: operator CSRs, one assembled stiffness
   matrix and the converged DIVA solution of an ISMIP-HOM-C-like set-up.  The GPU tests compare
   the CUDA path with these files directly, so parity is also checked against bytes that do not
   change when the oracle's code does.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

KNOWN = {
    "spmv_7x7_eq1": {
        "source": "ut_mpi_CSR_matrix_vector_multiplication.f90:188-250",
        "rows": [[[1, 1.0]], [[2, 2.0], [3, 3.0]], [[2, 4.0], [5, 1.0]], [[3, 2.0], [4, 3.0], [6, 4.0]],
                 [[4, 1.0], [5, 2.0]], [[5, 3.0], [7, 4.0]], [[6, 1.0], [7, 2.0]]],
        "n": 7, "x": [1, 2, 3, 4, 1, 2, 3], "y": [1, 13, 9, 26, 6, 15, 8]},
    "spmv_7x7_eq2": {
        "source": "ut_mpi_CSR_matrix_vector_multiplication.f90:252-321",
        "rows": [[[1, 1.0], [5, 5.0]], [[2, 2.0], [3, 3.0], [6, 6.0]], [[2, 4.0], [5, 1.0]],
                 [[1, 5.0], [3, 2.0], [4, 3.0], [6, 4.0], [7, 5.0]], [[2, 6.0], [4, 1.0], [5, 2.0]],
                 [[1, 5.0], [5, 3.0], [7, 4.0]], [[2, 6.0], [6, 1.0], [7, 2.0]]],
        "n": 7, "x": [1, 2, 3, 4, 1, 2, 3], "y": [6, 25, 9, 46, 18, 20, 20]},
    "petsc_matmult_6x6": {
        "source": "ut_petsc.f90:85-149 (two ranks: rows 1-2 | 3-6)",
        "rows": [[[1, 1.0]], [[1, 2.0], [2, 3.0]], [[2, 4.0], [4, 5.0]], [[4, 6.0], [5, 7.0]], [[5, 8.0]],
                 [[5, 9.0], [6, 10.0]]],
        "n": 6, "x": [1, 2, 3, 4, 5, 6], "y": [1, 8, 28, 59, 40, 105]},
    "tridiagonal_solve_7": {
        "source": "ut_mpi_CSR_matrix_solving.f90:217-270",
        "rows": [[[1, 1.0]]] + [[[i - 1, -1.0], [i, 2.0], [i + 1, -1.0]] for i in range(2, 7)] + [[[7, 1.0]]],
        "n": 7, "b": [1] * 7, "x": [1, 3.5, 5, 5.5, 5, 3.5, 1], "tol": 1e-5},
}


def main():
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as fh:
        json.dump(KNOWN, fh, indent=1)
    import ufe_pkg
    ufe_pkg.load()
    from ufemism2_0_b200 import experiments, synthetic
    import oracle as O
    O.build()
    L = 40e3
    mesh = synthetic.lattice_mesh(-L, L, -L, L, 9, 7, jitter=0.25, seed=424242)
    _, C, _ = experiments.ISMIP_HOM("C", L, 9)
    ice = synthetic.geometry_ISMIP_HOM_C(mesh, L)
    C.stress_balance_PETSc_rtol, C.stress_balance_PETSc_abstol = 1e-12, 1e-12
    O.calc_all_matrix_operators_mesh(mesh)
    out = {"V": mesh.V, "Tri": mesh.Tri, "VBI": mesh.VBI, "L": np.array(L)}
    for nm, A in mesh.ops.items():
        out[nm + "_ptr"], out[nm + "_ind"], out[nm + "_val"] = A.ptr, A.ind, A.val
    D = O.new_DIVA_state(mesh)
    nv, _ = O.solve_DIVA(mesh, ice, C, D, "direct")
    out["n_visc_its"] = np.array(nv)
    for k in ("u_vav_b", "v_vav_b", "u_3D_b", "eta_3D_a", "N_b", "beta_eff_b", "tau_dx_b"):
        out[k] = np.asarray(D[k])
    A, bb = O.assemble_stiffness(mesh, C, D["N_b"], D["dN_dx_b"], D["dN_dy_b"], D["beta_eff_b"], D["tau_dx_b"],
                                 D["tau_dy_b"], D["u_b_prev"], D["v_b_prev"])
    out["A_ptr"], out["A_ind"], out["A_val"], out["A_bb"] = A.ptr, A.ind, A.val, bb
    np.savez_compressed(os.path.join(HERE, "oracle_lattice_9x7.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
