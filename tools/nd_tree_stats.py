"""CPU-only: device-layout statistics of the nested-dissection tree (what ufe_nd_solver_create will allocate and what
the Gauss-Jordan sweeps will cost) for the stiffness pattern of a synthetic mesh, per leaf size.
usage: python tools/nd_tree_stats.py antarctic:<nV> | mismipplus:<h>  [leaf ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ufe_pkg; ufe_pkg.load()
import numpy as np
import scipy.sparse as sp
from ufemism2_0_b200 import experiments, nd


def layout(T):
    nl = T.n_levels
    n = np.zeros(nl, int); p = np.zeros(nl, int); nb = np.zeros(nl, int)
    for q in T.nodes:
        n[q.level] += 1; p[q.level] = max(p[q.level], 2 * q.sep.size); nb[q.level] = max(nb[q.level], 2 * q.bnd.size)
    p = np.maximum(32, (p + 31) // 32 * 32); g = (p + nb + 63) // 64 * 64
    by = 8.0 * n * g * g
    return {"levels": nl, "fronts": int(n.sum()), "GB": by.sum() / 1e9, "leaf_level_GB": by[-1] / 1e9, "flops": float((2.0 * n * g * g * p).sum()),
            "sweep_traffic_GB": float((2 * by * (p / 32)).sum() / 1e9), "launches": int((2 * (p // 32) + 3).sum()), "max_g": int(g.max())}


if __name__ == "__main__":
    kind, val = sys.argv[1].split(":")
    mesh, C, ice = experiments.MISMIPplus(float(val)) if kind == "mismipplus" else experiments.antarctic(int(float(val)))
    nT, nV = mesh.nTri, mesh.nV
    # stand-in for the stiffness block pattern (no device, no oracle here): triangles that share a vertex
    # (13 per row on a regular triangulation; the M2_*_b_b stencils of the reference have about 10)
    Tri = np.asarray(mesh.Tri) - 1
    P = sp.csr_matrix((np.ones(3 * nT), (np.repeat(np.arange(nT), 3), Tri.ravel())), shape=(nT, nV))
    B = (P @ P.T).tocsr(); B.sort_indices()
    bptr, bind = B.indptr.astype(np.int32), B.indices.astype(np.int32)
    print(f"nV {nV} nTri {nT} unknowns {2 * nT} block nnz/row {B.nnz / nT:.1f}")
    for leaf in [int(a) for a in sys.argv[2:]] or [24, 48, 96, 192]:
        t = time.time(); T = nd.analyse(np.asarray(mesh.TriGC), bptr, bind, leaf); dt = time.time() - t
        L = layout(T)
        print(f"leaf {leaf:4d}: " + "  ".join(f"{k} {v:.3g}" if isinstance(v, float) else f"{k} {v}" for k, v in L.items()) + f"  (analyse + copy {dt:.1f}s)", flush=True)
