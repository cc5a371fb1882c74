"""Round-2 planning probe (CPU, scipy): fill and flops of an exact LU of the DIVA stiffness matrix under a geometric
nested-dissection ordering of the triangles (recursive coordinate bisection on the centroids, separators = the
triangles whose 2-ring stencil crosses the cut), for the synthetic Antarctic-shaped mesh at growing sizes.
Prints nnz(L+U), the implied bytes, and the flop count from the column counts, next to the banded (x-sorted)
ordering that bjacobi_lu uses today."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments
import oracle as O


def nd_order(xy, adj, leaf=200):
    """Recursive bisection: returns a permutation (elimination order) of the nodes; separators last."""
    n = xy.shape[0]
    order = []

    def rec(idx):
        if idx.size <= leaf:
            order.extend(idx.tolist())
            return
        ext = xy[idx].max(axis=0) - xy[idx].min(axis=0)
        ax = int(np.argmax(ext))
        med = np.median(xy[idx, ax])
        left = idx[xy[idx, ax] <= med]
        right = idx[xy[idx, ax] > med]
        inleft = np.zeros(n, dtype=bool); inleft[left] = True
        # separator: nodes of `left` adjacent (in the matrix graph) to a node of `right`
        sub = adj[right]
        touched = np.unique(sub.indices)
        sep = touched[inleft[touched]]
        insep = np.zeros(n, dtype=bool); insep[sep] = True
        rec(left[~insep[left]])
        rec(right)
        order.extend(sep.tolist())

    rec(np.arange(n))
    return np.array(order)


for arg in sys.argv[1:] or ["10000", "40000", "mismipplus:2000"]:
    if arg.startswith("mismipplus:"):
        mesh, C, ice = experiments.MISMIPplus(float(arg.split(":")[1]))
    else:
        mesh, C, ice = experiments.antarctic(int(float(arg)))
    mesh.ops = O.calc_all_matrix_operators_mesh(mesh)
    nT = mesh.nTri
    z = np.zeros(nT)
    A, b = O.assemble_stiffness(mesh, C, np.ones(nT), z, z, np.ones(nT), z, z, z, z)
    M = A.to_scipy().tocsr()
    M.data[:] = 1.0
    M = (M + M.T).tocsr()                                   # structural symmetrisation
    M = M + sp.identity(M.shape[0]) * 100.0                 # any non-singular values: only the pattern matters here
    # block graph on triangles (2x2 u,v blocks)
    B = sp.csr_matrix((np.ones(M.nnz), M.indices // 2, M.indptr), shape=(M.shape[0], nT))
    G = (sp.csr_matrix((np.ones(2 * nT), (np.arange(2 * nT) // 2, np.arange(2 * nT))), shape=(nT, 2 * nT)) @ B).tocsr()
    t = time.time(); p = nd_order(mesh.TriGC, G); t_ord = time.time() - t
    perm = np.stack([2 * p, 2 * p + 1], axis=1).ravel()
    res = {}
    orderings = [("nested dissection", perm)]
    if 2 * nT <= 70_000:
        orderings.append(("x-sorted (banded)", np.arange(2 * nT)))
    for name, P in orderings:
        Mp = M[P][:, P].tocsc()
        t = time.time()
        lu = spla.splu(Mp, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        nnz = lu.L.nnz + lu.U.nnz
        cl = np.diff(lu.L.tocsc().indptr).astype(float)      # column counts of L
        ru = np.diff(lu.U.tocsr().indptr).astype(float)      # row counts of U
        flops = float(2.0 * (cl * ru).sum())
        res[name] = (nnz, flops, time.time() - t)
    print(f"nV {mesh.nV} unknowns {2 * nT}  ordering {t_ord:.1f}s")
    for name, (nnz, flops, tt) in res.items():
        print(f"   {name:20s} nnz(L+U) {nnz:.3e} = {nnz * 8 / 1e9:.2f} GB  ({nnz / (2 * nT):.0f} per row)  flops {flops:.3e}  (splu {tt:.1f}s)", flush=True)
