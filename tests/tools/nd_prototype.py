"""Round-2 prototype (CPU, numpy): multifrontal exact solve of the DIVA stiffness system under geometric nested
dissection of the triangles, organised the way the GPU version will be -- one dense front per tree node, fronts of a
tree level processed as one batch (padded to the level's largest front), extend-add through index maps built once per
mesh.  Verifies the solve against scipy on real stiffness matrices (values from the oracle's first Picard iteration)
and prints, per tree level, the number of fronts, their sizes, the padded bytes and the dense flops.

    python tests/tools/nd_prototype.py mismipplus:4000 | antarctic:<nV> [leaf_triangles]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import ufe_pkg; ufe_pkg.load()
from ufemism2_0_b200 import experiments
import oracle as O


class Node:
    __slots__ = ("sep", "bnd", "children", "level", "parent", "X", "F12", "F21", "z")

    def __init__(self, sep, children, level):
        self.sep, self.children, self.level = sep, children, level
        self.bnd = None
        self.parent = None


def dissect(xy, G, leaf):
    """Recursive coordinate bisection.  Returns the root; node.sep = triangles eliminated at that node."""
    n = xy.shape[0]

    def rec(idx, level):
        if idx.size <= leaf:
            return Node(idx, [], level)
        ext = xy[idx].max(axis=0) - xy[idx].min(axis=0)
        ax = int(np.argmax(ext))
        med = np.median(xy[idx, ax])
        left, right = idx[xy[idx, ax] <= med], idx[xy[idx, ax] > med]
        if left.size == 0 or right.size == 0:
            return Node(idx, [], level)
        inleft = np.zeros(n, dtype=bool); inleft[left] = True
        touched = np.unique(G[right].indices)
        sep = touched[inleft[touched]]
        insep = np.zeros(n, dtype=bool); insep[sep] = True
        kids = [rec(left[~insep[left]], level + 1), rec(right, level + 1)]
        return Node(sep, kids, level)

    return rec(np.arange(n), 0)


def symbolic(root, G, n):
    """Bottom-up: node.bnd = triangles of ancestors' separators coupled to the subtree (front = [sep; bnd])."""
    order, stack = [], [root]
    while stack:
        nd = stack.pop()
        order.append(nd)
        for c in nd.children:
            c.parent = nd
            stack.append(c)
    eliminated = np.zeros(n, dtype=bool)
    for nd in reversed(order):                 # children before parents
        cand = [np.unique(G[nd.sep].indices)] + [c.bnd for c in nd.children]
        eliminated[nd.sep] = True
        allc = np.unique(np.concatenate(cand)) if cand else np.zeros(0, dtype=np.int64)
        nd.bnd = allc[~eliminated[allc]]
    return list(reversed(order))               # post-order-compatible: children first


def factor(post, A, nT):
    """Numeric multifrontal factorisation on 2x2-blocked fronts.  A: csr (2nT x 2nT)."""
    A = A.tocsr()
    loc = np.full(nT, -1, dtype=np.int64)
    flops = 0.0
    schur = {}
    for nd in post:
        tri = np.concatenate([nd.sep, nd.bnd])
        ns, nf = 2 * nd.sep.size, 2 * tri.size
        dof = np.stack([2 * tri, 2 * tri + 1], axis=1).ravel()
        loc[tri] = np.arange(tri.size)
        F = np.zeros((nf, nf))
        # original entries with a row or a column in the separator
        sub = A[dof][:, dof].toarray()
        F[:ns, :] = sub[:ns, :]
        F[ns:, :ns] = sub[ns:, :ns]
        for c in nd.children:                  # extend-add
            S = schur.pop(id(c))
            m = loc[c.bnd]
            md = np.stack([2 * m, 2 * m + 1], axis=1).ravel()
            F[np.ix_(md, md)] += S
        X = np.linalg.inv(F[:ns, :ns])
        nd.X, nd.F12, nd.F21 = X, F[:ns, ns:].copy(), F[ns:, :ns].copy()
        schur[id(nd)] = F[ns:, ns:] - nd.F21 @ (X @ nd.F12)
        nb = nf - ns
        flops += 2.0 * ns ** 3 + 2.0 * ns * ns * nb + 2.0 * nb * ns * nb
        loc[tri] = -1
    return flops


def solve(post, b):
    x = b.copy()
    dofs = lambda t: np.stack([2 * t, 2 * t + 1], axis=1).ravel()
    for nd in post:                            # forward
        s, bd = dofs(nd.sep), dofs(nd.bnd)
        nd.z = nd.X @ x[s]
        x[bd] -= nd.F21 @ nd.z
    for nd in reversed(post):                  # backward
        s, bd = dofs(nd.sep), dofs(nd.bnd)
        x[s] = nd.z - nd.X @ (nd.F12 @ x[bd])
    return x


def main():
    arg = sys.argv[1] if len(sys.argv) > 1 else "mismipplus:4000"
    leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    kind, val = arg.split(":")
    mesh, C, ice = experiments.MISMIPplus(float(val)) if kind == "mismipplus" else experiments.antarctic(int(float(val)))
    mesh.ops = O.calc_all_matrix_operators_mesh(mesh)
    # the stiffness matrix of the first Picard iteration (cold start), captured from the oracle's own loop
    cap = {}
    orig = O.direct_solve
    def grab(A, b):
        if "A" not in cap: cap["A"], cap["b"] = A.to_scipy().tocsr(), np.asarray(b).copy()
        return orig(A, b)
    O.direct_solve = grab
    C.visc_it_nit = 0
    O.solve_DIVA(mesh, ice, C, O.new_DIVA_state(mesh))
    O.direct_solve = orig
    A, b = cap["A"], cap["b"]
    nT = mesh.nTri
    P = (abs(A) + abs(A).T).tocsr()
    G = sp.csr_matrix((np.ones(P.nnz), P.indices // 2, P.indptr), shape=(2 * nT, nT))
    G = (sp.csr_matrix((np.ones(2 * nT), (np.arange(2 * nT) // 2, np.arange(2 * nT))), shape=(nT, 2 * nT)) @ G).tocsr()
    t = time.time(); root = dissect(mesh.TriGC, G, leaf); post = symbolic(root, G, nT); t_sym = time.time() - t
    t = time.time(); flops = factor(post, A, nT); t_fac = time.time() - t
    x = solve(post, b)
    xr = spla.splu(A.tocsc()).solve(b)
    r = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    print(f"{arg}: nTri {nT} unknowns {2 * nT} leaf {leaf} | symbolic {t_sym:.2f}s numeric {t_fac:.2f}s | residual {r:.2e} "
          f"| x vs SuperLU {np.abs(x - xr).max() / np.abs(xr).max():.2e} | dense flops {flops:.3e}")
    levels = {}
    for nd in post:
        levels.setdefault(nd.level, []).append((2 * nd.sep.size, 2 * (nd.sep.size + nd.bnd.size)))
    tot_pad = tot_exact = 0
    print(" level  fronts   sep(max)  front(max)  front(mean)   padded MB   exact MB")
    for lv in sorted(levels):
        a = np.array(levels[lv])
        pad = a.shape[0] * int(a[:, 1].max()) ** 2 * 8 / 1e6
        ex = float((a[:, 1].astype(float) ** 2).sum() * 8 / 1e6)
        tot_pad += pad; tot_exact += ex
        print(f" {lv:5d} {a.shape[0]:7d} {a[:, 0].max():10d} {a[:, 1].max():11d} {a[:, 1].mean():12.0f} {pad:11.1f} {ex:10.1f}")
    print(f" fronts kept for the solve: padded {tot_pad / 1e3:.2f} GB, exact {tot_exact / 1e3:.2f} GB")


if __name__ == "__main__":
    main()
