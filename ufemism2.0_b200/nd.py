"""ctypes mirror of the staged nested-dissection symbolic analysis (include/ufe_diva.h, ``ufe_nd_*``; host-only).

``analyse`` returns the elimination tree of the stiffness matrix' block graph in post-order with, per node, ``sep``
(triangles eliminated there), ``bnd`` (rest of the dense front) and ``up`` (extend-add positions in the parent's
front), plus the assembly map of every block entry; ``tests/test_host_logic.py`` drives a numpy numeric phase from
these maps to check them.  ``Solver`` is the numeric multifrontal phase on the device (``ufe_nd_solver_*``,
csrc/ufe_nd_numeric.cu): analyse once per mesh, ``factor`` per matrix, ``solve`` per right-hand side."""
from __future__ import annotations

import ctypes as ct
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import check, vp


@dataclass
class NdNode:
    level: int
    parent: int
    sep: np.ndarray
    bnd: np.ndarray
    up: np.ndarray


@dataclass
class NdTree:
    nodes: list
    n_levels: int
    max_front: int
    padded_front_bytes: float
    entry_node: np.ndarray
    entry_row: np.ndarray
    entry_col: np.ndarray
    owners: dict = None


def analyse(centroids: np.ndarray, bptr: np.ndarray, bind: np.ndarray, leaf_triangles: int = 96, nranks=()) -> NdTree:
    """centroids: (nT,2); bptr / bind: 0-based block CSR pattern over triangles.  nranks: rank counts for which the
    owner of every node is also returned (tree.owners[n] = array over the nodes)."""
    nT = centroids.shape[0]
    x = np.ascontiguousarray(centroids[:, 0], dtype=np.float64)
    y = np.ascontiguousarray(centroids[:, 1], dtype=np.float64)
    bptr = np.ascontiguousarray(bptr, dtype=np.int32)
    bind = np.ascontiguousarray(bind, dtype=np.int32)
    lib = capi.lib()
    T = ct.c_void_p()
    check(lib.ufe_nd_analyse(nT, vp(x), vp(y), vp(bptr), vp(bind), int(leaf_triangles), ct.byref(T)))
    try:
        nn, nl, mf, pb = ct.c_int32(), ct.c_int32(), ct.c_int32(), ct.c_double()
        check(lib.ufe_nd_tree_info(T, ct.byref(nn), ct.byref(nl), ct.byref(mf), ct.byref(pb)))
        P = ct.POINTER(ct.c_int32)
        nodes = []
        for i in range(nn.value):
            lv, pa, ns, nb = ct.c_int32(), ct.c_int32(), ct.c_int32(), ct.c_int32()
            s, b, u = P(), P(), P()
            check(lib.ufe_nd_tree_node(T, i, ct.byref(lv), ct.byref(pa), ct.byref(ns), ct.byref(nb), ct.byref(s), ct.byref(b), ct.byref(u)))
            arr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0, dtype=np.int32)
            nodes.append(NdNode(lv.value, pa.value, arr(s, ns.value), arr(b, nb.value), arr(u, nb.value if pa.value >= 0 else 0)))
        en, er, ec = P(), P(), P()
        check(lib.ufe_nd_tree_entry_map(T, ct.byref(en), ct.byref(er), ct.byref(ec)))
        nnzb = int(bptr[-1])
        tree = NdTree(nodes, nl.value, mf.value, pb.value, np.ctypeslib.as_array(en, shape=(nnzb,)).copy(),
                      np.ctypeslib.as_array(er, shape=(nnzb,)).copy(), np.ctypeslib.as_array(ec, shape=(nnzb,)).copy())
        tree.owners = {}
        for n in nranks:
            o = np.zeros(nn.value, dtype=np.int32)
            check(lib.ufe_nd_tree_owners(T, int(n), vp(o)))
            tree.owners[int(n)] = o
    finally:
        lib.ufe_nd_tree_free.restype = None
        lib.ufe_nd_tree_free(T)
    return tree


def block_pattern(ptr: np.ndarray, ind: np.ndarray, nT: int):
    """0-based block CSR over triangles (sorted columns) of a 0-based scalar CSR pattern with the (u,v) interleaved
    unknown numbering n = 2 ti + uv (mesh_translation_tables.f90:181-198, here 0-based)."""
    rows = np.repeat(np.arange(2 * nT, dtype=np.int64), np.diff(ptr)) // 2
    key = np.unique(rows * nT + np.asarray(ind, dtype=np.int64) // 2)
    br, bc = key // nT, key % nT
    bptr = np.zeros(nT + 1, dtype=np.int32)
    np.cumsum(np.bincount(br, minlength=nT), out=bptr[1:])
    return bptr, bc.astype(np.int32)


class Solver:
    """Exact multifrontal solve of a stiffness system on the device.  ptr / ind: scalar CSR pattern (N = 2 nT) in the
    reference's 1-based convention (type_sparse_matrix_CSR_dp: ptr(1) = 1, column indices from 1), as every other entry
    point of the C ABI takes it."""

    def __init__(self, centroids: np.ndarray, ptr: np.ndarray, ind: np.ndarray, leaf_triangles: int = 96):
        self._lib = capi.lib()
        self._T, self._S = ct.c_void_p(), ct.c_void_p()
        nT = centroids.shape[0]
        self.N = 2 * nT
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        ind = np.ascontiguousarray(ind, dtype=np.int32)
        if ptr[0] != 1:
            raise ValueError("nd.Solver: ptr / ind must be 1-based (ptr[0] == 1)")
        bptr, bind = block_pattern(ptr - 1, ind - 1, nT)
        x = np.ascontiguousarray(centroids[:, 0], dtype=np.float64)
        y = np.ascontiguousarray(centroids[:, 1], dtype=np.float64)
        check(self._lib.ufe_nd_analyse(nT, vp(x), vp(y), vp(bptr), vp(bind), int(leaf_triangles), ct.byref(self._T)))
        nn, nl, mf, pb = ct.c_int32(), ct.c_int32(), ct.c_int32(), ct.c_double()
        check(self._lib.ufe_nd_tree_info(self._T, ct.byref(nn), ct.byref(nl), ct.byref(mf), ct.byref(pb)))
        self.n_fronts, self.n_levels, self.max_front = nn.value, nl.value, mf.value
        try:
            check(self._lib.ufe_nd_solver_create(self._T, self.N, vp(ptr), vp(ind), ct.byref(self._S)))
        except Exception:
            self.close()
            raise

    def factor(self, val: np.ndarray):
        val = np.ascontiguousarray(val, dtype=np.float64)
        check(self._lib.ufe_nd_solver_factor(self._S, vp(val)))

    def solve(self, b: np.ndarray, n_refine: int = 1):
        """returns (x, |b - A x| / |b|)"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.N)
        rr = ct.c_double()
        check(self._lib.ufe_nd_solver_solve(self._S, vp(b), vp(x), int(n_refine), ct.byref(rr)))
        return x, rr.value

    def info(self):
        f, s, by, fl = ct.c_double(), ct.c_double(), ct.c_double(), ct.c_double()
        check(self._lib.ufe_nd_solver_info(self._S, ct.byref(f), ct.byref(s), ct.byref(by), ct.byref(fl)))
        return {"factor_ms": f.value, "solve_ms": s.value, "front_bytes": by.value, "factor_flops": fl.value,
                "n_fronts": self.n_fronts, "n_levels": self.n_levels, "max_front": self.max_front}

    def close(self):
        self._lib.ufe_nd_solver_free.restype = None
        self._lib.ufe_nd_tree_free.restype = None
        if self._S:
            self._lib.ufe_nd_solver_free(self._S)
            self._S = ct.c_void_p()
        if self._T:
            self._lib.ufe_nd_tree_free(self._T)
            self._T = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
