"""ctypes mirror of the staged nested-dissection symbolic analysis (include/ufe_diva.h, ``ufe_nd_*``; host-only).

``analyse`` returns the elimination tree of the stiffness matrix' block graph in post-order with, per node, ``sep``
(triangles eliminated there), ``bnd`` (rest of the dense front) and ``up`` (extend-add positions in the parent's
front), plus the assembly map of every block entry.  The numeric multifrontal phase on the device is the next step
(DESIGN.md section 9); ``tests/test_host_logic.py`` drives a numpy numeric phase from these maps to check them."""
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


def analyse(centroids: np.ndarray, bptr: np.ndarray, bind: np.ndarray, leaf_triangles: int = 96) -> NdTree:
    """centroids: (nT,2); bptr / bind: 0-based block CSR pattern over triangles."""
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
    finally:
        lib.ufe_nd_tree_free.restype = None
        lib.ufe_nd_tree_free(T)
    return tree
