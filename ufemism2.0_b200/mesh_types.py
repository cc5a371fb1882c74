"""Host-side mirror of the reference's ``type_mesh`` (primary + secondary data).

Mirrors ``src/UPSY/types/mesh_types.f90:16-279`` of the reference: every array is
column-major (Fortran order), 1-based, with 0 meaning "none" in ``TriC`` and padding in
``C`` / ``iTri``.  Only the members the DIVA/SSA velocity solve reads are kept.

Conventions (``src/UPSY/mesh/mesh_dummy_meshes.f90:61-113``):
  * ``Tri(ti,:)`` counter-clockwise; ``TriC(ti,n)`` = triangle across the edge opposite
    vertex ``n``;
  * ``C(vi,:)`` / ``iTri(vi,:)`` counter-clockwise, starting at the domain border for
    border vertices; ``iTri(vi,k)`` lies between ``C(vi,k)`` and ``C(vi,k+1)``;
  * ``VBI``: 1=N, 2=NE, 3=E, 4=SE, 5=S, 6=SW, 7=W, 8=NW, 0=interior;
  * ``TriGC`` = mean of the three vertices (``mesh_secondary.f90:432``);
  * ``TriBI`` by tracing the border from the SW corner (``mesh_secondary.f90:72-135``).

``build_mesh_from_triangles`` derives all connectivity from (V, Tri, VBI) with vectorised
numpy so that ~1e6-vertex synthetic meshes are built in seconds.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

NC_MEM = 32  # mesh%nC_mem, reference default (mesh_types.f90)


@dataclass
class Mesh:
    """``type_mesh`` subset. All index arrays int32, Fortran-ordered, 1-based."""

    nV: int
    nTri: int
    nC_mem: int
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    V: np.ndarray        # (nV,2) f64 F
    Tri: np.ndarray      # (nTri,3) i32 F
    TriC: np.ndarray     # (nTri,3) i32 F
    C: np.ndarray        # (nV,nC_mem) i32 F
    nC: np.ndarray       # (nV,) i32
    iTri: np.ndarray     # (nV,nC_mem) i32 F
    niTri: np.ndarray    # (nV,) i32
    VBI: np.ndarray      # (nV,) i32
    TriBI: np.ndarray    # (nTri,) i32
    TriGC: np.ndarray    # (nTri,2) f64 F
    nz: int = 12
    zeta: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # operator matrices (filled by calc_all_matrix_operators_mesh); dict name -> CSR
    ops: dict = field(default_factory=dict)

    # translation tables, mesh_translation_tables.f90:181-198
    def tiuv2n(self, ti, uv):
        return 2 * (ti - 1) + uv


def zeta_regular(nz: int) -> np.ndarray:
    """``initialise_scaled_vertical_coordinate_regular`` (mesh_zeta.f90:58-80)."""
    return np.array([(k - 1) / (nz - 1) for k in range(1, nz + 1)], dtype=np.float64)


def zeta_irregular_log(nz: int, R: float) -> np.ndarray:
    """``initialise_scaled_vertical_coordinate_irregular_log`` (mesh_zeta.f90:82-118)."""
    if R == 1.0:
        return zeta_regular(nz)
    z = np.zeros(nz)
    for k in range(1, nz + 1):
        sigma = (k - 1) / (nz - 1)
        z[nz - k] = 1.0 - (R ** sigma - 1.0) / (R - 1.0)
    return z


def build_mesh_from_triangles(V, Tri, VBI, xmin, xmax, ymin, ymax, nz=12,
                              nC_mem=NC_MEM, zeta=None) -> Mesh:
    """Derive TriC, C, nC, iTri, niTri, TriGC, TriBI from vertices + CCW triangles.

    V: (nV,2) float64; Tri: (nTri,3) 1-based CCW; VBI: (nV,) border indices.
    """
    V = np.asfortranarray(V, dtype=np.float64)
    Tri = np.asfortranarray(Tri, dtype=np.int32)
    VBI = np.ascontiguousarray(VBI, dtype=np.int32)
    nV, nTri = V.shape[0], Tri.shape[0]
    T0 = Tri.astype(np.int64) - 1

    # --- TriC: neighbour across the edge opposite local vertex n
    # edge opposite vertex n is (n+1, n+2) (cyclic)
    ea = np.stack([T0[:, 1], T0[:, 2], T0[:, 0]], axis=1)  # start of edge opposite n
    eb = np.stack([T0[:, 2], T0[:, 0], T0[:, 1]], axis=1)
    lo = np.minimum(ea, eb).ravel()
    hi = np.maximum(ea, eb).ravel()
    key = lo * nV + hi
    tri_of = np.repeat(np.arange(nTri, dtype=np.int64), 3)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same_next = np.zeros(ks.shape, dtype=bool)
    same_next[:-1] = ks[1:] == ks[:-1]
    TriC_flat = np.zeros(3 * nTri, dtype=np.int32)
    i_first = np.nonzero(same_next)[0]
    a, b = order[i_first], order[i_first + 1]
    TriC_flat[a] = tri_of[b] + 1
    TriC_flat[b] = tri_of[a] + 1
    TriC = np.asfortranarray(TriC_flat.reshape(nTri, 3))

    # --- TriGC
    TriGC = np.asfortranarray((V[T0[:, 0]] + V[T0[:, 1]] + V[T0[:, 2]]) / 3.0)

    # --- iTri / C: sort the triangles around each vertex counter-clockwise
    v_of = T0.ravel()                              # vertex of each (tri, n) corner
    t_of = tri_of
    nxt = np.stack([T0[:, 1], T0[:, 2], T0[:, 0]], axis=1).ravel()   # p: next vertex CCW in tri
    prv = np.stack([T0[:, 2], T0[:, 0], T0[:, 1]], axis=1).ravel()   # q: previous
    gcx = np.repeat(TriGC[:, 0], 3) - V[v_of, 0]
    gcy = np.repeat(TriGC[:, 1], 3) - V[v_of, 1]
    theta = np.arctan2(gcy, gcx)
    # start direction for border vertices: S,SW -> east; E,SE -> north; N,NE -> west; W,NW -> south
    theta0_tab = np.array([-np.pi, np.pi, np.pi, 0.5 * np.pi, 0.5 * np.pi, 0.0, 0.0,
                           -0.5 * np.pi, -0.5 * np.pi])
    th0 = theta0_tab[VBI[v_of]]
    rel = np.mod(theta - th0, 2.0 * np.pi)
    order = np.lexsort((rel, v_of))
    v_s, t_s, p_s, q_s = v_of[order], t_of[order], nxt[order], prv[order]
    niTri = np.bincount(v_of, minlength=nV).astype(np.int32)
    start = np.zeros(nV + 1, dtype=np.int64)
    np.cumsum(niTri, out=start[1:])
    pos = np.arange(v_s.shape[0], dtype=np.int64) - start[v_s]
    if niTri.max() + 1 > nC_mem:
        raise ValueError("vertex degree exceeds nC_mem")
    iTri = np.zeros((nV, nC_mem), dtype=np.int32, order="F")
    C = np.zeros((nV, nC_mem), dtype=np.int32, order="F")
    iTri[v_s, pos] = t_s + 1
    C[v_s, pos] = p_s + 1
    is_border = VBI > 0
    nC = niTri.copy()
    nC[is_border] += 1
    last = start[1:] - 1                           # last sorted corner of each vertex
    bv = np.nonzero(is_border)[0]
    C[bv, niTri[bv]] = q_s[last[bv]] + 1

    # --- TriBI: trace the border from the SW corner (mesh_secondary.f90:72-135)
    TriBI = np.zeros(nTri, dtype=np.int32)
    sw = np.nonzero(VBI == 6)[0]
    if sw.size:
        vi_SW = int(sw[0])
        vi = vi_SW
        corner = {}
        n_cycles = 0
        while True:
            n_cycles += 1
            if n_cycles > nV:
                raise RuntimeError("got stuck tracing the domain border")
            TriBI[iTri[vi, : niTri[vi]] - 1] = VBI[vi]
            vi = int(C[vi, nC[vi] - 1]) - 1
            if VBI[vi] in (4, 2, 8):
                corner[int(VBI[vi])] = vi
            if vi == vi_SW:
                break
        corner[6] = vi_SW
        for bi, vc in corner.items():
            if niTri[vc] == 1:
                TriBI[iTri[vc, 0] - 1] = bi

    if zeta is None:
        zeta = zeta_regular(nz)
    return Mesh(nV=nV, nTri=nTri, nC_mem=nC_mem, xmin=float(xmin), xmax=float(xmax),
                ymin=float(ymin), ymax=float(ymax), V=V, Tri=Tri, TriC=TriC, C=C, nC=nC,
                iTri=iTri, niTri=niTri, VBI=VBI, TriBI=TriBI, TriGC=TriGC, nz=nz,
                zeta=np.ascontiguousarray(zeta, dtype=np.float64))


def dummy_mesh_5(xmin, xmax, ymin, ymax, nz=12) -> Mesh:
    """The reference's 5-vertex seed mesh (mesh_dummy_meshes.f90:23-123), rebuilt from
    its (V, Tri, VBI) only; used to check the connectivity builder against the arrays
    the reference hard-codes."""
    V = np.array([[xmin, ymin], [xmax, ymin], [xmax, ymax], [xmin, ymax],
                  [(xmin + xmax) / 2, (ymin + ymax) / 2]], dtype=np.float64)
    Tri = np.array([[1, 2, 5], [2, 3, 5], [3, 4, 5], [4, 1, 5]], dtype=np.int32)
    VBI = np.array([6, 4, 2, 8, 0], dtype=np.int32)
    return build_mesh_from_triangles(V, Tri, VBI, xmin, xmax, ymin, ymax, nz=nz)
