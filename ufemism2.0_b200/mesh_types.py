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


# --------------------------------------------------------------------------------------
# secondary data read by the ice-thickness path: edges (c-grid),
# Voronoi cell areas, shared Voronoi boundary lengths, connection vectors
# --------------------------------------------------------------------------------------
@dataclass
class MeshEdges:
    """``type_mesh`` members read by ``calc_ice_flux_divergence_matrix_upwind``
    (conservation_of_mass_utilities.f90:21-131) and ``map_velocities_from_b_to_c_2D``
    (map_velocities_to_c_grid.f90:17-69); all Fortran-ordered, 1-based, 0 = none."""

    nE: int
    VE: np.ndarray      # (nV,nC_mem) i32: edge of connection ci
    EV: np.ndarray      # (nE,4) i32: [vi, vj, vl, vr]
    ETri: np.ndarray    # (nE,2) i32: [til, tir]
    EBI: np.ndarray     # (nE,) i32
    Tricc: np.ndarray   # (nTri,2) f64 circumcentres
    A: np.ndarray       # (nV,) f64 Voronoi cell areas
    Cw: np.ndarray      # (nV,nC_mem) f64 shared Voronoi boundary lengths
    D_x: np.ndarray     # (nV,nC_mem) f64
    D_y: np.ndarray
    D: np.ndarray


def _circumcentres(V, T0):
    """Circumcentre of every triangle (plane_geometry.f90:282-309 computes the same point as the
    intersection of two perpendicular bisectors)."""
    p, q, r = V[T0[:, 0]], V[T0[:, 1]], V[T0[:, 2]]
    ax, ay = q[:, 0] - p[:, 0], q[:, 1] - p[:, 1]
    bx, by = r[:, 0] - p[:, 0], r[:, 1] - p[:, 1]
    d = 2.0 * (ax * by - ay * bx)
    a2, b2 = ax * ax + ay * ay, bx * bx + by * by
    cx = p[:, 0] + (by * a2 - ay * b2) / d
    cy = p[:, 1] + (ax * b2 - bx * a2) / d
    return np.asfortranarray(np.stack([cx, cy], axis=1))


def calc_mesh_edges(mesh: Mesh) -> MeshEdges:
    """Vectorised ``construct_mesh_edges`` (edges/mesh_edges.f90:19-194), ``calc_Voronoi_cell_areas``,
    ``calc_connection_widths``, ``calc_connection_lengths`` (mesh_secondary.f90:137-186, 246-366) for
    the synthetic meshes.  Uses the ordering conventions of C / iTri (iTri(vi,k) lies between
    C(vi,k) and C(vi,k+1)) instead of the reference's searches; tests/test_thickness.py checks it against a loop-for-loop
    restatement of those routines."""
    nV, nTri, ncm = mesh.nV, mesh.nTri, mesh.nC_mem
    C, nC, iTri, niTri, VBI = mesh.C, mesh.nC.astype(np.int64), mesh.iTri, mesh.niTri.astype(np.int64), mesh.VBI
    V = mesh.V
    T0 = mesh.Tri.astype(np.int64) - 1
    Tricc = _circumcentres(V, T0)
    tol = 1e-9 * max(mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin)
    if (Tricc[:, 0].min() < mesh.xmin - tol or Tricc[:, 0].max() > mesh.xmax + tol or
            Tricc[:, 1].min() < mesh.ymin - tol or Tricc[:, 1].max() > mesh.ymax + tol):
        # mesh_utilities.f90:150-153 crashes on such a mesh
        raise ValueError("found triangle circumcentre outside the mesh domain")
    np.clip(Tricc[:, 0], mesh.xmin, mesh.xmax, out=Tricc[:, 0])
    np.clip(Tricc[:, 1], mesh.ymin, mesh.ymax, out=Tricc[:, 1])

    wmax = int(nC.max())
    ci = np.arange(wmax, dtype=np.int64)[None, :]
    valid = ci < nC[:, None]
    vi_of = np.broadcast_to(np.arange(nV, dtype=np.int64)[:, None], (nV, wmax))
    vj = C[:, :wmax].astype(np.int64) - 1
    border = (VBI > 0)[:, None]

    # --- edges: numbered in the order (vi ascending, ci ascending) of their first visit, which is
    # from the lower-numbered end vertex
    first = valid & (vj > vi_of)
    fi, fc = np.nonzero(first)                       # row-major = (vi, ci) lexicographic
    nE = fi.size
    if nE != int(nC.sum()) // 2:
        raise ValueError("inconsistent connectivity: sum(nC)/2 != number of edges")
    VE = np.zeros((nV, ncm), dtype=np.int32, order="F")
    VE[fi, fc] = np.arange(1, nE + 1, dtype=np.int32)
    # the same edge seen from vj: find cj with C(vj,cj) == vi
    a, b = fi, vj[fi, fc]
    hit = (C[b, :wmax].astype(np.int64) - 1 == a[:, None]) & valid[b]
    cj = hit.argmax(axis=1)
    if not hit.any(axis=1).all():
        raise ValueError("asymmetric connectivity list")
    VE[b, cj] = VE[fi, fc]

    # left / right triangles and opposite vertices of vi -> vj, seen from vi
    nCi, isb = nC[fi], border[fi, 0]
    left_k = fc                                      # iTri(vi,ci) lies left of vi -> C(vi,ci)
    has_left = left_k < niTri[fi]
    til = np.where(has_left, iTri[fi, np.minimum(left_k, ncm - 1)], 0).astype(np.int32)
    vil = np.where(has_left, C[fi, (fc + 1) % np.maximum(nCi, 1)], 0).astype(np.int32)
    right_k = np.where(isb, fc - 1, (fc - 1) % np.maximum(nCi, 1))
    has_right = right_k >= 0
    tir = np.where(has_right, iTri[fi, np.maximum(right_k, 0)], 0).astype(np.int32)
    vir = np.where(has_right, C[fi, np.maximum(right_k, 0)], 0).astype(np.int32)
    EV = np.asfortranarray(np.stack([a + 1, b + 1, vil, vir], axis=1).astype(np.int32))
    ETri = np.asfortranarray(np.stack([til, tir], axis=1).astype(np.int32))

    # edge border index (mesh_edges.f90:196-228)
    ba, bb_ = VBI[a], VBI[b]
    EBI = np.zeros(nE, dtype=np.int32)
    for code, group in ((1, (8, 1, 2)), (3, (2, 3, 4)), (5, (4, 5, 6)), (7, (6, 7, 8))):
        m = (EBI == 0) & (ba > 0) & (bb_ > 0) & np.isin(ba, group) & np.isin(bb_, group)
        EBI[m] = code

    # --- connection vectors
    D_x = np.zeros((nV, ncm), order="F"); D_y = np.zeros((nV, ncm), order="F"); D = np.zeros((nV, ncm), order="F")
    vjc = np.where(valid, vj, 0)
    dx_ = np.where(valid, V[vjc, 0] - V[:, 0][:, None], 0.0)
    dy_ = np.where(valid, V[vjc, 1] - V[:, 1][:, None], 0.0)
    D_x[:, :wmax], D_y[:, :wmax] = dx_, dy_
    D[:, :wmax] = np.sqrt(dx_ ** 2 + dy_ ** 2)

    # --- shared Voronoi boundary lengths per edge (find_shared_Voronoi_boundary, mesh_utilities.f90:305-372)
    ddom = ((mesh.xmax - mesh.xmin) + (mesh.ymax - mesh.ymin)) / 100.0
    cw_e = np.zeros(nE)
    inner = EBI == 0
    if (inner & ((til == 0) | (tir == 0))).any():
        raise ValueError("non-border edge with a single adjacent triangle")
    cw_e[inner] = np.hypot(*(Tricc[til[inner] - 1] - Tricc[tir[inner] - 1]).T)
    bt = np.where(til > 0, til, tir)[~inner] - 1
    cc1 = Tricc[bt]
    cc2 = cc1.copy()
    eb = EBI[~inner]
    cc2[eb == 1, 1] = mesh.ymax + ddom
    cc2[eb == 3, 0] = mesh.xmax + ddom
    cc2[eb == 5, 1] = mesh.ymin - ddom
    cc2[eb == 7, 0] = mesh.xmin - ddom
    cw_e[~inner] = np.hypot(*(cc1 - cc2).T)
    Cw = np.zeros((nV, ncm), order="F")
    Cw[:, :wmax] = np.where(valid, cw_e[np.maximum(VE[:, :wmax].astype(np.int64) - 1, 0)], 0.0)

    # --- Voronoi cell areas (calc_Voronoi_cell with dx = 0, mesh_utilities.f90:72-303)
    wt = int(niTri.max())
    kt = np.arange(wt, dtype=np.int64)[None, :]
    tv = kt < niTri[:, None]
    tcc = Tricc[np.maximum(iTri[:, :wt].astype(np.int64) - 1, 0)]          # (nV, wt, 2)
    rel = np.where(tv[:, :, None], tcc - V[:, None, :], 0.0)
    # polygon points relative to the vertex: [first projection] + circumcentres + [last projection]
    P = np.zeros((nV, wt + 2, 2))
    P[:, 1:wt + 1] = rel
    bidx = np.nonzero(VBI > 0)[0]
    if bidx.size:
        vb = VBI[bidx]
        firstp = rel[bidx, 0].copy()
        lastp = rel[bidx, niTri[bidx] - 1].copy()
        Vb = V[bidx]
        for codes, axis, lim in (((1, 2), 1, mesh.ymax), ((3, 4), 0, mesh.xmax), ((5, 6), 1, mesh.ymin), ((7, 8), 0, mesh.xmin)):
            m = np.isin(vb, codes)
            firstp[m, axis] = lim - Vb[m, axis]
        for codes, axis, lim in (((2, 3), 0, mesh.xmax), ((4, 5), 1, mesh.ymin), ((6, 7), 0, mesh.xmin), ((8, 1), 1, mesh.ymax)):
            m = np.isin(vb, codes)
            lastp[m, axis] = lim - Vb[m, axis]
        P[bidx, 0] = firstp
        # the last projection goes right after the last circumcentre; later slots stay at the vertex
        # itself (zero vector), which adds nothing to the cross products, like the corner points do
        P[bidx, niTri[bidx] + 1] = lastp
    # free vertices: close the polygon (slot 0 = last circumcentre so that pair (0,1) is the wrap-around)
    fidx = np.nonzero(VBI == 0)[0]
    P[fidx, 0] = rel[fidx, niTri[fidx] - 1]
    cross = P[:, 1:, 0] * P[:, :-1, 1] - P[:, 1:, 1] * P[:, :-1, 0]
    A = np.abs(cross).sum(axis=1) / 2.0

    return MeshEdges(nE=nE, VE=VE, EV=EV, ETri=ETri, EBI=EBI, Tricc=Tricc, A=np.ascontiguousarray(A), Cw=Cw,
                     D_x=D_x, D_y=D_y, D=D)
