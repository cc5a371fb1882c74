"""Deterministic synthetic meshes and ice geometries for tests and benchmarks.

The reference's mesh generator (Ruppert refinement + Lloyd, ``src/UPSY/mesh/``) is out of
scope; the velocity solve consumes a finished ``type_mesh``.  This
module produces ``type_mesh``-conformant meshes (see ``mesh_types.py``) of any size:
a jittered lattice triangulation of a rectangle, x-sorted like
``mesh_contiguous_domains.f90:45,140`` so that index ranges are vertical strips.

Geometries follow the reference's closed forms where they exist
(``src/UFEMISM/reference_geometries/idealised_geometries.f90``): ISMIP-HOM A (:243-267),
ISMIP-HOM C/D (:294-317), MISMIP+ bed (:357-401), SSA_icestream (:186-206); the
thickness fields for MISMIP+ and the Antarctic-scale dome are synthetic.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesh_types import Mesh, build_mesh_from_triangles

ice_density = 910.0          # parameters.f90:52
seawater_density = 1028.0    # parameters.f90:54
grav = 9.81                  # parameters.f90:49
pi = 3.141592653589793       # parameters.f90:44

SEED = 20261017


def lattice_mesh(xmin, xmax, ymin, ymax, nx, ny, jitter=0.2, seed=SEED, nz=12, delaunay=False) -> Mesh:
    """nx x ny lattice, alternating cell diagonals, interior vertices jittered by
    ``jitter*h*U(-1,1)`` (PCG64), vertices and triangles sorted by x (centroid x).
    ``delaunay=True`` re-triangulates the jittered points (scipy.spatial.Delaunay) so that the
    Voronoi dual the ice-thickness path reads (cell areas, shared boundary lengths) is a proper
    tessellation, as it is on the reference's own (Delaunay-refined) meshes."""
    assert nx >= 3 and ny >= 3
    dx = (xmax - xmin) / (nx - 1)
    dy = (ymax - ymin) / (ny - 1)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    x = xmin + ii * dx
    y = ymin + jj * dy
    # exact borders
    x[-1, :] = xmax
    y[:, -1] = ymax
    rng = np.random.Generator(np.random.PCG64(seed))
    jx = rng.uniform(-1.0, 1.0, size=x.shape) * jitter * dx
    jy = rng.uniform(-1.0, 1.0, size=y.shape) * jitter * dy
    interior = np.zeros(x.shape, dtype=bool)
    interior[1:-1, 1:-1] = True
    x = np.where(interior, x + jx, x)
    y = np.where(interior, y + jy, y)

    VBI = np.zeros((nx, ny), dtype=np.int32)
    VBI[:, -1] = 1
    VBI[-1, :] = 3
    VBI[:, 0] = 5
    VBI[0, :] = 7
    VBI[-1, -1] = 2
    VBI[-1, 0] = 4
    VBI[0, 0] = 6
    VBI[0, -1] = 8

    vid = (ii * ny + jj).astype(np.int64)
    a = vid[:-1, :-1].ravel()
    b = vid[1:, :-1].ravel()
    c = vid[1:, 1:].ravel()
    d = vid[:-1, 1:].ravel()
    even = ((ii[:-1, :-1] + jj[:-1, :-1]) % 2 == 0).ravel()
    t1 = np.where(even[:, None], np.stack([a, b, c], 1), np.stack([a, b, d], 1))
    t2 = np.where(even[:, None], np.stack([a, c, d], 1), np.stack([b, c, d], 1))
    Tri0 = np.concatenate([t1, t2], axis=0)
    V = np.stack([x.ravel(), y.ravel()], axis=1)
    VBI = VBI.ravel()
    if delaunay:
        from scipy.spatial import Delaunay
        Tri0 = Delaunay(V).simplices.astype(np.int64)
        p, q, r = V[Tri0[:, 0]], V[Tri0[:, 1]], V[Tri0[:, 2]]
        area2 = (q[:, 0] - p[:, 0]) * (r[:, 1] - p[:, 1]) - (q[:, 1] - p[:, 1]) * (r[:, 0] - p[:, 0])
        Tri0 = Tri0[np.abs(area2) > 1e-9 * dx * dy]              # slivers between collinear border points
        area2 = area2[np.abs(area2) > 1e-9 * dx * dy]
        cw = area2 < 0
        Tri0[cw] = Tri0[cw][:, [0, 2, 1]]                         # counter-clockwise

    # x-sort vertices (mesh_contiguous_domains.f90:45-138)
    vperm = np.argsort(V[:, 0], kind="stable")
    inv = np.empty_like(vperm)
    inv[vperm] = np.arange(vperm.size)
    V = V[vperm]
    VBI = VBI[vperm]
    Tri0 = inv[Tri0]
    # x-sort triangles by geometric centre (mesh_contiguous_domains.f90:140-230)
    gcx = V[Tri0, 0].sum(axis=1) / 3.0
    tperm = np.argsort(gcx, kind="stable")
    Tri0 = Tri0[tperm]
    return build_mesh_from_triangles(V, (Tri0 + 1).astype(np.int32), VBI, xmin, xmax,
                                     ymin, ymax, nz=nz)


# ----------------------------------------------------------------------------------
# ice-model inputs consumed read-only by the velocity solve (type_ice_model subset,
# src/UFEMISM/types/ice_model_types.f90:208+, and type_bed_roughness_model)
# ----------------------------------------------------------------------------------
@dataclass
class IceInputs:
    Hi: np.ndarray
    Hb: np.ndarray
    Hs: np.ndarray
    SL: np.ndarray
    Hib: np.ndarray
    fraction_gr: np.ndarray          # (nV)
    fraction_gr_b: np.ndarray        # (nTri)
    effective_pressure: np.ndarray   # (nV)
    mask_grounded_ice: np.ndarray    # int32 (nV)
    mask_floating_ice: np.ndarray
    mask_icefree_land: np.ndarray
    mask_icefree_ocean: np.ndarray
    Ti: np.ndarray                   # (nV,nz) F  (Huybrechts1992 only)
    till_friction_angle: np.ndarray  # (nV)
    alpha_sq: np.ndarray
    beta_sq: np.ndarray


def ice_surface_elevation(Hi, Hb, SL):
    """``ice_geometry_basics.f90``: Hs = Hi + max(SL - rho_i/rho_w Hi, Hb)."""
    return Hi + np.maximum(SL - ice_density / seawater_density * Hi, Hb)


def thickness_above_floatation(Hi, Hb, SL):
    return Hi - np.maximum(0.0, (SL - Hb) * (seawater_density / ice_density))


def _finish_inputs(mesh: Mesh, Hi, Hb, SL, phi=10.0, alpha_sq=0.5, beta_sq=1.0e4,
                   Ti_val=260.0, hydrology="Martin2011") -> IceInputs:
    nV = mesh.nV
    Hi = np.ascontiguousarray(Hi, dtype=np.float64)
    Hb = np.ascontiguousarray(Hb, dtype=np.float64)
    SL = np.ascontiguousarray(SL, dtype=np.float64)
    Hs = ice_surface_elevation(Hi, Hb, SL)
    Hib = Hs - Hi
    TAF = thickness_above_floatation(Hi, Hb, SL)
    grounded = (TAF > 0.0) & (Hi > 0.0)
    floating = (TAF <= 0.0) & (Hi > 0.0)
    land = (Hi <= 0.0) & (Hb >= SL)
    ocean = (Hi <= 0.0) & (Hb < SL)
    # sub-grid grounded fraction: linear ramp across |TAF| < 50 m
    fg = np.clip(0.5 + TAF / 100.0, 0.0, 1.0)
    fg = np.where(Hi > 0.0, fg, np.where(Hb >= SL, 1.0, 0.0))
    T0 = mesh.Tri.astype(np.int64) - 1
    fg_b = (fg[T0[:, 0]] + fg[T0[:, 1]] + fg[T0[:, 2]]) / 3.0
    # effective pressure (basal_hydrology_main.f90:45-48, 97-101), Hi_eff = Hi
    if hydrology == "Martin2011":
        lam = np.clip(1.0 - (Hb - SL - 0.0) / (1000.0 - 0.0), 0.0, 1.0)
        pw = 0.96 * ice_density * grav * Hi * lam
    else:  # 'none'
        pw = np.zeros(nV)
    Neff = np.maximum(0.0, ice_density * grav * Hi - pw)
    Ti = np.full((nV, mesh.nz), Ti_val, dtype=np.float64, order="F")
    return IceInputs(
        Hi=Hi, Hb=Hb, Hs=Hs, SL=SL, Hib=Hib, fraction_gr=fg,
        fraction_gr_b=np.ascontiguousarray(fg_b), effective_pressure=Neff,
        mask_grounded_ice=grounded.astype(np.int32), mask_floating_ice=floating.astype(np.int32),
        mask_icefree_land=land.astype(np.int32), mask_icefree_ocean=ocean.astype(np.int32),
        Ti=Ti, till_friction_angle=np.full(nV, phi), alpha_sq=np.full(nV, alpha_sq),
        beta_sq=np.full(nV, beta_sq))


def geometry_ISMIP_HOM_A(mesh: Mesh, L: float) -> IceInputs:
    """idealised_geometries.f90:243-267."""
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    Hs = 2000.0 - x * np.tan(0.5 * pi / 180.0)
    Hb = Hs - 1000.0 + 500.0 * np.sin(x * 2.0 * pi / L) * np.sin(y * 2.0 * pi / L)
    Hi = Hs - Hb
    return _finish_inputs(mesh, Hi, Hb, np.full(mesh.nV, -10000.0), hydrology="none")


def geometry_ISMIP_HOM_C(mesh: Mesh, L: float) -> IceInputs:
    """idealised_geometries.f90:294-317 (experiments C and D share the geometry)."""
    x = mesh.V[:, 0]
    Hs = 2000.0 - x * np.tan(0.1 * pi / 180.0)
    Hb = Hs - 1000.0
    Hi = Hs - Hb
    return _finish_inputs(mesh, Hi, Hb, np.full(mesh.nV, -10000.0), hydrology="none")


def geometry_SSA_icestream(mesh: Mesh, H=2000.0, dhdx=-3.0e-4) -> IceInputs:
    """idealised_geometries.f90:186-206."""
    x = mesh.V[:, 0]
    Hi = np.full(mesh.nV, H)
    Hb = dhdx * x
    return _finish_inputs(mesh, Hi, Hb, np.full(mesh.nV, -10000.0), hydrology="none")


def mismipplus_bed(x, y):
    """idealised_geometries.f90:369-397."""
    B0, B2, B4, B6 = -150.0, -728.8, 343.91, -50.57
    xbar, fc, dc, wc, zbdeep = 300000.0, 4000.0, 500.0, 24000.0, -720.0
    xt = x / xbar
    Bx = B0 + B2 * xt ** 2 + B4 * xt ** 4 + B6 * xt ** 6
    By = dc / (1.0 + np.exp(-2.0 * (y - wc) / fc)) + dc / (1.0 + np.exp(2.0 * (y + wc) / fc))
    return np.maximum(Bx + By, zbdeep)


def geometry_MISMIPplus(mesh: Mesh, x_gl=450e3, H0=1500.0, H_shelf=300.0, calving_front=None, profile="surface",
                        s0=1350.0) -> IceInputs:
    """MISMIP+ bed (idealised_geometries.f90:357-401) with a synthetic ice body.

    ``profile="surface"`` (default, round 2): the ice SURFACE is prescribed as a function of x only --
    a parabola from ``s0`` at the divide to the floatation surface at ``x_gl`` (zero slope at the divide,
    finite slope at the grounding line), then the surface of a shelf thinning linearly to ``H_shelf`` at
    x = 640 km and flat beyond -- and the thickness follows from it: ``Hi = min(s - Hb, s / (1 - rho_i /
    rho_w))``, i.e. grounded where the bed is above the floating draft (the channel walls), floating
    elsewhere.  The surface is level across the channel, so the driving stress is along the flow as in a
    MISMIP+ steady state (speeds up to ~100 m/yr).

    ``profile="vialov"`` (round 1): thickness a function of x only, Vialov shape with an infinite slope at
    the grounding line; on the 500 m high channel walls this gives 500 m of surface relief across 8 km,
    velocities pinned at the ``vel_max`` cap and a Picard iteration that never converges.  Kept for the
    round-1 test cases.

    ``calving_front=640e3`` removes the ice beyond it (ice-free ocean)."""
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    Hb = mismipplus_bed(x, y)
    SL = np.zeros(mesh.nV)
    Hb_gl = mismipplus_bed(np.array([x_gl]), np.array([0.0]))[0]
    if profile == "surface":
        r = 1.0 - ice_density / seawater_density
        s_gl = -Hb_gl * (seawater_density / ice_density) * r + 4.0     # just grounded on the centre line at x_gl
        s_sh = H_shelf * r
        sg = s_gl + (s0 - s_gl) * (1.0 - np.clip(x / x_gl, 0.0, 1.0) ** 2)
        t = np.clip((x - x_gl) / (640e3 - x_gl), 0.0, 1.0)
        sf = s_gl + (s_sh - s_gl) * t
        s = np.where(x <= x_gl, sg, sf)
        Hi = np.minimum(s - Hb, s / r)
    elif profile == "vialov":
        # floatation thickness at the grounding line keeps the profile continuous there
        H_gl = max(H_shelf + 50.0, -Hb_gl * seawater_density / ice_density + 20.0)
        s = np.clip(x / x_gl, 0.0, 1.0)
        Hg = H_gl + (H0 - H_gl) * (1.0 - s ** (4.0 / 3.0)) ** (3.0 / 8.0)
        t = np.clip((x - x_gl) / (640e3 - x_gl), 0.0, 1.0)
        Hf = H_gl + (H_shelf - H_gl) * t
        Hi = np.where(x <= x_gl, Hg, Hf)
    else:
        raise ValueError(profile)
    if calving_front is not None:
        Hi = np.where(x > calving_front, 0.0, Hi)
    return _finish_inputs(mesh, Hi, Hb, SL, hydrology="Martin2011")


def geometry_antarctic_dome(mesh: Mesh, seed=SEED) -> IceInputs:
    """Antarctic-shaped synthetic geometry: radial dome on an
    undulating bed, floating where thin."""
    x, y = mesh.V[:, 0], mesh.V[:, 1]
    r = np.sqrt(x * x + y * y)
    R = 0.79 * min(mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin) / 2.0
    s = np.clip(r / R, 0.0, 1.0)
    Hi = 3500.0 * (1.0 - s ** (4.0 / 3.0)) ** (3.0 / 8.0)
    Hi = np.where(r < R, np.maximum(Hi, 150.0), 0.0)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    ph = rng.uniform(0.0, 2.0 * pi, size=4)
    noise = 40.0 * (np.sin(2 * pi * x / 2.3e5 + ph[0]) * np.cos(2 * pi * y / 1.9e5 + ph[1])
                    + np.sin(2 * pi * x / 0.9e5 + ph[2]) * np.sin(2 * pi * y / 1.1e5 + ph[3]))
    Hb = -200.0 + 600.0 * np.sin(2 * pi * x / 1e6) * np.cos(2 * pi * y / 8e5) - 1e-4 * r + noise
    return _finish_inputs(mesh, Hi, Hb, np.zeros(mesh.nV), hydrology="Martin2011")
