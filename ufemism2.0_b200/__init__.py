"""B200-native DIVA/SSA ice-velocity solve behind UFEMISM's call surface (host mirror).

Sub-modules: ``mesh_types`` (type_mesh), ``config`` (config keys), ``synthetic`` (test /
bench meshes + geometries), ``capi`` (ctypes binding of include/ufe_diva.h),
``diva`` (initialise_DIVA_solver / solve_DIVA / solve_SSA mirrors).
"""
__all__ = ["mesh_types", "config", "synthetic"]
