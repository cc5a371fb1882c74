#!/usr/bin/env python
"""Benchmark of the DIVA velocity solve (BASELINE.json metric: DIVA velocity solves/sec on
the MISMIP+ 2 km mesh; SpMV HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload mismipplus_2km|mismip_8km|ismip_hom_a|antarctic_1m|antarctic:<nV>]

A "step" is one full cold-start ``solve_DIVA`` call (Picard loop to
``visc_it_norm_dUV_tol`` or ``visc_it_nit``, DIVA_main.f90:88-262) on synthetic geometry of
the named mesh.  Legs of the b200 arm:

  value : solves/s with every input resident in HBM (``ufe_diva_reset_state`` +
          ``ufe_diva_solve_resident``), device time (CUDA events inside the library), max over
          ranks, barrier + synchronize on both sides;
  e2e   : the same solve through the reference-facing call ``ufe_diva_solve`` with pinned HOST
          buffers - H2D of all ice inputs + state and D2H of all results inside the timed region;
  roofline : the Krylov MatMult kernel (k_kspmv) on the resident stiffness matrix, CUDA events on
          the launching stream, L2 flushed between launches (ufe_bench_spmv);
  wide_mesh_nd_lu : reported beside the metric: one converged cold-start solve_DIVA on a wide (square, Antarctic-shaped)
                    mesh with the multifrontal nested-dissection preconditioner (krylov_pc = nd_lu), one GPU.
  thickness_update(_large_mesh) : SURVEY.md 8f rank 2, reported beside the metric, not part of it: one
          ``calc_dHi_dt_semiimplicit`` call on the velocities the solve left on the device (host buffers in and
          out), its device-time split and the achieved GB/s of k_thk_divq (N = 1 only);
  cpu_baseline : the oracle's restatement of the reference CPU path (GMRES(30) + block-Jacobi
          ILU(0), per-iteration re-assembly) on a bounded sample, rank 0, N = 1 only.

``--impl reference`` times only the CPU restatement (the reference itself - Fortran + MPI +
PETSc + NetCDF - cannot be built in this image; DESIGN.md "Oracle").
Multi-GPU: launched by torchrun, one rank per GPU; the mesh is partitioned by
``partition_list`` ranges inside the library; NCCL unique id broadcast through
torch.distributed.  One solve is shared by all ranks => "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


COMM_LABEL = {0: "NCCL", 1: "peer memory (IPC/NVLink) inside the Krylov loop, NCCL outside",
              2: "redundant: the system is below the partitioning threshold (131072 unknowns), every rank solves it whole, no communication"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mismipplus_2km")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target size of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spmv-reps", type=int, default=200)
    ap.add_argument("--no-large-roofline", action="store_true",
                    help="skip the SpMV roofline leg on the ~1M-vertex mesh (the configuration the >= 70 %% target is quoted on)")
    ap.add_argument("--large-vertices", type=int, default=1_000_000)
    ap.add_argument("--wide-vertices", type=int, default=100_000,
                    help="size of the wide-mesh full-solve leg (krylov_pc = nd_lu); 0 skips it")
    return ap.parse_args()


def make_workload(name):
    import ufe_pkg
    ufe_pkg.load()
    from ufemism2_0_b200 import experiments
    if name == "mismipplus_2km":
        mesh, C, ice = experiments.MISMIPplus(2e3)
        label = "MISMIP+ 800x80 km, uniform 2 km synthetic mesh (config_MISMIPplus_2km_spinup.cfg keys), cold-start DIVA solve"
    elif name == "mismip_8km":
        mesh, C, ice = experiments.MISMIP_8km()
        label = "MISMIP 2000x2000 km, 8 km synthetic mesh (config_MISMIP_8km_spinup_for_scaling.cfg keys), cold-start DIVA solve"
    elif name == "ismip_hom_a":
        mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 41)
        label = "ISMIP-HOM A, L = 160 km, 41x41 lattice, cold-start DIVA solve"
    elif name.startswith("antarctic"):
        nV = 1_000_000 if name == "antarctic_1m" else int(name.split(":")[1])
        mesh, C, ice = experiments.antarctic(nV)
        label = f"synthetic Antarctic-scale dome, ~{nV} vertices (config_ant_template.cfg keys), cold-start DIVA solve"
    else:
        raise SystemExit(f"unknown workload {name}")
    return mesh, C, ice, label


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned_like(a):
    """Copy of ``a`` in page-locked host memory (torch is plumbing here, not compute)."""
    import torch
    a = np.asfortranarray(a)
    t = torch.empty(a.size, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
    v = t.numpy().reshape(a.shape, order="F")
    v[...] = a
    _PINNED.append(t)
    return v


_PINNED = []

# dram__bytes_read.sum + dram__bytes_write.sum of one k_kspmv_bell launch on the 1 M-vertex mesh (one GPU),
# from the committed ncu capture profiles/r1_kspmv_bell_ncu_full_summary.txt
NCU_TRAFFIC_BYTES_1M = 745554944   # k_kspmv_bell<1,4>: 711.08 MB read + 34.48 MB written (0.79 x the algorithmic 939.3 MB)


def cpu_reference_leg(mesh, C, ice, seconds, n_visc_full=None):
    """Time the oracle's restatement of the reference CPU path (test infrastructure used here
    as the reported baseline only): per-iteration closures + re-assembly + GMRES(30) with
    block-Jacobi/ILU(0) over `cores` strips (PETSc defaults, petsc_basic.f90:66-141)."""
    import copy
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    cores = min(os.cpu_count() or 1, 32)       # 32 = the authors' Snellius task count
    if not hasattr(mesh, "ops") or not mesh.ops:
        O.calc_all_matrix_operators_mesh(mesh)
    # grow the sample (number of Picard iterations of the same cold-start solve) until it
    # costs about `seconds`
    k, t_used, its_done, kry = 1, 0.0, 0, 0
    while True:
        C2 = copy.copy(C)
        C2.visc_it_nit = k - 1                 # loop exits when it > visc_it_nit
        D = O.new_DIVA_state(mesh)
        t0 = time.perf_counter()
        nv, na = O.solve_DIVA(mesh, ice, C2, D, "ksp", nranks=cores)
        t_used = time.perf_counter() - t0
        its_done, kry = nv, na
        if t_used > seconds / 2 or nv < k or k >= 64:
            break
        k = max(k + 1, int(k * min(4.0, seconds / max(t_used, 1e-3))))
    finished = its_done < k or (n_visc_full is not None and its_done >= n_visc_full)
    scale = 1.0 if finished or not n_visc_full else n_visc_full / its_done
    t_full = t_used * scale
    sample = (f"first {its_done} Picard iteration(s) of the same cold-start solve ({kry} GMRES its, "
              f"{t_used:.1f} s on {cores} threads)"
              + ("" if scale == 1.0 else f", extrapolated x{scale:.1f} to the {n_visc_full} Picard iterations the GPU solve took"))
    return {"value": 1.0 / t_full, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
            "seconds_sampled": t_used}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, C, ice, label = make_workload(args.workload)
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 2))):
        base = cpu_reference_leg(mesh, C, ice, max(5.0, args.cpu_seconds), n_visc_full=C.visc_it_nit + 1)
        vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": "DIVA velocity solves/sec", "value": v, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": label, "nV": mesh.nV, "nTri": mesh.nTri, "nz": mesh.nz},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference path (oracle port); the Fortran+PETSc reference cannot be built in this image"}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ufe_pkg
    ufe_pkg.load()
    from ufemism2_0_b200 import capi, diva

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (capi.ct.c_char * 128)()
            capi.check(capi.lib().ufe_comm_get_unique_id(buf))
            uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        comm = (rank, world, local, bytes(uid.cpu().tolist()))

    mesh, C, ice, label = make_workload(args.workload)
    t0 = time.perf_counter()
    S = diva.initialise_DIVA_solver(mesh, C, comm)
    t_create = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- leg 1: resident (value) ----------------
    S.upload(ice, state=True)
    infos = []
    for _ in range(args.warmup):
        S.reset_state_resident()
        S.solve_DIVA_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        S.reset_state_resident()
        info = S.solve_DIVA_resident()
        infos.append(info)
        dev_ms += info.ms_total
    barrier()
    wall_value = time.perf_counter() - w0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    value = args.steps / (dev_ms * 1e-3)

    # ---------------- several ranks: the same solve with the rows partitioned over the ranks, when the default left the
    # (small) system unpartitioned -- reported beside `value`, which is what the library does by default
    partitioned = None
    if world > 1 and infos[-1].reserved == 2:
        uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (capi.ct.c_char * 128)()
            capi.check(capi.lib().ufe_comm_get_unique_id(buf))
            uid2 = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid2, 0)
        os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"] = "0"
        SP = diva.initialise_DIVA_solver(mesh, C, (rank, world, local, bytes(uid2.cpu().tolist())))
        del os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"]
        SP.upload(ice, state=True)
        for _ in range(min(args.warmup, 3)):
            SP.reset_state_resident()
            SP.solve_DIVA_resident()
        barrier()
        p_ms, pi = 0.0, None
        for _ in range(args.steps):
            SP.reset_state_resident()
            pi = SP.solve_DIVA_resident()
            p_ms += pi.ms_total
        barrier()
        p_ms = max_over_ranks(p_ms)
        partitioned = {"value": args.steps / (p_ms * 1e-3), "unit": "solves/s", "ms_per_step": p_ms / args.steps,
                       "n_visc_its": pi.n_visc_its, "n_Axb_its": pi.n_Axb_its, "comm": COMM_LABEL.get(pi.reserved, str(pi.reserved)),
                       "what": "rows partitioned over the ranks by partition_list (UFE_REDUNDANT_MAX_UNKNOWNS=0), replicated exact factorisation"}
        SP.close()

    # ---------------- warm solve (time-stepping pattern): thickness perturbed by 0.1 %, state carried over
    import copy as _copy
    from ufemism2_0_b200 import synthetic as _syn
    ice_w = _copy.copy(ice)
    ice_w.Hi = ice.Hi * 1.001
    ice_w.Hs = _syn.ice_surface_elevation(ice_w.Hi, ice.Hb, ice.SL)
    ice_w.Hib = ice_w.Hs - ice_w.Hi
    warm_ms, warm_infos = 0.0, []
    for k in range(args.steps):
        S.upload(ice_w if k % 2 == 0 else ice, state=False)      # alternate so every warm solve sees a changed geometry
        wi = S.solve_DIVA_resident()
        warm_infos.append(wi)
        warm_ms += wi.ms_total
    warm_ms = max_over_ranks(warm_ms)
    warm = {"value": args.steps / (warm_ms * 1e-3), "unit": "solves/s", "ms_per_step": warm_ms / args.steps,
            "n_visc_its": [w.n_visc_its for w in warm_infos], "n_Axb_its": [w.n_Axb_its for w in warm_infos],
            "what": "second and later solve_DIVA calls from the previous velocities after a 0.1 % thickness change (device-resident)"}
    S.upload(ice, state=False)

    # ---------------- leg 2: end to end through ufe_diva_solve, pinned host buffers -------
    for n in ("Hi", "Hs", "Hib", "SL", "fraction_gr", "fraction_gr_b", "effective_pressure", "Ti",
              "till_friction_angle", "alpha_sq", "beta_sq", "mask_grounded_ice", "mask_floating_ice",
              "mask_icefree_land"):
        a = getattr(ice, n)
        setattr(ice, n, pinned_like(a.astype(np.int32) if n.startswith("mask") else a))
    state_names = ["u_vav_b", "v_vav_b", "tau_bx_b", "tau_by_b", "eta_3D_b", "u_base_b", "v_base_b", "u_3D_b",
                   "v_3D_b", "du_dx_a", "du_dy_a", "dv_dx_a", "dv_dy_a", "du_dz_3D_a", "dv_dz_3D_a", "eta_3D_a",
                   "basal_friction_coefficient_a"]
    for n in state_names:
        setattr(S, n, pinned_like(getattr(S, n)))
    nV, nT, nz = mesh.nV, mesh.nTri, mesh.nz
    h2d = 8 * (6 * nV + nT) + 4 * 3 * nV + 8 * 3 * nV + 8 * (6 * nT + nT * nz)
    if C.choice_ice_rheology_Glen == "Huybrechts1992":
        h2d += 8 * nV * nz
    d2h = 8 * (6 * nT + nT * nz) + 8 * 2 * nT * nz + 8 * 4 * nV + 8 * 2 * nV * nz + 8 * nV * nz + 8 * nV

    def e2e_step():
        for n in state_names[:7]:
            getattr(S, n)[...] = 0.0          # 'zero' initial velocities, host side
        return S.solve_DIVA(ice)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    e2e_infos = [e2e_step() for _ in range(args.steps)]
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - w0)
    e2e = {"value": args.steps / e2e_wall, "unit": "solves/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_wall / args.steps,
           "ms_h2d": float(np.mean([i.ms_h2d for i in e2e_infos])), "ms_d2h": float(np.mean([i.ms_d2h for i in e2e_infos]))}

    # ---------------- roofline: Krylov MatMult kernel on the resident stiffness matrix ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6500.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6500 GB/s (B200_PROFILING.md)"

    def spmv_roofline(solver, note):
        ms, nbytes = solver.bench_spmv(args.spmv_reps, flush_l2=True)
        ms = max_over_ranks(ms)
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "k_kspmv_bell (stiffness-matrix SpMV of the Krylov loop, blocked sliced-ELL)",
                "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms, "note": note}

    roofline_small = spmv_roofline(S, "bench workload's own matrix, per rank; L2 flushed (512 MiB memset) between launches; "
                                      "launch-latency-bound at this size")
    roofline = roofline_small
    krylov_iteration = other_kernels = None

    # ---------------- the thickness update that follows the velocity solve ----
    def thickness_leg(solver, msh, Cfg, geo, reps=5):
        """calc_dHi_dt_semiimplicit on the velocities the last solve left on the device (single rank)."""
        from ufemism2_0_b200 import mesh_types
        try:
            solver.set_mesh_edges(mesh_types.calc_mesh_edges(msh))
        except ValueError as e:
            return {"skipped": str(e)}
        n = msh.nV
        f = dict(Hi=geo.Hi, Hb=geo.Hb, SL=geo.SL, SMB=np.full(n, 0.3), BMB=np.zeros(n), LMB=np.zeros(n),
                 fraction_margin=np.ones(n), mask_noice=np.zeros(n, dtype=np.int32), dHi_dt_target=np.zeros(n))
        out, t = None, []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            out = solver.calc_dHi_dt_semiimplicit(f, 1.0)
            t.append(time.perf_counter() - t0)
        tm = solver.thickness_timing()
        dq = tm.pop("divq_algorithmic_bytes")
        return {"what": "calc_dHi_dt_semiimplicit, dt = 1 yr, f_s = %g, rtol %g, host buffers in and out" % (Cfg.dHi_semiimplicit_fs, Cfg.dHi_PETSc_rtol),
                "nV": n, "wall_ms_per_call": 1e3 * float(np.mean(t[1:])), "n_Axb_its": out["n_Axb_its"], "flags": out["flags"],
                "device_ms": tm, "k_thk_divq": {"algorithmic_bytes": dq, "achieved_GBs": dq / (tm["ms_divq"] * 1e-3) / 1e9,
                                                  "frac": dq / (tm["ms_divq"] * 1e-3) / 1e9 / peak,
                                                  "note": "includes the 1-thread init kernel launched before it"}}

    thickness = thickness_large = None
    if world == 1:
        thickness = thickness_leg(S, mesh, C, ice)
    if not args.no_large_roofline:
        # the configuration the north_star quotes the SpMV roofline on: ~1 M vertices, N ~ 4 M unknowns,
        # ~76 M non-zeros, partitioned over the ranks; one truncated Picard iteration assembles the matrix
        import copy
        from ufemism2_0_b200 import experiments
        meshL, CL, iceL = experiments.antarctic(args.large_vertices)
        CL = copy.copy(CL)
        CL.visc_it_nit, CL.b200_krylov_maxits, CL.b200_krylov_pc = 0, 20, "bjacobi2"
        commL = None
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                buf = (capi.ct.c_char * 128)()
                capi.check(capi.lib().ufe_comm_get_unique_id(buf))
                uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, 0)
            commL = (rank, world, local, bytes(uid.cpu().tolist()))
        SL = diva.initialise_DIVA_solver(meshL, CL, commL)
        SL.solve_DIVA(iceL, outputs=False)
        # whole Krylov iteration: a capped BiCGStab run on the same matrix; bytes per
        # iteration without fusion credit = 2 B_spmv + 16 vector passes of 8 B per unknown
        CK = copy.copy(CL)
        CK.b200_krylov_maxits = 200
        SL.set_config(CK)
        SL.reset_state_resident()
        ik = SL.solve_DIVA_resident()
        roofline = spmv_roofline(SL, f"synthetic Antarctic-scale mesh nV={meshL.nV} nTri={meshL.nTri} (N={2 * meshL.nTri} unknowns), rows "
                                     f"partitioned over {world} rank(s), figure per rank; L2 flushed (512 MiB memset) between launches")
        # DRAM bytes per launch of this kernel at this size on one GPU from `ncu --set full`
        # (profiles/r1_kspmv_bell_ncu_full_summary.txt); null when the sizes differ
        if world == 1 and args.large_vertices == 1_000_000:
            roofline["traffic"] = NCU_TRAFFIC_BYTES_1M
        # closures and assembly of one Picard iteration at this size (reported, not optimised this round);
        # bytes: SURVEY.md 8d B_asm = 44 nnz(M2) + 8 nnz(A) + 68 n_b (+ 8 nnz(A): the blocked copy is written too)
        nT_loc, nV_loc, nzL = meshL.nTri / world, meshL.nV / world, meshL.nz
        nnz_M2 = 10.0 * nT_loc
        asm_bytes = 44 * nnz_M2 + 2 * 8 * 4 * nnz_M2 + 68 * nT_loc
        clo_bytes = nV_loc * (6 * 28 + 6 * (6 + 3 * nzL) * 8 + (7 + 5 * nzL) * 8) + nT_loc * (3 * 28 + 3 * (3 + 3 * nzL) * 8 + (5 + 3 * nzL) * 8)
        other_kernels = {
            "assembly": {"ms": max_over_ranks(ik.ms_assembly), "algorithmic_bytes": asm_bytes,
                         "achieved_GBs": asm_bytes / (max_over_ranks(ik.ms_assembly) * 1e-3) / 1e9},
            "closures": {"ms": max_over_ranks(ik.ms_closures), "algorithmic_bytes": clo_bytes,
                         "achieved_GBs": clo_bytes / (max_over_ranks(ik.ms_closures) * 1e-3) / 1e9},
            "note": "k_assemble and k_vertex_diva + k_triangle_diva, one Picard iteration on the same large mesh, per rank"}
        if ik.n_Axb_its > 0:
            n_loc = 2 * meshL.nTri / world
            it_bytes = 2.0 * roofline["algorithmic_bytes_per_launch"] + 16 * 8.0 * n_loc
            it_ms = max_over_ranks(ik.ms_krylov / ik.n_Axb_its)
            krylov_iteration = {"method": "bicgstab+bjacobi2", "its_timed": ik.n_Axb_its, "ms_per_iteration": it_ms,
                                "algorithmic_bytes_per_iteration": it_bytes, "achieved": it_bytes / (it_ms * 1e-3) / 1e9,
                                "unit": "GB/s", "frac": it_bytes / (it_ms * 1e-3) / 1e9 / peak,
                                "note": "per rank, same mesh as `roofline`; includes the host polls between iteration batches"}
        if world == 1:
            thickness_large = thickness_leg(SL, meshL, CL, iceL, reps=2)
        SL.close()
        del meshL, iceL
    # reported beside the metric: a converged cold-start DIVA solve on a WIDE mesh (x-sorted bandwidth too large for the
    # banded exact preconditioner) with the multifrontal nested-dissection preconditioner, one GPU.  Never fatal.
    wide = None
    if world == 1 and args.wide_vertices > 0:
        try:
            import copy
            from ufemism2_0_b200 import experiments
            meshW, CW, iceW = experiments.antarctic(args.wide_vertices)
            CW = copy.copy(CW)
            CW.b200_krylov_pc, CW.b200_krylov_pc_lag, CW.b200_krylov_maxits = "nd_lu", 0, 200
            SW = diva.initialise_DIVA_solver(meshW, CW)
            t0 = time.time(); iw = SW.solve_DIVA(iceW); tw = time.time() - t0
            SW.close()
            wide = {"workload": f"synthetic Antarctic-shaped mesh nV={meshW.nV} (N={2 * meshW.nTri} unknowns), cold start, krylov_pc=nd_lu",
                    "solves_per_s": 1.0 / tw, "wall_s": tw, "n_visc_its": iw.n_visc_its, "n_Axb_its": iw.n_Axb_its, "flags": iw.flags,
                    "picard_converged": bool(iw.n_visc_its < CW.visc_it_nit), "ms_krylov": iw.ms_krylov,
                    "note": "host buffers, includes the once-per-mesh symbolic analysis inside the first linear solve"}
            del meshW, iceW
        except Exception as e:      # reported, the metric above does not depend on it
            wide = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        S.close()
        if world > 1:
            dist.destroy_process_group()
        return

    last = infos[-1]
    line = {
        "metric": "DIVA velocity solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "nV": nV, "nTri": nT, "nz": nz, "unknowns": 2 * nT,
                   "krylov": f"{C.b200_krylov_method}+{C.b200_krylov_pc}",
                   "krylov_pc_used": {0: "jacobi", 1: "bjacobi2", 2: "bjacobi_lu", 4: "nd_lu"}.get(last.krylov_pc_used, str(last.krylov_pc_used)),
                   "rtol": C.stress_balance_PETSc_rtol, "abstol": C.stress_balance_PETSc_abstol,
                   "picard_tol": C.visc_it_norm_dUV_tol, "visc_it_nit": C.visc_it_nit,
                   "l2": "step working set rewritten every Picard iteration; SpMV roofline leg flushes L2 (512 MiB) between launches"},
        "solve": {"n_visc_its": last.n_visc_its, "n_Axb_its": last.n_Axb_its, "flags": last.flags, "L2_uv": last.L2_uv,
                  "ms_closures": last.ms_closures, "ms_assembly": last.ms_assembly, "ms_krylov": last.ms_krylov,
                  "wall_ms_per_step": 1e3 * wall_value / args.steps, "create_s": t_create},
        "comm": ("single GPU" if world == 1 else COMM_LABEL.get(last.reserved, str(last.reserved))),
        "partitioned": partitioned,
        "e2e": e2e, "warm": warm, "gpu_launches": int(sum(i.gpu_launches for i in infos)),
        "roofline": roofline, "roofline_bench_workload": roofline_small, "krylov_iteration": krylov_iteration,
        "other_kernels_large_mesh": other_kernels, "thickness_update": thickness,
        "thickness_update_large_mesh": thickness_large, "wide_mesh_nd_lu": wide, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_leg(mesh, C, ice, args.cpu_seconds, n_visc_full=last.n_visc_its)
    print(json.dumps(line), flush=True)
    S.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
