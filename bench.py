#!/usr/bin/env python
"""Benchmark of the DIVA velocity solve (BASELINE.json metric: DIVA velocity solves/sec on
the MISMIP+ 2 km mesh; SpMV HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload mismipplus_2km|mismip_8km|ismip_hom_a|antarctic_1m|antarctic:<nV>]

A "step" is one full cold-start ``solve_DIVA`` call (Picard loop to ``visc_it_norm_dUV_tol``,
DIVA_main.f90:88-262) on synthetic geometry of the named mesh (Delaunay mesh, converging Picard
loop: ufemism2.0_b200/experiments.py).  Legs of the b200 arm:

  value : solves/s with every input resident in HBM (``ufe_diva_reset_state`` +
          ``ufe_diva_solve_resident``), device time (CUDA events inside the library), max over
          ranks, barrier + synchronize on both sides.  N > 1: the 64 000-unknown system does not
          shard usefully (every kernel is launch-latency-bound), so the default multi-GPU run is
          N independent replicas of the workload, one per GPU, no communication ("scaling":
          "weak"); the row-partitioned solve of the SAME system is reported beside it
          (``partitioned``) and the sharded path is measured where it matters, on the 1 M-vertex
          mesh (``antarctic_1m``: one converged solve partitioned over the N ranks, strong scaling);
  e2e   : the same solve through the reference-facing call ``ufe_diva_solve`` with pinned HOST
          buffers - H2D of all ice inputs + state and D2H of all results inside the timed region;
  parity: u, v of the timed workload against the oracle's direct-solve Picard loop (computed in a
          separate CPU process while the GPU legs run; outside every timed region);
  roofline : the Krylov MatMult kernel (k_kspmv_bell) on the stiffness matrix of the ~1 M-vertex
          mesh, CUDA events on the launching stream, L2 flushed between launches (ufe_bench_spmv);
  antarctic_1m : one CONVERGED cold-start solve_DIVA on the ~1 M-vertex Antarctic-shaped mesh with
          the multifrontal nested-dissection preconditioner, rows and elimination sub-trees
          partitioned over the N ranks;
  secondary : ISMIP-HOM A (L = 160 km): GPU and CPU port on a second workload;
  cpu_baseline : the oracle's restatement of the reference CPU path (GMRES(30) + block-Jacobi
          ILU(0), per-iteration re-assembly), the WHOLE solve, run to completion on the host
          cores (no extrapolation), rank 0, N = 1 only.

``--impl reference`` times only the CPU restatement (the reference itself - Fortran + MPI +
PETSc + NetCDF - cannot be built in this image; DESIGN.md "Oracle").
"""
import argparse
import copy
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
PC_LABEL = {0: "jacobi", 1: "bjacobi2", 2: "bjacobi_lu", 4: "nd_lu"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mismipplus_2km")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed workload")
    ap.add_argument("--no-secondary", action="store_true", help="skip the ISMIP-HOM A leg")
    ap.add_argument("--spmv-reps", type=int, default=200)
    ap.add_argument("--no-large", action="store_true",
                    help="skip the legs on the ~1M-vertex mesh (SpMV roofline, converged partitioned solve)")
    ap.add_argument("--large-vertices", type=int, default=1_000_000)
    ap.add_argument("--large-solves", type=int, default=1, help="timed converged solves on the large mesh")
    ap.add_argument("--oracle-worker", nargs=2, metavar=("WORKLOAD", "OUT"), help=argparse.SUPPRESS)
    return ap.parse_args()


def make_workload(name):
    import ufe_pkg
    ufe_pkg.load()
    from ufemism2_0_b200 import experiments
    if name == "mismipplus_2km":
        mesh, C, ice = experiments.MISMIPplus(2e3)
        label = "MISMIP+ 800x80 km, uniform 2 km synthetic Delaunay mesh (config_MISMIPplus_2km_spinup.cfg keys), cold-start DIVA solve"
    elif name == "mismipplus_8km":
        mesh, C, ice = experiments.MISMIPplus(8e3)
        label = "MISMIP+ 800x80 km, uniform 8 km synthetic Delaunay mesh, cold-start DIVA solve"
    elif name == "mismip_8km":
        mesh, C, ice = experiments.MISMIP_8km()
        label = "MISMIP 2000x2000 km, 8 km synthetic Delaunay mesh (config_MISMIP_8km_spinup_for_scaling.cfg keys), cold-start DIVA solve"
    elif name == "ismip_hom_a":
        mesh, C, ice = experiments.ISMIP_HOM("A", 160e3, 41)
        label = "ISMIP-HOM A, L = 160 km, 41x41 Delaunay mesh (config_ISMIP_HOM_A_160_DIVA.cfg keys), cold-start DIVA solve"
    elif name.startswith("antarctic"):
        nV = 1_000_000 if name == "antarctic_1m" else int(float(name.split(":")[1]))
        mesh, C, ice = experiments.antarctic(nV)
        label = f"synthetic Antarctic-scale dome, ~{nV} vertices, Delaunay mesh (config_ant_template.cfg keys), cold-start DIVA solve"
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


_PINNED = []


def pinned_like(a):
    """Copy of ``a`` in page-locked host memory (torch is plumbing here, not compute)."""
    import torch
    a = np.asfortranarray(a)
    t = torch.empty(a.size, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
    v = t.numpy().reshape(a.shape, order="F")
    v[...] = a
    _PINNED.append(t)
    return v


# dram__bytes_read.sum + dram__bytes_write.sum of one k_kspmv_bell launch on the 1 M-vertex mesh (one GPU),
# from the committed ncu capture profiles/r1_kspmv_bell_ncu_full_summary.txt (same kernel, same matrix shape)
NCU_TRAFFIC_BYTES_1M = 745554944   # k_kspmv_bell<1,4>: 711.08 MB read + 34.48 MB written (0.79 x the algorithmic 939.3 MB)


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle (test infrastructure) as checker and as the reported CPU baseline
# ------------------------------------------------------------------------------------------------------------------
def _oracle(native=False):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build(native=native)
    O.lib()
    return O


def oracle_worker(workload, out):
    """Separate CPU process: the oracle's direct-solve Picard loop of a workload (the parity reference)."""
    mesh, C, ice, _ = make_workload(workload)
    O = _oracle()
    O.calc_all_matrix_operators_mesh(mesh)
    D = O.new_DIVA_state(mesh)
    tr = []
    t0 = time.perf_counter()
    nv, _ = O.solve_DIVA(mesh, ice, C, D, "direct", trace=tr)
    np.savez(out, u=D["u_vav_b"], v=D["v_vav_b"], n_visc_its=nv, seconds=time.perf_counter() - t0,
             L2_uv=tr[-1][1] if tr else np.nan)


def parity_of(S, ref):
    """rel-L2 and max-norm distance of the GPU's u, v from the oracle's (both normalised by the oracle's velocity)."""
    du, dv = S.u_vav_b - ref["u"], S.v_vav_b - ref["v"]
    nrm = max(float(np.sqrt(np.sum(ref["u"] ** 2 + ref["v"] ** 2))), 1e-300)
    mx = max(float(np.max(np.hypot(ref["u"], ref["v"]))), 1e-300)
    return {"rel_L2": float(np.sqrt(np.sum(du ** 2 + dv ** 2)) / nrm), "max_norm": float(np.max(np.hypot(du, dv)) / mx),
            "n_visc_its_oracle": int(ref["n_visc_its"]), "oracle_seconds": float(ref["seconds"]),
            "oracle": "direct-solve Picard loop (scipy SuperLU per iteration), same mesh / inputs / config", "tolerance": 1e-6}


def cpu_reference_leg(mesh, C, ice):
    """The oracle's restatement of the reference CPU path, the WHOLE cold-start solve run to completion: per Picard
    iteration closures + re-assembly + GMRES(30) with block-Jacobi/ILU(0) over `cores` strips, zero initial guess,
    PETSc's default stopping rule (petsc_basic.f90:66-141).  Built with the reference's performance flags
    (-O3 -march=native, compile_UFEMISM.csh:86-98) on this machine."""
    O = _oracle(native=True)
    cores = min(os.cpu_count() or 1, 32)       # 32 = the authors' Snellius task count
    if not getattr(mesh, "ops", None):
        O.calc_all_matrix_operators_mesh(mesh)
    D = O.new_DIVA_state(mesh)
    tr = []
    t0 = time.perf_counter()
    nv, na = O.solve_DIVA(mesh, ice, C, D, "ksp", nranks=cores, trace=tr)
    t = time.perf_counter() - t0
    its = [x[2] for x in tr]
    conv = bool(tr and tr[-1][1] < C.visc_it_norm_dUV_tol)
    return {"value": 1.0 / t, "unit": "solves/s", "cores": cores, "kind": "port", "seconds": t,
            "sample": f"the whole cold-start solve, run to completion: {nv} Picard iterations, {na} GMRES(30)/bjacobi-ILU(0) iterations "
                      f"({min(its) if its else 0}-{max(its) if its else 0} per linear solve, none at the 10 000 cap: {all(i < 10000 for i in its)}), "
                      f"{t:.1f} s on {cores} threads; no extrapolation",
            "n_visc_its": int(nv), "n_Axb_its": int(na), "picard_converged": conv, "L2_uv": float(tr[-1][1]) if tr else None,
            "krylov_its_per_solve_max": int(max(its)) if its else 0, "extrapolated": False,
            "u": D["u_vav_b"], "v": D["v_vav_b"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, C, ice, label = make_workload(args.workload)
    t0 = time.perf_counter()
    base = cpu_reference_leg(mesh, C, ice)          # one whole solve is the bounded sample (tens of seconds)
    base.pop("u"); base.pop("v")
    v = base["value"]
    line = {"impl": "reference", "metric": "DIVA velocity solves/sec", "value": v, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": 1, "steps_requested": args.steps, "warmup": 0, "ms_per_step": 1e3 / v,
            "higher_is_better": True, "scaling": "weak" if args.gpus > 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": label, "nV": mesh.nV, "nTri": mesh.nTri, "nz": mesh.nz},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0,
            "note": "CPU restatement of the reference path (oracle port, -O3 -march=native), one complete solve on the host cores; "
                    "the Fortran+PETSc reference cannot be built in this image"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.oracle_worker:
        return oracle_worker(*args.oracle_worker)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import ufe_pkg
    ufe_pkg.load()
    from ufemism2_0_b200 import capi, diva, experiments
    from ufemism2_0_b200 import synthetic as _syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def new_comm():
        """NCCL unique id broadcast through torch.distributed -> (rank, nranks, device, id) for ufe_diva_create"""
        if world == 1:
            return None
        dev = torch.device("cuda", local)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = (capi.ct.c_char * 128)()
            capi.check(capi.lib().ufe_comm_get_unique_id(buf))
            uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        return (rank, world, local, bytes(uid.cpu().tolist()))

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the parity reference of the timed workload is computed by a CPU process while the GPU legs run
    want_parity = rank == 0 and not args.no_parity and not args.workload.startswith("antarctic")
    ora_proc = ora_out = None
    if want_parity:
        fd, ora_out = tempfile.mkstemp(suffix=".npz"); os.close(fd)
        ora_proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--oracle-worker", args.workload, ora_out],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)

    mesh, C, ice, label = make_workload(args.workload)
    big_workload = 2 * mesh.nTri > 131072            # large systems are one partitioned solve, small ones N replicas
    t0 = time.perf_counter()
    # replicas: a one-rank handle on THIS rank's GPU (comm = (rank 0 of 1, device))
    S = diva.initialise_DIVA_solver(mesh, C, new_comm() if big_workload else (0, 1, local, None))
    t_create = time.perf_counter() - t0

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
    units = args.steps * (1 if big_workload else world)          # replicas: every rank completed `steps` solves
    value = units / (dev_ms * 1e-3)
    last = infos[-1]
    S.download()
    uv_resident = (S.u_vav_b.copy(), S.v_vav_b.copy())

    # ---------------- several ranks, small system: the same solve with the rows partitioned over the ranks
    partitioned = multi_parity = None
    if world > 1 and not big_workload:
        os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"] = "0"
        SP = diva.initialise_DIVA_solver(mesh, C, new_comm())
        del os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"]
        SP.upload(ice, state=True)
        for _ in range(min(args.warmup, 2)):
            SP.reset_state_resident()
            SP.solve_DIVA_resident()
        barrier()
        p_ms, pi = 0.0, None
        for _ in range(min(args.steps, 3)):
            SP.reset_state_resident()
            pi = SP.solve_DIVA_resident()
            p_ms += pi.ms_total
        barrier()
        p_ms = max_over_ranks(p_ms) / min(args.steps, 3)
        SP.download()
        du, dv = SP.u_vav_b - uv_resident[0], SP.v_vav_b - uv_resident[1]
        nrm = max(float(np.sqrt(np.sum(uv_resident[0] ** 2 + uv_resident[1] ** 2))), 1e-300)
        partitioned = {"value": 1e3 / p_ms, "unit": "solves/s", "ms_per_step": p_ms, "n_visc_its": pi.n_visc_its, "n_Axb_its": pi.n_Axb_its,
                       "flags": pi.flags, "krylov_pc_used": PC_LABEL.get(pi.krylov_pc_used, str(pi.krylov_pc_used)),
                       "comm": COMM_LABEL.get(pi.reserved, str(pi.reserved)),
                       "rel_L2_vs_one_gpu": float(np.sqrt(np.sum(du ** 2 + dv ** 2)) / nrm),
                       "what": "ONE solve of the same system, rows and elimination sub-trees partitioned over the ranks (partition_list strips, "
                               "UFE_REDUNDANT_MAX_UNKNOWNS=0); strong scaling of a launch-latency-bound system, reported as is"}
        SP.close()

    # ---------------- warm solve (time-stepping pattern): thickness perturbed by 0.1 %, state carried over
    ice_w = copy.copy(ice)
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
    warm = {"value": units / (warm_ms * 1e-3), "unit": "solves/s", "ms_per_step": warm_ms / args.steps,
            "n_visc_its": [w.n_visc_its for w in warm_infos], "n_Axb_its": [w.n_Axb_its for w in warm_infos],
            "flags": [w.flags for w in warm_infos],
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
    e2e = {"value": units / e2e_wall, "unit": "solves/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_wall / args.steps,
           "ms_h2d": float(np.mean([i.ms_h2d for i in e2e_infos])), "ms_d2h": float(np.mean([i.ms_d2h for i in e2e_infos]))}
    uv_e2e = (np.array(S.u_vav_b), np.array(S.v_vav_b))

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

    # ---------------- several ranks: parity of the partitioned path on a case the oracle finishes in seconds
    if world > 1:
        meshP, CP, iceP = experiments.MISMIPplus(8e3)
        os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"] = "0"
        SQ = diva.initialise_DIVA_solver(meshP, CP, new_comm())
        del os.environ["UFE_REDUNDANT_MAX_UNKNOWNS"]
        iq = SQ.solve_DIVA(iceP)
        if rank == 0:
            O = _oracle()
            O.calc_all_matrix_operators_mesh(meshP)
            D = O.new_DIVA_state(meshP)
            nvq, _ = O.solve_DIVA(meshP, iceP, CP, D, "direct")
            multi_parity = parity_of(SQ, {"u": D["u_vav_b"], "v": D["v_vav_b"], "n_visc_its": nvq, "seconds": 0.0})
            multi_parity.update(workload="MISMIP+ 8 km, rows partitioned over the ranks", n_visc_its=iq.n_visc_its, flags=iq.flags,
                                krylov_pc_used=PC_LABEL.get(iq.krylov_pc_used, str(iq.krylov_pc_used)))
        SQ.close()
        barrier()

    # ---------------- the ~1 M-vertex mesh: SpMV roofline (north_star target) and one converged partitioned solve
    large = krylov_iteration = other_kernels = None
    if not args.no_large:
        tl0 = time.perf_counter()
        meshL, CL, iceL = experiments.antarctic(args.large_vertices)
        t_mesh = time.perf_counter() - tl0
        SL = diva.initialise_DIVA_solver(meshL, CL, new_comm())
        SL.upload(iceL, state=True)
        barrier()
        tl0 = time.perf_counter()
        i0 = SL.solve_DIVA_resident()                          # first solve: includes the once-per-mesh symbolic analysis
        t_first = max_over_ranks(time.perf_counter() - tl0)
        l_ms, li = 0.0, i0
        for _ in range(args.large_solves):
            SL.reset_state_resident()
            barrier()
            li = SL.solve_DIVA_resident()
            l_ms += li.ms_total
        barrier()
        l_ms = max_over_ranks(l_ms) / max(args.large_solves, 1)
        large = {"workload": f"synthetic Antarctic-scale dome, nV={meshL.nV} nTri={meshL.nTri} (N={2 * meshL.nTri} unknowns), Delaunay mesh, "
                             f"config_ant_template.cfg keys, cold start, ONE solve partitioned over {world} rank(s)",
                 "value": 1e3 / l_ms, "unit": "solves/s", "scaling": "strong", "ms_per_solve": l_ms, "n_visc_its": li.n_visc_its,
                 "n_Axb_its": li.n_Axb_its, "flags": li.flags, "picard_converged": bool(li.flags == 0), "L2_uv": li.L2_uv,
                 "krylov_pc_used": PC_LABEL.get(li.krylov_pc_used, str(li.krylov_pc_used)),
                 "ms_closures": max_over_ranks(li.ms_closures), "ms_assembly": max_over_ranks(li.ms_assembly),
                 "ms_krylov_incl_factorisation": max_over_ranks(li.ms_krylov),
                 "first_solve_s_incl_symbolic_analysis": t_first, "mesh_generation_s": t_mesh,
                 "comm": "single GPU" if world == 1 else COMM_LABEL.get(li.reserved, str(li.reserved))}
        roofline = spmv_roofline(SL, f"synthetic Antarctic-scale mesh nV={meshL.nV} nTri={meshL.nTri} (N={2 * meshL.nTri} unknowns), rows "
                                     f"partitioned over {world} rank(s), figure per rank; L2 flushed (512 MiB memset) between launches")
        if world == 1 and args.large_vertices == 1_000_000:
            roofline["traffic"] = NCU_TRAFFIC_BYTES_1M
            roofline["traffic_source"] = "ncu --set full capture of this kernel on this matrix shape, profiles/r1_kspmv_bell_ncu_full_summary.txt"
        # closures and assembly of one Picard iteration at this size; bytes: SURVEY.md 8d
        nT_loc, nV_loc, nzL = meshL.nTri / world, meshL.nV / world, meshL.nz
        nnz_M2 = 10.0 * nT_loc
        asm_bytes = 44 * nnz_M2 + 2 * 8 * 4 * nnz_M2 + 68 * nT_loc
        clo_bytes = nV_loc * (6 * 28 + 6 * (6 + 3 * nzL) * 8 + (7 + 5 * nzL) * 8) + nT_loc * (3 * 28 + 3 * (3 + 3 * nzL) * 8 + (5 + 3 * nzL) * 8)
        nit = max(li.n_visc_its, 1)
        other_kernels = {
            "assembly": {"ms_per_picard_iteration": max_over_ranks(li.ms_assembly) / nit, "algorithmic_bytes": asm_bytes,
                         "achieved_GBs": asm_bytes / (max_over_ranks(li.ms_assembly) / nit * 1e-3) / 1e9},
            "closures": {"ms_per_picard_iteration": max_over_ranks(li.ms_closures) / nit, "algorithmic_bytes_gather_counted": clo_bytes,
                         "achieved_GBs_gather_counted": clo_bytes / (max_over_ranks(li.ms_closures) / nit * 1e-3) / 1e9},
            "note": "k_assemble and k_vertex_diva + k_triangle_diva per Picard iteration of the large solve, per rank"}
        SL.close()
        del meshL, iceL

    # ---------------- second workload: ISMIP-HOM A, GPU and CPU port on the same inputs (rank 0, one GPU)
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and args.workload != "ismip_hom_a":
        meshS, CS, iceS, labelS = make_workload("ismip_hom_a")
        SS = diva.initialise_DIVA_solver(meshS, CS, (0, 1, local, None))
        SS.upload(iceS, state=True)
        for _ in range(2):
            SS.reset_state_resident(); SS.solve_DIVA_resident()
        s_ms, si = 0.0, None
        for _ in range(args.steps):
            SS.reset_state_resident(); si = SS.solve_DIVA_resident(); s_ms += si.ms_total
        SS.download()
        cb = cpu_reference_leg(meshS, CS, iceS)
        O = _oracle()
        D = O.new_DIVA_state(meshS)
        nvs, _ = O.solve_DIVA(meshS, iceS, CS, D, "direct")
        secondary = {"workload": labelS, "nTri": meshS.nTri, "value": args.steps / (s_ms * 1e-3), "unit": "solves/s",
                     "n_visc_its": si.n_visc_its, "n_Axb_its": si.n_Axb_its, "flags": si.flags,
                     "parity": parity_of(SS, {"u": D["u_vav_b"], "v": D["v_vav_b"], "n_visc_its": nvs, "seconds": 0.0}),
                     "cpu_baseline": {k: v for k, v in cb.items() if k not in ("u", "v")}}
        SS.close()

    if rank != 0:
        S.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- parity of the timed workload (oracle process started at the beginning) -------------
    parity = None
    if ora_proc is not None:
        err = ora_proc.communicate()[1]
        if ora_proc.returncode == 0:
            ref = np.load(ora_out)
            class _V:  # noqa: E701
                pass
            a, b = _V(), _V()
            a.u_vav_b, a.v_vav_b = uv_resident
            b.u_vav_b, b.v_vav_b = uv_e2e
            parity = parity_of(a, ref)
            parity["e2e_rel_L2"] = parity_of(b, ref)["rel_L2"]
            parity["n_visc_its"] = last.n_visc_its
            parity["pass"] = bool(parity["rel_L2"] < 1e-6 and parity["max_norm"] < 1e-6 and parity["e2e_rel_L2"] < 1e-6)
        else:
            parity = {"error": err.decode(errors="replace")[-400:]}
        try:
            os.unlink(ora_out)
        except OSError:
            pass

    line = {
        "metric": "DIVA velocity solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if (big_workload or world == 1) else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": label + ("" if (big_workload or world == 1) else f"; {world} independent replicas, one per GPU, no communication"),
                   "nV": nV, "nTri": nT, "nz": nz, "unknowns": 2 * nT,
                   "krylov": f"{C.b200_krylov_method}+{C.b200_krylov_pc}",
                   "krylov_pc_used": PC_LABEL.get(last.krylov_pc_used, str(last.krylov_pc_used)),
                   "rtol": C.stress_balance_PETSc_rtol, "abstol": C.stress_balance_PETSc_abstol,
                   "picard_tol": C.visc_it_norm_dUV_tol, "visc_it_nit": C.visc_it_nit,
                   "l2": "step working set (fronts, matrix, fields) rewritten every Picard iteration; SpMV roofline leg flushes L2 (512 MiB) between launches"},
        "solve": {"n_visc_its": last.n_visc_its, "n_Axb_its": last.n_Axb_its, "flags": last.flags, "picard_converged": bool(last.flags == 0),
                  "L2_uv": last.L2_uv, "ms_closures": last.ms_closures, "ms_assembly": last.ms_assembly, "ms_krylov_incl_factorisation": last.ms_krylov,
                  "wall_ms_per_step": 1e3 * wall_value / args.steps, "create_s": t_create},
        "parity": parity, "multi_gpu_parity": multi_parity,
        "comm": ("single GPU" if world == 1 else ("independent replicas" if not big_workload else COMM_LABEL.get(last.reserved, str(last.reserved)))),
        "partitioned": partitioned,
        "e2e": e2e, "warm": warm, "gpu_launches": int(sum(i.gpu_launches for i in infos)),
        "roofline": roofline, "roofline_bench_workload": roofline_small, "antarctic_1m": large,
        "other_kernels_large_mesh": other_kernels, "secondary": secondary, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_leg(mesh, C, ice)
        ref_u, ref_v = cb.pop("u"), cb.pop("v")
        nrm = max(float(np.sqrt(np.sum(ref_u ** 2 + ref_v ** 2))), 1e-300)
        cb["rel_L2_of_gpu_vs_this_cpu_solve"] = float(np.sqrt(np.sum((uv_resident[0] - ref_u) ** 2 + (uv_resident[1] - ref_v) ** 2)) / nrm)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    S.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
