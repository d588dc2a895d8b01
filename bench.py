#!/usr/bin/env python
"""Benchmark of the LTE time-step hot path (BASELINE.json metric: LTE timesteps/s at 655,362 cells, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One bench "step" = one output interval of the reference's loop: --substeps (default 100) consecutive LTE
time steps (ab3Explicit dumps every totalIter/outputTime steps; the shipped input.in gives 290). `value`
counts LTE time steps per second with everything resident in HBM; `e2e` is the same interval driven
through the C ABI with HOST buffers: state H2D (odis_set_state), the interval's steps, and the D2H reads a
dump needs (eta, edge velocities, dissipation). Under torchrun (N > 1) the SAME grid is cut into N
contiguous space-filling-curve parts, one per GPU, with a one-ring halo; the boundary-edge velocities are
exchanged once per step by direct stores into the neighbours' memory over NVLink, issued by the edge kernel
itself (BASELINE config 3: "1/2/4/8 B200 face-partitioned"): total work is fixed, so "scaling" is "strong". NCCL is used only for the barrier / max-over-ranks timing and to
pass the IPC handles around.

`--impl reference` times the reference's own CPU solver (oracle/_ref/odis_ref_l<L>: the unmodified
reference sources) on the same workload, on rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LTE timesteps/sec at grid L8 (FP64), 1-8 B200; achieved HBM GB/s vs peak"
UNIT = "timesteps/s"

# Enceladus subsurface ocean (literature values, not in the reference)
ENCELADUS = dict(radius=252.1e3, shell=23e3, h=38e3, g=0.113, omega=5.307e-5, ecc=0.0047, alpha=1e-7, love_reduct=0.9)
# Shell pressure coefficients beta_l, l = 0..8, for a 23 km shell: column 23 of the reference's
# input_files/LOVE_SHELL_COEFFS/ENCELADUS/beta_hs_1km_to_50km_lmax30.txt (row l; the reference reads one column of it as
# input_files/beta.txt, src/boundaryConditions.cpp:139-158). The term applies factor_l = 1 - beta_l (:373).
BETA_23KM = [0.0, 0.0, 2.970754525850653494e+01, 3.846475509963270412e+01, 4.902765616416872518e+01, 6.693284467155693562e+01,
             9.640980297755693584e+01, 1.414891137087704465e+02, 2.060167851185478298e+02]


def shell_factor(l_max: int) -> np.ndarray:
    f = 1.0 - np.array(BETA_23KM[:l_max + 1])
    f[:2] = 0.0                                   # degrees 0, 1 are never applied
    return f


def workload_params(mesh, member: int = 0, n_members: int = 1) -> dict:
    """Solver scalars for sweep member `member` (ocean thickness x drag, log-spaced as in SURVEY §8d item 5)."""
    h, alpha = ENCELADUS["h"], ENCELADUS["alpha"]
    if n_members > 1:
        h = float(np.logspace(np.log10(10e3), np.log10(38e3), n_members)[member])
        alpha = float(np.logspace(-8, -6, n_members)[member])
    dmin = float(mesh.tables["face_node_dist"].min())
    # dt from the wave CFL rule in the reference's (commented) code, src/mesh.cpp:1593-1594, on the thickest ocean
    dt = 0.2 * dmin / math.sqrt(ENCELADUS["g"] * ENCELADUS["h"])
    return dict(g=ENCELADUS["g"], h=h, alpha=alpha, dt=dt, radius=mesh.radius, omega=ENCELADUS["omega"],
                love_reduct=ENCELADUS["love_reduct"], ecc=ENCELADUS["ecc"], obl=0.0, shell_thickness=ENCELADUS["shell"],
                semimajor_axis=238.02e6, potential=5, friction=0, surface=2, init_load=0, reorder=1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        if shutil.which("nvidia-smi"):
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                              "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read, daemon=True)
                self.t.start()
            except OSError:
                self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch() -> float | None:
    """dram bytes per edge_step launch from the committed ncu summary, if one exists."""
    p = os.path.join(ROOT, "profiles", "edge_step_summary.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


def cpu_baseline_port(mesh, prm: dict, budget_s: float = 15.0, sh_degree: int = 0) -> dict:
    """The oracle's plain-C restatement of the reference loop (bit-identical to the reference solver, see
    tests/test_oracle_pinned.py) timed on one host core on a bounded sample of the same workload."""
    from oracle.lte_oracle import LteOracle
    keys = ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl", "shell_thickness", "potential", "friction", "surface", "init_load")
    o = LteOracle(mesh.tables, {k: prm[k] for k in keys})
    if sh_degree >= 2:
        from oracle import sh_oracle
        Y = sh_oracle.basis(mesh.tables["node_pos_sph"], sh_degree)
        o.set_self_gravity(Y, sh_oracle.apply_operator(Y, shell_factor(sh_degree)))
    o.set_state()
    t0 = time.perf_counter(); o.step(3); probe = (time.perf_counter() - t0) / 3
    n = int(max(5, min(2000, budget_s / max(probe, 1e-9))))
    t0 = time.perf_counter(); o.step(n); el = time.perf_counter() - t0
    return {"value": n / el, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} LTE steps of the same {mesh.n_cells}-cell workload"
                      + (f" (self-gravity term to degree {sh_degree} included)" if sh_degree >= 2 else "") + ", oracle/lte_oracle.c (gcc -O2), single thread"}


def variant_probe(args) -> None:
    """Child process of the N = 1 bench (own CUDA context, so a fault in an opt-in kernel selection cannot take the headline
    measurement with it): times one selection and checks it against the default selection; prints one JSON line."""
    import geodesicodis_b200 as odis
    name = args.variant_probe
    out = {"probe": name}
    if name == "headline_selections":
        pos, fr, cen = odis.generate_grid(args.level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"] - ENCELADUS["shell"])
        prm, L, S = workload_params(mesh), max(args.sh_degree, 2), args.substeps
        ref_eta = None
        # selections that reuse validated synchronisation first; the new mbarrier byte accounting (16-bit ids) last
        for key, sel in (("default", 0), ("cell_update_64_registers", 64), ("cell_update_l2_prefetch", 512),
                         ("cell_update_l2_prefetch_64_registers", 576), ("self_gravity_3_launch", 16), ("self_gravity_3_launch_64_registers", 80),
                         ("self_gravity_3_launch_cell_l2_prefetch", 16 + 512), ("edge_ids_16bit", 128), ("edge_ids_16bit_cell_l2_prefetch", 640),
                         ("self_gravity_3_launch_edge_ids_16bit_cell_l2_prefetch", 16 + 128 + 512)):
            sv = odis.Solver(mesh, dict(prm, kernel_select=sel))
            sv.enable_self_gravity(L, shell_factor(L))
            sv.step(2 * S)
            eta = sv.field(odis.FIELD_ETA)
            if ref_eta is None:
                ref_eta = eta
            n = 10 * S
            l0 = sv.launches
            rec = {"timesteps_per_s": round(n / (sv.step_timed(n) * 1e-3), 1), "launches_per_step": (sv.launches - l0) / n,
                   "max_rel_diff_eta_vs_default_after_%d_steps" % (2 * S): float(np.abs(eta - ref_eta).max() / np.abs(ref_eta).max())}
            e, c, g = sv.step_profiled_sh(200)
            rec["avg_us"] = {"edge": round(e / 200 * 1e3, 2), "cell": round(c / 200 * 1e3, 2), "self_gravity_launches": round(g / 200 * 1e3, 2)}
            out[key] = rec
            print(json.dumps(out), flush=True)                   # the parent reads the last complete line
            sv.close()
    elif name == "e2e_pipelined":
        # the e2e interval of the main line (state H2D, S steps, eta / v / dissipation D2H, all through the C ABI with page-locked host
        # buffers) with the copies on the second stream: odis_stage_state / odis_commit_state bring interval k+1's state in while
        # interval k steps, odis_snapshot_begin / _wait take interval k's fields out while interval k+1 steps. Same bytes per interval.
        import torch
        pos, fr, cen = odis.generate_grid(args.level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"] - ENCELADUS["shell"])
        prm, L, S = workload_params(mesh), args.sh_degree, args.substeps
        N, F = mesh.n_cells, mesh.n_edges
        sv = odis.Solver(mesh, prm)
        if L >= 2:
            sv.enable_self_gravity(L, shell_factor(L))
        sv.step(2 * S)
        def pin(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            return torch.from_numpy(a).pin_memory().numpy() if torch.cuda.is_available() else a      # (no CUDA: the emulated library in tests)
        h_v, h_eta = pin(sv.field(odis.FIELD_VELOCITY)), pin(sv.field(odis.FIELD_ETA))
        h_dv, h_de = pin(sv.field(odis.FIELD_DVDT).ravel()), pin(sv.field(odis.FIELD_DETADT).ravel())
        it0, Ke, fields = sv.iter, 20, sv.SNAP_ETA | sv.SNAP_VELOCITY

        def run(n):
            check = 0.0
            sv.stage_state(h_v, h_eta, h_dv, h_de)
            for k in range(n):
                sv.commit_state(iter=it0 + k * S)
                if k + 1 < n:
                    sv.stage_state(h_v, h_eta, h_dv, h_de)
                sv.step(S)
                sv.snapshot_begin(k & 1, fields)
                if k > 0:
                    check += sv.snapshot_wait((k - 1) & 1, copy=False)["dissipation_avg"]
            last = sv.snapshot_wait((n - 1) & 1, copy=False)
            sv.synchronize()
            return check + last["dissipation_avg"], last

        run(2)
        t0 = time.perf_counter()
        _, last = run(Ke)
        el = time.perf_counter() - t0
        # the same interval, synchronous calls, for the comparison and as the check of the pipelined result
        sv.set_state(h_v, h_eta, h_dv, h_de, iter=it0 + (Ke - 1) * S)
        sv.step(S)
        same = bool(np.array_equal(sv.field(odis.FIELD_ETA), last["eta"]) and np.array_equal(sv.field(odis.FIELD_VELOCITY), last["velocity"]))
        out.update({"value": round(Ke * S / el, 2), "unit": UNIT, "intervals_timed": Ke, "h2d_bytes_per_step": 8 * (4 * F + 4 * N),
                    "d2h_bytes_per_step": 8 * (F + N + 1), "fields_identical_to_synchronous_calls": same,
                    "note": "pipeline fill (the first, unoverlapped upload) is inside the timed region"})
    elif name == "nonlinear":
        level = args.probe_level                                 # 8: 163,842 cells (BASELINE 'L7'), the shipped input.in physics
        pos, fr, cen = odis.generate_grid(level)
        r = 6.37122e6
        mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
        nl = odis.nonlinear_tables(mesh, 0.5)
        dmin = float(mesh.tables["face_node_dist"].min())
        prm = dict(g=9.80616, h=8e3, alpha=1e-7, dt=0.2 * dmin / math.sqrt(9.80616 * 8e3), radius=r, omega=7.292e-5, love_reduct=1.0, ecc=0.01,
                   obl=math.radians(-2.0), shell_thickness=0.0, semimajor_axis=0.0, potential=1, friction=0, surface=0, init_load=0, reorder=1)
        res = {}
        for key, sel in (("6_launch_default", 0), ("4_launch", 32)):
            sv = odis.Solver(mesh, dict(prm, kernel_select=sel))
            sv.enable_advection(nl)
            sv.step(60)
            res[key] = sv.field(odis.FIELD_ETA)
            out[key + "_timesteps_per_s"] = round(400 / (sv.step_timed(400) * 1e-3), 1)
            sv.close()
        out["cells"] = mesh.n_cells
        out["fields_identical"] = bool(np.array_equal(res["6_launch_default"], res["4_launch"]))
    elif name == "other_configs":
        # the remaining single-GPU shapes of BASELINE.json's configs, each a short device-resident timing (kernels that have run on B200s
        # before): [1] Enceladus free-surface ocean, 163,842 cells, linear drag, no self-gravity; [4] one GPU's share of the ensemble
        # sweep: 32 members (ocean thickness x drag) on 40,962 cells, self-gravity to degree 8 as batched FP64 tensor-core GEMMs
        pos, fr, cen = odis.generate_grid(args.probe_level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"])
        prm = dict(workload_params(mesh), surface=0, shell_thickness=0.0, love_reduct=1.0)
        sv = odis.Solver(mesh, prm)
        sv.step(200)
        ms = sv.step_timed(2000) / 2000
        _, alg = sv.footprint()
        out["free_surface_%d_cells" % mesh.n_cells] = {"timesteps_per_s": round(1e3 / ms, 1), "cell_updates_per_s": round(mesh.n_cells * 1e3 / ms, 1),
                                            "algorithmic_GBps": round(alg / (ms * 1e-3) / 1e9, 1)}
        print(json.dumps(out), flush=True)
        sv.close()
        pos, fr, cen = odis.generate_grid(args.probe_level - 1)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"] - ENCELADUS["shell"])
        M = 32
        plist = [workload_params(mesh, m, M) for m in range(M)]
        ens = odis.Ensemble(mesh, plist)
        ens.enable_self_gravity(8, shell_factor(8))
        ens.step(40)
        ms = ens.step_timed(400) / 400
        info = ens.info()
        out["ensemble_32_members_%d_cells_sh8" % mesh.n_cells] = {"batched_steps_per_s": round(1e3 / ms, 1), "member_steps_per_s": round(M * 1e3 / ms, 1),
                                                      "algorithmic_GBps": round(info["algorithmic_bytes_per_step"] / (ms * 1e-3) / 1e9, 1)}
        ens.close()
    else:
        out["error"] = "unknown probe"
    print(json.dumps(out), flush=True)


PROBE_BUDGET_S = 420.0          # all child-process probes of one bench run together (each also has its own limit)
_probe_deadline = [None]


def run_probe(name: str, args, limit: float = 180.0) -> dict:
    """Runs `bench.py --variant-probe name` in a subprocess; any failure is recorded instead of raised. A probe that does not return
    (a kernel selection that has never run on this hardware) is killed at its limit and what it had printed until then is kept."""
    if _probe_deadline[0] is None:
        _probe_deadline[0] = time.time() + PROBE_BUDGET_S
    limit = min(limit, _probe_deadline[0] - time.time())
    if limit < 20.0:
        return {"probe": name, "error": "skipped: the probes' time budget of this bench run is spent"}
    cmd = [sys.executable, os.path.abspath(__file__), "--variant-probe", name, "--level", str(args.level), "--sh-degree", str(args.sh_degree),
           "--substeps", str(args.substeps)]
    last_line = lambda text: ([l for l in (text or "").splitlines() if l.startswith("{") and l.endswith("}")] or [None])[-1]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=limit,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = last_line(r.stdout)
        res = json.loads(line) if line else {"probe": name}
        if r.returncode != 0 or not line:                           # keep what was measured before the failure
            res["error"] = f"exit {r.returncode}: {(r.stderr or r.stdout)[-300:]}"
        return res
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else e.stdout
        try:
            res = json.loads(last_line(out)) if last_line(out) else {"probe": name}
        except Exception:
            res = {"probe": name}
        res["error"] = f"killed after {limit:.0f} s; the entries above had been measured by then"
        return res
    except Exception as e:                                           # spawn failure, bad JSON
        return {"probe": name, "error": repr(e)[:300]}


def run_ours(args) -> None:
    import torch
    import geodesicodis_b200 as odis

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: geodesicodis_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    radius = ENCELADUS["radius"] - ENCELADUS["shell"]                 # LID_LOVE: boundaryConditions.cpp:126
    if world > 1 and args.level >= 10:
        # large grids: the node's first rank builds the tables once, the others map them (np.load mmap) from /dev/shm
        shared = f"/dev/shm/odis_b200_mesh_l{args.level}_{os.environ.get('MASTER_PORT', '0')}"
        if local_rank == 0:
            pos, fr, cen = odis.generate_grid(args.level)
            odis.Mesh.from_arrays(pos, fr, cen, radius).save(shared)
            del pos, fr, cen
        dist.barrier()
        mesh = odis.Mesh.load(shared)
    else:
        pos, fr, cen = odis.generate_grid(args.level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, radius)
    prm = workload_params(mesh)
    if args.kernel_select:                          # experiments only (odis_params.reserved[0]); the default line is selection 0
        prm = dict(prm, kernel_select=args.kernel_select)
    solver = odis.Solver(mesh, prm, device=local_rank, rank=rank, world=world)
    if world > 1:                                   # every rank publishes its halo buffers; neighbours map them
        blobs = [None] * world
        dist.all_gather_object(blobs, solver.halo_blob())
        solver.halo_connect(blobs)
        dist.barrier()
    L = args.sh_degree
    if L >= 2:                                      # self-gravity / shell-pressure term (BASELINE config 3), matrix-free kernels
        solver.enable_self_gravity(L, shell_factor(L))
    N, F = mesh.n_cells, mesh.n_edges
    S, K, W = args.substeps, args.steps, max(args.warmup, 3)
    dev_bytes, alg_bytes = solver.footprint()

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(W):
        solver.step(S)
    barrier()
    launches0 = solver.launches
    sampler = ClockSampler(local_rank)
    t_wall0 = time.time()
    ms = solver.step_timed(K * S)               # CUDA events on the solver's own stream
    torch.cuda.synchronize()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = solver.launches - launches0
    if not math.isfinite(solver.dissipation_avg()):
        raise SystemExit("bench.py: the run blew up (non-finite dissipation); the timing would be meaningless")
    barrier()
    ms = max_over_ranks(ms)
    value = K * S / (ms * 1e-3)                  # all ranks advance the same K*S steps of the one global grid

    # ---- per-kernel timing for the roofline (live, CUDA events around every launch) --------------
    nprof = min(K * S, 400)
    edge_ms, cell_ms, sh_ms = solver.step_profiled_sh(nprof)
    edge_us, cell_us, sh_us = edge_ms / nprof * 1e3, cell_ms / nprof * 1e3, sh_ms / nprof * 1e3
    peak, peak_src = measured_peak()
    part = solver.partition()
    edge_alg = 200 * part["own_edges"]          # per-edge algorithmic bytes x edges per launch (this rank's)
    achieved = edge_alg / (edge_us * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": "edge_step_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic_per_launch(), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": edge_alg, "avg_launch_us": round(edge_us, 2),
                "cell_step_kernel": {"algorithmic_bytes_per_launch": 128 * part["own_cells"], "avg_launch_us": round(cell_us, 2),
                                     "achieved": round(128 * part["own_cells"] / (cell_us * 1e-6) / 1e9, 1)},
                "self_gravity_kernels": {"avg_us_per_step": round(sh_us, 2), "launches_per_step": 3 if L >= 2 else 0, "bound": "fp64 pipe (matrix-free: "
                                         "the harmonic basis is rebuilt per cell by recurrence instead of streaming 8*(l_max+1)^2 B per cell)"},
                "whole_step": {"algorithmic_bytes": alg_bytes, "achieved": round(alg_bytes * value / 1e9, 1),
                               "frac": round(alg_bytes * value / 1e9 / peak, 4), "frac_of_8TBs_nominal": round(alg_bytes * value / 8e12, 4),
                               "note": "per GPU: this rank's share of the grid; for N>1 the halo exchange is part of the two kernels"}}

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    pin = lambda n: torch.zeros(n, dtype=torch.float64).pin_memory().numpy()
    h_v, h_eta, h_dv, h_de = pin(F), pin(N), pin(3 * F), pin(3 * N)
    def whole(field):                            # a partitioned solver returns its own entries, zeros elsewhere
        a = solver.field(field)
        if dist is not None:
            t = torch.from_numpy(a).cuda()
            dist.all_reduce(t)
            a = t.cpu().numpy()
        return a

    h_v[:] = whole(odis.FIELD_VELOCITY); h_eta[:] = whole(odis.FIELD_ETA)
    h_dv[:] = whole(odis.FIELD_DVDT).ravel(); h_de[:] = whole(odis.FIELD_DETADT).ravel()
    it0 = solver.iter
    Ke = max(1, min(K, 20))

    def e2e_interval(k: int):
        solver.set_state(h_v, h_eta, h_dv, h_de, iter=it0 + k * S)          # H2D of the interval's inputs
        solver.step(S)
        solver.field(odis.FIELD_ETA, out=h_eta)                              # D2H of what a dump reads, straight into the pinned buffers
        solver.field(odis.FIELD_VELOCITY, out=h_v)
        return solver.dissipation_avg()

    e2e_interval(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        e2e_interval(k + 1)
    torch.cuda.synchronize()
    el = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": round(Ke * S / el, 2), "unit": UNIT, "h2d_bytes_per_step": 8 * (4 * F + 4 * N),
           "d2h_bytes_per_step": 8 * (F + N + 1), "intervals_timed": Ke}

    variants = None
    if world == 1 and not args.no_variants:
        variants = {}
        for name, deg in (("no_self_gravity", 0), ("self_gravity_degree_8", 8)):
            if deg == L:
                continue
            alt = odis.Solver(mesh, workload_params(mesh), device=local_rank)
            if deg >= 2:
                alt.enable_self_gravity(deg, shell_factor(deg))
            alt.step(2 * S)
            n = max(S, min(K * S, 1000))
            variants[name] = {"timesteps_per_s": round(n / (alt.step_timed(n) * 1e-3), 1)}
            alt.close()
    if variants is not None and not args.no_probes:
        # opt-in kernel selections that are not the default, each timed in its own process (not part of `value`)
        torch.cuda.synchronize()
        variants["other_baseline_configs"] = run_probe("other_configs", args, 120.0)      # kernels that have run on B200s before
        variants["e2e_pipelined"] = run_probe("e2e_pipelined", args, 120.0)
        variants["opt_in_selections"] = [run_probe("nonlinear", args, 120.0), run_probe("headline_selections", args, 240.0)]
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (icosahedral-bisection grid generated in the reference's grid_lN.txt conventions; zero initial state, tidal forcing)",
                "config": {"workload": f"Enceladus subsurface ocean (LID_LOVE, 23 km shell), ECC tide, linear drag, {N} cells / {F} edges "
                                       f"(reference file level {args.level} = BASELINE 'L{args.level - 1}'); "
                                       + (f"self-gravity / shell-pressure term by spherical harmonics to degree {L} (least-squares analysis + synthesis every step; "
                                          f"factors 1 - beta_l of the reference's 23 km Enceladus table)" if L >= 2 else "no self-gravity term"),
                           "sh_degree": L, "kernel_select": args.kernel_select,
                           "cells": N, "edges": F, "lte_steps_per_bench_step": S, "dt_s": prm["dt"],
                           "cache": f"working set {dev_bytes / 1e6:.0f} MB device, {alg_bytes / 1e6:.0f} MB streamed per LTE step > 126 MB L2 (no flush needed)",
                           "parallelism": "1 GPU" if world == 1 else
                           f"{world} GPUs, grid cut into {world} space-filling-curve parts, one-ring halo, one exchange per step (boundary-edge "
                           f"velocities pushed by the edge kernel itself as NVLink peer stores; ghost cells updated locally) "
                           f"(rank 0: {part['own_cells']} own + {part['ghost_cells']} ghost cells, {part['n_peers']} neighbours)"},
                "cell_updates_per_s": round(value * N, 1), "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        if variants:
            line["variants"] = variants
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_port(mesh, prm, sh_degree=L)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        if world > 1 and args.level >= 10 and local_rank == 0:
            shutil.rmtree(shared, ignore_errors=True)
        dist.destroy_process_group()


def run_reference(args) -> None:
    """The reference's own CPU implementation of the path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import geodesicodis_b200 as odis
    from oracle.build_oracle import reference_binary
    level = args.level
    S = args.ref_substeps
    K, W = args.steps, max(args.warmup, 0)
    nsteps = (K + W) * S
    pos, fr, cen = odis.generate_grid(level)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"] - ENCELADUS["shell"])
    prm = workload_params(mesh)
    N, F = mesh.n_cells, mesh.n_edges
    cores = os.cpu_count() or 1
    binary = reference_binary(level, openmp=True)          # -fopenmp build: the reference's own omp loops + row-parallel sparse products
    omp = binary is not None
    if binary is None:
        binary, cores = reference_binary(level), 1
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Enceladus subsurface ocean (LID_LOVE), ECC tide, linear drag, {N} cells / {F} edges; the reference's "
                                   "self-gravity term is commented out at HEAD (src/spatialOperators.cpp:387-462), so its loop runs without it",
                       "cells": N, "edges": F, "lte_steps_per_bench_step": S}}
    if binary is not None:
        d = tempfile.mkdtemp(prefix="odis_ref_bench_")
        try:
            os.makedirs(d + "/input_files"); os.makedirs(d + "/DATA")
            odis.write_grid_file(f"{d}/input_files/grid_l{level}.txt", pos, fr, cen)
            # the reference quantises dt to period/(100k): ask for our dt, then bound the loop to nsteps
            period = 2 * round(math.pi / ENCELADUS["omega"])
            dt, total = odis.quantise_time_step(float(period), prm["dt"])
            keys = {"radius": ENCELADUS["radius"], "k2": 0.0, "h2": 0.0, "love reduction factor": ENCELADUS["love_reduct"],
                    "angular velocity": ENCELADUS["omega"], "surface gravity": ENCELADUS["g"], "semimajor axis": 238.02e6,
                    "eccentricity": ENCELADUS["ecc"], "obliquity": 0.0, "ocean thickness": ENCELADUS["h"], "shell thickness": ENCELADUS["shell"],
                    "friction coefficient": ENCELADUS["alpha"], "friction type": "LINEAR", "potential": "ECC", "surface type": "LID_LOVE",
                    "advection": "false", "solver type": "AB3", "sh degree": 2, "geodesic grid level": level, "output time": 1,
                    "dissipation output": "false", "dissipation avg output": "true", "kinetic avg output": "false",
                    "displacement output": "false", "velocity output": "false", "velocity cartesian output": "false",
                    "sh coefficient output": "false", "initial conditions": "NONE", "dummy1 output": "false",
                    "simulation end time": repr((nsteps - 0.5) / total), "time step": repr(prm["dt"]), "core number": 1, "rbf epsilon": 0.5}
            with open(d + "/input.in", "w") as f:
                f.write("".join(f"{k}; {v}; bench;\n" for k, v in keys.items()))
            t0 = time.perf_counter()
            subprocess.run([binary, "--quiet-restart"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                           env=dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close"))
            wall = time.perf_counter() - t0
            timing = dict(l.split() for l in open(d + "/DATA/ref_timing.txt") if l.strip())
            loop_s = float(timing["loop_seconds"]) - float(timing["dump_seconds_inside_loop"])
            value = nsteps / loop_s
            kind = "reference"
            build = (f"-fopenmp as in its Makefile:32-33, {cores} threads: its own omp loops + row-parallel sparse products" if omp
                     else "serial as in the active line of its Makefile")
            sample = (f"{nsteps} LTE steps in the reference's own ab3Explicit loop (unmodified sources, g++ -O3 -march=native, {build}) "
                      f"at {N} cells; loop {loop_s:.1f} s of {wall:.0f} s wall (the rest is the reference's mesh construction)")
        finally:
            shutil.rmtree(d, ignore_errors=True)
    else:
        base = cpu_baseline_port(mesh, prm, budget_s=20.0)
        value, kind, cores, sample = base["value"], "port", 1, base["sample"] + " (oracle/_ref binary for this level not present)"
    line.update({"value": round(value, 3), "ms_per_step": round(1e3 * S / value, 2),
                 "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=9, help="reference grid-file level; 9 = 655,362 cells (BASELINE 'L8')")
    ap.add_argument("--substeps", type=int, default=100, help="LTE time steps per bench step (one output interval)")
    ap.add_argument("--ref-substeps", type=int, default=2, help="LTE time steps per bench step for --impl reference")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra device-resident timings with other --sh-degree values")
    ap.add_argument("--no-probes", action="store_true", help="skip the subprocess timings of the opt-in kernel selections")
    ap.add_argument("--kernel-select", type=int, default=0, help="opt-in kernel selection bits (include/odis_b200.h, odis_params.reserved[0]); "
                    "0 = the default kernels. E.g. 16: self-gravity step in 3 launches (also on partitioned grids), 128: 16-bit stencil ids")
    ap.add_argument("--variant-probe", default="", help=argparse.SUPPRESS)
    ap.add_argument("--probe-level", type=int, default=8, help=argparse.SUPPRESS)
    ap.add_argument("--sh-degree", type=int, default=2, help="self-gravity term by spherical harmonics to this degree (the shipped input.in's "
                    "'sh degree' is 2); 0 = off, as at reference HEAD where the term is commented out")
    args = ap.parse_args()
    if args.variant_probe:
        variant_probe(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
