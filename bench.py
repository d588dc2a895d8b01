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

# torchrun gives every rank ONE OpenMP thread unless the caller says otherwise; the host-side set-up (grid generator, mesh tables,
# renumbering, the ranks' table builds) would then crawl on one core each. Every rank takes its share of the cores instead
# (before libgomp is loaded; nothing here is timed: the device times come from CUDA events).
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

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


def ncu_traffic_per_launch(kernel: str) -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at 655,362 cells on one GPU, from the committed
    `ncu --set full` summary (profiles/edge_step_summary.json: one entry per kernel name), if one exists."""
    p = os.path.join(ROOT, "profiles", "edge_step_summary.json")
    try:
        d = json.load(open(p))
        return float(d["kernels"][kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def sg_alg_bytes(N: int, L: int) -> int:
    """Matrix-free self-gravity term: 4 basis inputs per cell for the analysis (folded into the cell update for L <= 4) + 4 basis
    inputs and {eta,U} read + written by the synthesis."""
    return 0 if L < 2 else (32 + 64) * N


def workload_config(N: int, F: int, level: int, L: int, S: int, dt: float, world: int, kernel_select: int = 0) -> dict:
    """The `config` of a bench line; both arms build it from the same inputs so that the driver sees one configuration."""
    alg = 200 * F + 128 * N + sg_alg_bytes(N, L)
    return {"workload": f"Enceladus subsurface ocean (LID_LOVE, 23 km shell), ECC tide, linear drag, {N} cells / {F} edges "
                        f"(reference file level {level} = BASELINE 'L{level - 1}'); "
                        + (f"self-gravity / shell-pressure term by spherical harmonics to degree {L} (least-squares analysis + synthesis every step; "
                           f"factors 1 - beta_l of the reference's 23 km Enceladus table)" if L >= 2 else "no self-gravity term"),
            "sh_degree": L, "kernel_select": kernel_select, "cells": N, "edges": F, "lte_steps_per_bench_step": S, "dt_s": dt,
            "cache": f"inputs larger than L2: {alg / 1e6:.0f} MB algorithmic bytes streamed per LTE step ({alg / world / 1e6:.0f} MB per GPU) vs 126 MB L2; "
                     "no flush between steps" + ("" if alg / world > 126e6 else " (per-GPU share below the L2 size: the tables may stay resident, stated)"),
            "parallelism": "1 GPU" if world == 1 else
            f"{world} GPUs, the one grid cut into {world} space-filling-curve parts, one-ring halo, one exchange per step (boundary-edge "
            f"velocities pushed by the edge kernel itself as NVLink peer stores; ghost cells updated locally; harmonic sums all-reduced through peer memory)"}


def oracle_run(mesh, prm: dict, sh_degree: int, nsteps: int, budget_s: float | None = None):
    """The oracle's plain-C restatement of the reference loop (bit-identical to the reference solver, tests/test_oracle_pinned.py) from
    the zero state: returns (oracle, steps taken, seconds). budget_s: size the run to about that much CPU time instead of nsteps."""
    from oracle.lte_oracle import LteOracle
    keys = ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl", "shell_thickness", "potential", "friction", "surface", "init_load")
    o = LteOracle(mesh.tables, {k: prm[k] for k in keys})
    if sh_degree >= 2:
        from oracle import sh_oracle
        Y = sh_oracle.basis(mesh.tables["node_pos_sph"], sh_degree)
        o.set_self_gravity(Y, sh_oracle.apply_operator(Y, shell_factor(sh_degree)))
    o.set_state()
    done, el = 0, 0.0
    if budget_s is not None:
        t0 = time.perf_counter(); o.step(3); probe = (time.perf_counter() - t0) / 3
        done, el = 3, probe * 3
        nsteps = int(max(5, min(2000, budget_s / max(probe, 1e-9)))) - 3
    t0 = time.perf_counter(); o.step(nsteps); el += time.perf_counter() - t0
    return o, done + nsteps, el


def cpu_baseline_and_parity(odis, mesh, prm: dict, sh_degree: int, device: int, budget_s: float = 15.0) -> tuple[dict, dict]:
    """cpu_baseline: the oracle timed on one host core on a bounded sample of the headline workload. parity: the SAME oracle runs are
    kept and compared with the CUDA path after the same number of steps from the same (zero) state — with the self-gravity term
    (tolerance 1e-10: the term has no reference arithmetic to follow) and without it (bit equality required)."""
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    o, n, el = oracle_run(mesh, prm, sh_degree, 0, budget_s)
    base = {"value": n / el, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} LTE steps of the same {mesh.n_cells}-cell workload from the zero state"
                      + (f" (self-gravity term to degree {sh_degree} included)" if sh_degree >= 2 else "") + ", oracle/lte_oracle.c (gcc -O2), single thread"}
    parity = {"oracle": "oracle/lte_oracle.c, pinned bit for bit to the reference's own solver (tests/test_oracle_pinned.py)", "tolerance_rel": 1e-10}

    def compare(oracle, nsteps, degree):
        sv = odis.Solver(mesh, prm, device=device)
        if degree >= 2:
            sv.enable_self_gravity(degree, shell_factor(degree))
        sv.step(nsteps)
        v, eta = sv.field(odis.FIELD_VELOCITY), sv.field(odis.FIELD_ETA)
        ov, oe = oracle.field(0), oracle.field(1)
        out = {"steps": nsteps, "max_rel_eta": rel(eta, oe), "max_rel_v": rel(v, ov),
               "bit_identical": bool(np.array_equal(v, ov) and np.array_equal(eta, oe)),
               "within_tolerance": bool(rel(eta, oe) <= 1e-10 and rel(v, ov) <= 1e-10)}
        sv.close()
        return out

    if sh_degree >= 2:
        parity["with_self_gravity_term"] = compare(o, n, sh_degree)
        o2, n2, _ = oracle_run(mesh, prm, 0, 40)
        parity["without_self_gravity_term"] = compare(o2, n2, 0)
    else:
        parity["without_self_gravity_term"] = compare(o, n, 0)
    parity["ok"] = bool(parity["without_self_gravity_term"]["bit_identical"] and
                        parity.get("with_self_gravity_term", {"within_tolerance": True})["within_tolerance"])
    return base, parity


def e2e_pipelined(odis, sv, mesh, S: int, Ke: int, pin) -> dict:
    """One output interval end to end through the C ABI with page-locked HOST buffers, copies overlapped with stepping:
    odis_stage_state / odis_commit_state bring interval k+1's state in (H2D on the second stream) while interval k steps,
    odis_snapshot_begin / _wait take interval k's eta / v / dissipation out while interval k+1 steps. Same bytes per interval as the
    synchronous calls, whose result for the last interval is the check."""
    import torch
    N, F = mesh.n_cells, mesh.n_edges
    h_v, h_eta = pin(sv.field(odis.FIELD_VELOCITY)), pin(sv.field(odis.FIELD_ETA))
    h_dv, h_de = pin(sv.field(odis.FIELD_DVDT).ravel()), pin(sv.field(odis.FIELD_DETADT).ravel())
    it0, fields = sv.iter, sv.SNAP_ETA | sv.SNAP_VELOCITY

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
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, last = run(Ke)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    # the same interval with the synchronous calls: the comparison value and the check of the pipelined result
    def sync_interval(k):
        sv.set_state(h_v, h_eta, h_dv, h_de, iter=it0 + k * S)
        sv.step(S)
        e = sv.field(odis.FIELD_ETA)
        v = sv.field(odis.FIELD_VELOCITY)
        sv.dissipation_avg()
        return e, v
    sync_interval(0)
    t0 = time.perf_counter()
    for k in range(Ke):
        e, v = sync_interval(Ke - 1)
    torch.cuda.synchronize()
    el_sync = time.perf_counter() - t0
    same = bool(np.array_equal(e, last["eta"]) and np.array_equal(v, last["velocity"]))
    return {"value": round(Ke * S / el, 2), "unit": UNIT, "h2d_bytes_per_step": 8 * (4 * F + 4 * N), "d2h_bytes_per_step": 8 * (F + N + 1),
            "intervals_timed": Ke,
            "path": "odis_stage_state / odis_commit_state (H2D of the next interval's state on a second stream) + odis_step + odis_snapshot_begin / "
                    "_wait (D2H of eta, v, dissipation while the next interval steps); pipeline fill (the first, unoverlapped upload) inside the timed region",
            "synchronous_calls_value": round(Ke * S / el_sync, 2), "fields_identical_to_synchronous_calls": same}


def variant_probe(args) -> None:
    """Child process of the N = 1 bench (own CUDA context and host memory, so nothing here can disturb the headline measurement):
    prints one JSON line."""
    import geodesicodis_b200 as odis
    name = args.variant_probe
    out = {"probe": name}
    if name == "headline_selections":
        pos, fr, cen = odis.generate_grid(args.level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"] - ENCELADUS["shell"])
        prm, L, S = workload_params(mesh), max(args.sh_degree, 2), args.substeps
        ref_eta = None
        for key, sel in (("default", 0), ("stencil_ids_32bit_only", 128), ("direct_load_baseline_kernels_5_launches", 1)):
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
    elif name == "nonlinear":
        level = args.probe_level                                 # 8: 163,842 cells (BASELINE 'L7')
        # the headline physics (Enceladus ocean, ECC tide, linear drag; free surface) with `advection; true`: the nonlinear terms are
        # small against the linear ones here but cost the same, and the state stays finite (with the shipped input.in's Earth-sized
        # obliquity forcing the reference's nonlinear scheme itself goes non-finite within ~150 steps at this resolution)
        pos, fr, cen = odis.generate_grid(level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"])
        nl = odis.nonlinear_tables(mesh, 0.5)
        prm = dict(workload_params(mesh), surface=0, shell_thickness=0.0, love_reduct=1.0)
        res = {}
        # default: 4 launches per step (energy diagnostic and next potential folded into the edge / cell update); baseline selection: the
        # six gather kernels + diagnostics + potential pass = 8 launches. Every timed chunk restarts from the zero state: at this resolution
        # the reference's nonlinear scheme goes unstable after a few hundred steps, and a timing over NaN fields would prove little.
        for key, sel in (("4_launch_default", 0), ("8_launch_baseline", 1)):
            sv = odis.Solver(mesh, dict(prm, kernel_select=sel))
            sv.enable_advection(nl)
            sv.step(60)
            res[key] = sv.field(odis.FIELD_ETA)
            best = None
            for _ in range(3):
                sv.set_state()
                sv.step(20)
                ms = sv.step_timed(100) / 100
                best = ms if best is None else min(best, ms)
            out[key + "_timesteps_per_s"] = round(1.0 / (best * 1e-3), 1)
            out[key + "_us_per_step"] = round(best * 1e3, 2)
            out[key + "_finite"] = bool(math.isfinite(sv.dissipation_avg()))
            sv.close()
        out["cells"] = mesh.n_cells
        out["fields_identical"] = bool(np.array_equal(res["4_launch_default"], res["8_launch_baseline"]))
    elif name == "synthetic_l10_1gpu":
        # BASELINE config 4's grid on ONE GPU (the N = 1 point of the 10,485,762-cell scaling curve): free-surface LTE, ECC, linear drag
        t0 = time.time()
        pos, fr, cen = odis.generate_grid(11)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"])
        del pos, fr, cen
        prm = dict(workload_params(mesh), surface=0, shell_thickness=0.0, love_reduct=1.0)
        sv = odis.Solver(mesh, prm)
        sv.step(24)
        ms = sv.step_timed(200)
        _, alg = sv.footprint()
        peak, _ = measured_peak()
        out.update({"cells": mesh.n_cells, "timesteps_per_s": round(200 / (ms * 1e-3), 2), "steps_timed": 200, "n_gpus": 1,
                    "cell_updates_per_s": round(mesh.n_cells * 200 / (ms * 1e-3), 1), "algorithmic_GBps": round(alg * 200 / (ms * 1e-3) / 1e9, 1),
                    "frac_of_measured_hbm_peak": round(alg * 200 / (ms * 1e-3) / 1e9 / peak, 4), "finite": bool(math.isfinite(sv.dissipation_avg())),
                    "setup_s": round(time.time() - t0, 1)})
        sv.close()
    elif name == "other_configs":
        # the remaining single-GPU shapes of BASELINE.json's configs, each a short device-resident timing: [1] Enceladus free-surface
        # ocean, 163,842 cells, linear drag, no self-gravity; [4] one GPU's share of the ensemble sweep: 32 members (ocean thickness x drag)
        # on 40,962 cells, self-gravity to degree 8 as batched FP64 tensor-core GEMMs
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
        peak, _ = measured_peak()
        out["ensemble_32_members_%d_cells_sh8" % mesh.n_cells] = {"batched_steps_per_s": round(1e3 / ms, 1), "member_steps_per_s": round(M * 1e3 / ms, 1),
                                                      "algorithmic_GBps": round(info["algorithmic_bytes_per_step"] / (ms * 1e-3) / 1e9, 1),
                                                      "frac_of_measured_hbm_peak": round(info["algorithmic_bytes_per_step"] / (ms * 1e-3) / 1e9 / peak, 4)}
        ens.close()
    else:
        out["error"] = "unknown probe"
    print(json.dumps(out), flush=True)


PROBE_BUDGET_S = 520.0          # all child-process probes of one bench run together (each also has its own limit)
_probe_deadline = [None]


def run_probe(name: str, args, limit: float = 180.0) -> dict:
    """Runs `bench.py --variant-probe name` in a subprocess; any failure is recorded instead of raised. A probe that does not return
    is killed at its limit and what it had printed until then is kept."""
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
        # the device is drained BEFORE the collective: a partitioned step in flight waits in-kernel for the neighbours' flags, and the
        # merged cell kernel is a cooperative launch (all CTAs resident together) — an NCCL kernel squeezed in beside it would wait for a
        # peer whose own NCCL kernel cannot start behind ITS in-flight steps: a cycle. (include/odis_b200.h, odis_step: synchronize
        # partitioned solvers before any other operation that waits on another GPU.)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op="max") -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    def connect(solver):                            # every rank publishes its halo buffers; neighbours map them
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, solver.halo_blob())
            solver.halo_connect(blobs)
            dist.barrier()

    def whole(solver, field):                       # a partitioned solver returns its own entries, zeros elsewhere
        a = solver.field(field)
        if dist is not None:
            t = torch.from_numpy(a).cuda()
            dist.all_reduce(t)
            a = t.cpu().numpy()
        return a

    def timed_steps(solver, n):                     # CUDA events on the solver's own stream, max over ranks
        barrier()
        ms = solver.step_timed(n)
        torch.cuda.synchronize()
        barrier()
        return reduce_ranks(ms)

    radius = ENCELADUS["radius"] - ENCELADUS["shell"]                 # LID_LOVE: boundaryConditions.cpp:126
    pos, fr, cen = odis.generate_grid(args.level)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, radius)
    prm = workload_params(mesh)
    if args.kernel_select:                          # experiments only (odis_params.reserved[0]); the default line is selection 0
        prm = dict(prm, kernel_select=args.kernel_select)
    L = args.sh_degree
    N, F = mesh.n_cells, mesh.n_edges
    S, K, W = args.substeps, args.steps, max(args.warmup, 3)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

    # ---- N > 1: the partitioned run against the single-GPU solver, same steps from the same state (rank 0 holds the single one) -------
    parity_multi = None
    value_no_sg = None
    if world > 1:
        P = 60
        parity_multi = {"steps": P, "tolerance_rel": 1e-10}
        for key, deg in (("without_self_gravity_term", 0), ("with_self_gravity_term", L)):
            if key == "with_self_gravity_term" and L < 2:
                continue
            part_solver = odis.Solver(mesh, prm, device=local_rank, rank=rank, world=world)
            connect(part_solver)
            if deg >= 2:
                part_solver.enable_self_gravity(deg, shell_factor(deg))
            part_solver.step(P)
            v, eta = whole(part_solver, odis.FIELD_VELOCITY), whole(part_solver, odis.FIELD_ETA)
            if rank == 0:
                single = odis.Solver(mesh, prm, device=local_rank)
                if deg >= 2:
                    single.enable_self_gravity(deg, shell_factor(deg))
                single.step(P)
                sv_, se_ = single.field(odis.FIELD_VELOCITY), single.field(odis.FIELD_ETA)
                single.close()
                parity_multi[key] = {"max_rel_eta": rel(eta, se_), "max_rel_v": rel(v, sv_),
                                     "bit_identical": bool(np.array_equal(v, sv_) and np.array_equal(eta, se_)),
                                     "within_tolerance": bool(rel(eta, se_) <= 1e-10 and rel(v, sv_) <= 1e-10)}
            if deg == 0:                            # the same run without the term, timed: comparable with the reference arm (which has no term)
                for _ in range(W):
                    part_solver.step(S)
                value_no_sg = round(K * S / (timed_steps(part_solver, K * S) * 1e-3), 2)
            part_solver.synchronize()
            barrier()
            part_solver.close()
        if rank == 0:
            parity_multi["ok"] = bool(parity_multi["without_self_gravity_term"]["bit_identical"] and
                                      parity_multi.get("with_self_gravity_term", {"within_tolerance": True})["within_tolerance"])

    solver = odis.Solver(mesh, prm, device=local_rank, rank=rank, world=world)
    connect(solver)
    if L >= 2:                                      # self-gravity / shell-pressure term (BASELINE config 3), matrix-free kernels
        solver.enable_self_gravity(L, shell_factor(L))
    dev_bytes, alg_bytes = solver.footprint()

    # ---- device-resident throughput -------------------------------------------------------------
    sampler = ClockSampler(local_rank)              # from the warm-up on: at N = 8 the timed region is ~0.1 s, one nvidia-smi period
    t_warm0 = time.time()
    if world > 1:
        W = max(W, 30)                              # (reported as run)
    for _ in range(W):
        solver.step(S)
    barrier()
    launches0 = solver.launches
    t_wall0 = time.time()
    ms = solver.step_timed(K * S)               # CUDA events on the solver's own stream
    torch.cuda.synchronize()
    t_wall1 = time.time()
    clocks = sampler.stop(t_warm0 if t_wall1 - t_wall0 < 0.5 else t_wall0, t_wall1)
    clocks["window"] = "warm-up + timed region (same load)" if t_wall1 - t_wall0 < 0.5 else "timed region"
    launches = solver.launches - launches0
    if not math.isfinite(solver.dissipation_avg()):
        raise SystemExit("bench.py: the run blew up (non-finite dissipation); the timing would be meaningless")
    barrier()
    ms = reduce_ranks(ms)
    value = K * S / (ms * 1e-3)                  # all ranks advance the same K*S steps of the one global grid

    # ---- per-kernel timing for the roofline (live, CUDA events around every launch) --------------
    nprof = min(K * S, 400)
    edge_ms, cell_ms, sh_ms = solver.step_profiled_sh(nprof)
    edge_us, cell_us, sh_us = edge_ms / nprof * 1e3, cell_ms / nprof * 1e3, sh_ms / nprof * 1e3
    peak, peak_src = measured_peak()
    part = solver.partition()
    narrow = (args.kernel_select & (1 | 128)) == 0
    edge_bytes = (180 if narrow else 200)       # 16-bit stencil ids stream 20 B per edge instead of 40
    edge_alg = edge_bytes * part["own_edges"]   # per-edge algorithmic bytes x edges per launch (this rank's)
    edge_kernel = "edge_step_kernel" if (args.kernel_select & 1) else ("edge_step_pipe16_kernel" if narrow else "edge_step_pipe_kernel")
    achieved = edge_alg / (edge_us * 1e-6) / 1e9
    fused_sg = L >= 2 and L <= 4 and not (args.kernel_select & 1)
    launches_per_step = launches / (K * S)
    merged_sg = fused_sg and launches_per_step < 2.5        # solve + synthesis behind a grid barrier inside the cell update's launch
    # per cell: 128 B of the update; + 32 B basis inputs of the folded analysis; + 64 B of the merged synthesis (basis inputs, {eta,U} r/w)
    cell_alg = (128 + (32 if fused_sg else 0) + (64 if merged_sg else 0)) * part["own_cells"]
    roofline = {"bound": "hbm", "kernel": edge_kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic_per_launch(edge_kernel) if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": edge_alg, "avg_launch_us": round(edge_us, 2),
                "algorithmic_bytes_per_edge": edge_bytes,
                "cell_step_kernel": {"kernel": "cell_step_kernel" if (args.kernel_select & 1) else "cell_step_pipe_kernel",
                                     "algorithmic_bytes_per_launch": cell_alg, "avg_launch_us": round(cell_us, 2),
                                     "achieved": round(cell_alg / (cell_us * 1e-6) / 1e9, 1), "frac": round(cell_alg / (cell_us * 1e-6) / 1e9 / peak, 4),
                                     "note": "with the self-gravity term to degree <= 4 the harmonic analysis is accumulated inside this launch (+32 B per cell) "
                                             "and, merged, so are solve + synthesis behind a grid-wide barrier (+64 B per cell, mostly L2 hits)"},
                "self_gravity_kernels": {"avg_us_per_step": round(sh_us, 2) if not merged_sg else 0.0,
                                         "launches_per_step": ((0 if merged_sg else 1) if fused_sg else 3) if L >= 2 else 0,
                                         "algorithmic_bytes_per_step": (0 if merged_sg else 64 * part["own_cells"]) if L >= 2 else 0,
                                         "note": "matrix-free: the harmonic basis is rebuilt per cell by recurrence instead of streaming 8*(l_max+1)^2 B per cell"},
                "whole_step": {"algorithmic_bytes": alg_bytes, "achieved": round(alg_bytes * value / 1e9, 1),
                               "frac": round(alg_bytes * value / 1e9 / peak, 4), "frac_of_8TBs_nominal": round(alg_bytes * value / 8e12, 4),
                               "survey_b_alg_bytes": 200 * part["own_edges"] + 128 * part["own_cells"],
                               "frac_on_survey_b_alg": round((200 * part["own_edges"] + 128 * part["own_cells"]) * value / 1e9 / peak, 4),
                               "note": "per GPU: this rank's share of the grid; for N>1 the halo exchange is part of the two kernels"}}
    if world > 1:
        roofline["traffic_note"] = "no ncu capture of a partitioned rank (ncu is single-GPU here); the N = 1 line carries the measured DRAM bytes"

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    def pin(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return torch.from_numpy(a).pin_memory().numpy()
    Ke = max(1, min(K, 20))
    if world == 1:
        e2e = e2e_pipelined(odis, solver, mesh, S, Ke, pin)
    else:
        h_v, h_eta = pin(whole(solver, odis.FIELD_VELOCITY)), pin(whole(solver, odis.FIELD_ETA))
        h_dv, h_de = pin(whole(solver, odis.FIELD_DVDT).ravel()), pin(whole(solver, odis.FIELD_DETADT).ravel())
        o_v, o_eta = pin(np.zeros(F)), pin(np.zeros(N))
        it0 = solver.iter

        def e2e_interval(k: int):
            solver.set_state(h_v, h_eta, h_dv, h_de, iter=it0 + k * S)          # H2D of this rank's part of the interval's inputs
            solver.step(S)
            solver.field(odis.FIELD_ETA, out=o_eta)                              # D2H of this rank's part of what a dump reads
            solver.field(odis.FIELD_VELOCITY, out=o_v)
            return solver.dissipation_avg()

        e2e_interval(0)
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            e2e_interval(k + 1)
        torch.cuda.synchronize()
        el_sync = reduce_ranks(time.perf_counter() - t0)
        barrier()
        # the same intervals pipelined (what the N = 1 line reports): every rank packs + uploads its share of interval k+1's state on its
        # second stream while interval k steps, and its own entries of interval k come back through the snapshot slots meanwhile
        h_v2, h_eta2, h_dv2, h_de2 = h_v, h_eta, h_dv, h_de
        snap_fields = solver.SNAP_ETA | solver.SNAP_VELOCITY

        def pipelined(n):
            solver.stage_state(h_v2, h_eta2, h_dv2, h_de2)
            for k in range(n):
                solver.commit_state(iter=it0 + k * S)
                solver.step(S)                                  # enqueued first: the device steps while the host packs the next state
                solver.snapshot_begin(k & 1, snap_fields)
                if k + 1 < n:
                    solver.stage_state(h_v2, h_eta2, h_dv2, h_de2)
                if k > 0:
                    solver.snapshot_wait((k - 1) & 1, copy=False)
            last = solver.snapshot_wait((n - 1) & 1, copy=False)
            solver.synchronize()
            return last

        pipelined(2)
        barrier()
        t0 = time.perf_counter()
        last = pipelined(Ke)
        torch.cuda.synchronize()
        el = reduce_ranks(time.perf_counter() - t0)
        barrier()
        # check: the pipelined interval's own entries against the synchronous calls from the same state
        solver.set_state(h_v2, h_eta2, h_dv2, h_de2, iter=it0 + (Ke - 1) * S)
        solver.step(S)
        own_c, own_e = solver.partition_map()
        same = bool(np.array_equal(solver.field(odis.FIELD_ETA)[own_c], last["eta"]) and
                    np.array_equal(solver.field(odis.FIELD_VELOCITY)[own_e], last["velocity"]))
        same = reduce_ranks(1.0 if same else 0.0, "min") > 0.5
        loc_e, loc_c = part["own_edges"] + part["ghost_edges"], part["own_cells"] + part["ghost_cells"]
        h2d = reduce_ranks(8.0 * (loc_e + 3 * part["own_edges"] + 4 * loc_c), "sum")
        d2h = reduce_ranks(8.0 * (part["own_edges"] + part["own_cells"] + 1), "sum")
        e2e = {"value": round(Ke * S / el, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "intervals_timed": Ke,
               "path": "per rank: odis_stage_state (the rank packs its share of the next interval's GLOBAL host arrays and uploads only that, on "
                       "its second stream, while the current interval steps) / odis_commit_state + odis_step + odis_snapshot_begin / _wait (the "
                       "rank's own eta, v, dissipation come back compact while the next interval steps); bytes are summed over the ranks; "
                       "pipeline fill inside the timed region",
               "synchronous_calls_value": round(Ke * S / el_sync, 2), "fields_identical_to_synchronous_calls": same}

    # ---- BASELINE config 4 inside the same launch: synthetic 10,485,762-cell grid partitioned over the N GPUs -------------------------
    variants = None
    if world > 1 and not args.no_variants:
        variants = {"synthetic_l10": synthetic_l10_partitioned(odis, dist, torch, rank, local_rank, world, reduce_ranks, barrier)}
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
        value_no_sg = variants.get("no_self_gravity", {}).get("timesteps_per_s", round(value, 2) if L < 2 else None)
    cpu_base = parity = None
    if world == 1 and not args.no_cpu:
        cpu_base, parity = cpu_baseline_and_parity(odis, mesh, prm, L, local_rank)
    if variants is not None and world == 1 and not args.no_probes:
        # other configurations and selections, each timed in its own process (not part of `value`)
        torch.cuda.synchronize()
        variants["synthetic_l10_1gpu"] = run_probe("synthetic_l10_1gpu", args, 330.0)
        variants["other_baseline_configs"] = run_probe("other_configs", args, 120.0)
        variants["kernel_selections"] = [run_probe("nonlinear", args, 120.0), run_probe("headline_selections", args, 150.0)]
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (icosahedral-bisection grid generated in the reference's grid_lN.txt conventions; zero initial state, tidal forcing)",
                "config": workload_config(N, F, args.level, L, S, prm["dt"], world, args.kernel_select),
                "partition": {"rank0_own_cells": part["own_cells"], "rank0_ghost_cells": part["ghost_cells"], "rank0_neighbours": part["n_peers"]},
                "value_no_self_gravity": value_no_sg,
                "cell_updates_per_s": round(value * N, 1), "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        if parity is not None:
            line["parity"] = parity
        if parity_multi is not None:
            line["parity_vs_single_gpu"] = parity_multi
        if variants:
            line["variants"] = variants
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def synthetic_l10_partitioned(odis, dist, torch, rank, local_rank, world, reduce_ranks, barrier) -> dict:
    """BASELINE config 4: synthetic icosahedral-bisection grid with 10,485,762 cells (reference file level 11), free-surface LTE (no
    self-gravity term), partitioned over the N GPUs of this launch, 200 timed steps. Every stage ends in an all-reduce of an ok flag so
    that a failure on one rank ends the variant on all of them instead of hanging a collective."""
    out = {"cells": 10485762, "n_gpus": world, "steps_timed": 200, "config": "free-surface Enceladus ocean, ECC tide, linear drag, no self-gravity term"}
    shared = f"/dev/shm/odis_b200_mesh_l11_{os.environ.get('MASTER_PORT', '0')}"
    t0 = time.time()

    def stage(fn):
        ok, res, err = 1.0, None, ""
        try:
            res = fn()
        except Exception as e:                                   # noqa: BLE001
            ok, err = 0.0, repr(e)[:300]
        if reduce_ranks(ok, "min") < 0.5:
            raise RuntimeError(err or "another rank failed")
        return res

    solver = None
    try:
        def make_mesh():
            if local_rank == 0:                                  # one rank builds, all cores (the others wait at the stage's all-reduce)
                cores = os.cpu_count() or 8
                gomp = None
                try:
                    import ctypes
                    gomp = ctypes.CDLL("libgomp.so.1")
                    gomp.omp_set_num_threads(cores)
                except OSError:
                    pass
                pos, fr, cen = odis.generate_grid(11)
                odis.Mesh.from_arrays(pos, fr, cen, ENCELADUS["radius"], threads=cores).save(shared)
                if gomp is not None:
                    gomp.omp_set_num_threads(max(1, cores // world))
        stage(make_mesh)
        mesh = stage(lambda: odis.Mesh.load(shared))
        prm = dict(workload_params(mesh), surface=0, shell_thickness=0.0, love_reduct=1.0)
        solver = stage(lambda: odis.Solver(mesh, prm, device=local_rank, rank=rank, world=world))

        def connect():
            blobs = [None] * world
            dist.all_gather_object(blobs, solver.halo_blob())
            solver.halo_connect(blobs)
        stage(connect)
        out["setup_s"] = round(time.time() - t0, 1)
        stage(lambda: (solver.step(24), solver.synchronize()))      # drained before the next collective (see barrier())
        barrier()
        ms = stage(lambda: solver.step_timed(200))
        torch.cuda.synchronize()
        barrier()
        ms = reduce_ranks(ms)
        _, alg = solver.footprint()
        peak, _ = measured_peak()
        part = solver.partition()
        finite = reduce_ranks(1.0 if math.isfinite(solver.dissipation_avg()) else 0.0, "min") > 0.5
        out.update({"timesteps_per_s": round(200 / (ms * 1e-3), 2), "cell_updates_per_s": round(10485762 * 200 / (ms * 1e-3), 1),
                    "per_gpu_algorithmic_GBps": round(alg * 200 / (ms * 1e-3) / 1e9, 1),
                    "per_gpu_frac_of_measured_hbm_peak": round(alg * 200 / (ms * 1e-3) / 1e9 / peak, 4), "finite": bool(finite),
                    "rank0_own_cells": part["own_cells"], "rank0_ghost_cells": part["ghost_cells"]})
    except Exception as e:                                       # noqa: BLE001
        out["error"] = repr(e)[:300]
    finally:
        if solver is not None:
            try:
                solver.synchronize()
            except Exception:                                    # noqa: BLE001
                pass
        barrier()
        if solver is not None:
            solver.close()
        if local_rank == 0:
            shutil.rmtree(shared, ignore_errors=True)
    return out


def run_reference(args) -> None:
    """The reference's own CPU implementation of the path on this box's host cores (rank 0 only). The input files are written through the
    host-only library (no CUDA mapped into this process or the reference's)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    os.environ["ODIS_B200_HOST_ONLY"] = "1"
    import geodesicodis_b200 as odis
    from oracle.build_oracle import reference_binary
    level = args.level
    S, S_ref = args.substeps, args.ref_substeps
    K, W = args.steps, max(args.warmup, 0)
    nsteps = (K + W) * S_ref
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (icosahedral-bisection grid generated in the reference's grid_lN.txt conventions; zero initial state, tidal forcing)",
            "config": workload_config(N, F, level, args.sh_degree, S, prm["dt"], world, 0),
            "reference_notes": f"each bench step (one output interval of {S} LTE steps) is timed on a bounded sample of {S_ref} of its steps and scaled; "
                               "the reference's self-gravity term is commented out at its HEAD (src/spatialOperators.cpp:387-462), so its own loop runs "
                               "the configuration without that term (compare with the GPU arm's value_no_self_gravity)",
            "sampled_lte_steps_per_bench_step": S_ref}
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
        o, n, el = oracle_run(mesh, prm, 0, 0, 20.0)
        value, kind, cores = n / el, "port", 1
        sample = f"{n} LTE steps of the same {N}-cell workload, oracle/lte_oracle.c (gcc -O2), single thread (oracle/_ref binary for this level not present)"
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
    ap.add_argument("--kernel-select", type=int, default=0, help="kernel selection bits (include/odis_b200.h, odis_params.reserved[0]); "
                    "0 = the default kernels; 1: direct-load baseline kernels, 8: no CUDA-graph replay, 128: 32-bit stencil ids only")
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
