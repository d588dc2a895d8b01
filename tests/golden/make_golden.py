"""Generates the committed golden fixtures in tests/golden/ by running the reference's own solver
(oracle/_ref/odis_ref_l<L> = the unmodified reference sources, built by oracle/ref_build/Makefile) in
THIS container, where /root/reference exists. The fixtures travel to the GPU box; this script does not.

    python tests/golden/make_golden.py

Each case_<name>.npz holds: the input.in text, the grid (degrees, exactly as the reference parsed them),
the reference's derived scalars, sha256 digests of every mesh table (full tables for the small grids),
the complete FP64 solver state after the last step and the FP64 arrays handed to DumpData at every dump.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.build_oracle import build_reference  # noqa: E402
from oracle.refio import read_records, read_h5shim  # noqa: E402

REF_GRIDS = "/root/reference/input_files"
HERE = os.path.dirname(os.path.abspath(__file__))

BASE = {  # every key the reference reads; values overridden per case
    "radius": "252.1e3", "k2": "0.0", "h2": "0.0", "love reduction factor": "1.0", "angular velocity": "5.307e-5",
    "surface gravity": "0.113", "semimajor axis": "238.02e6", "eccentricity": "0.0047", "obliquity": "0.1",
    "orbital period": "118386.8", "ocean thickness": "38e3", "shell thickness": "0", "friction coefficient": "1e-7",
    "friction type": "LINEAR", "potential": "ECC", "surface type": "FREE", "advection": "false", "solver type": "AB3",
    "sh degree": "2", "geodesic grid level": "3", "output time": "10", "dissipation output": "false",
    "dissipation avg output": "true", "kinetic avg output": "false", "displacement output": "true",
    "velocity output": "true", "velocity cartesian output": "false", "sh coefficient output": "false",
    "initial conditions": "NONE", "dummy1 output": "false", "simulation end time": "1", "latitude spacing": "6.0",
    "time step": "100", "core number": "1", "rbf epsilon": "0.5",
}


def input_text(over: dict) -> str:
    cfg = dict(BASE)
    cfg.update(over)
    return "".join(f"{k}; \t {v}; \t note;\n" for k, v in cfg.items())


def parse_grid_degrees(path: str):
    """Degrees exactly as atof reads them (src/mesh.cpp:4043-4074)."""
    import re
    lat, lon, fr, cen = [], [], [], []
    with open(path) as f:
        next(f)
        for line in f:
            tok = re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line)
            if len(tok) < 21:
                continue
            lat.append(float(tok[1])); lon.append(float(tok[2]))
            fr.append([int(t) for t in tok[3:9]])
            cen.append([float(t) for t in tok[9:21]])
    return np.array(lat), np.array(lon), np.array(fr, dtype=np.int32), np.array(cen).reshape(-1, 6, 2)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


TABLES = ["node_pos_sph", "node_friends", "centroid_pos_sph", "control_volume_surf_area_map", "faces", "node_face_dir",
          "vertexes", "face_nodes", "face_vertexes", "face_interp_friends", "face_interp_weights", "face_len",
          "face_node_dist", "face_centre_m", "face_centre_pos_sph", "face_intercept_pos_sph", "face_area",
          "face_normal_vec_map", "vertex_pos_sph", "vertex_nodes", "vertex_R"]
SCALARS = ["radius", "angVel", "period", "g", "h", "alpha", "loveReduct", "shell_thickness", "e", "theta", "timeStep",
           "endTime", "totalIter", "outputTime", "tide_type", "fric_type", "surface_type", "advection"]


NL_TABLES = ["vertex_sinlat", "vertex_area", "vertex_R", "vertex_nodes", "face_vertexes"] + \
            [f"{op}.{part}" for op in ("operatorCurl", "operatorRBFinterp", "operatorDirectionalSecondDeriv") for part in ("indptr", "indices", "data")]


def run_case(name: str, level: int, over: dict, nsteps: int, every_step_dumps: bool, full_tables: bool,
             init_state: dict | None = None, nl_tables: bool = False, capture_start: bool = False):
    binary = os.path.join(ROOT, "oracle", "_ref", f"odis_ref_l{level}")
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(d + "/input_files"); os.makedirs(d + "/DATA")
        os.symlink(f"{REF_GRIDS}/grid_l{level}.txt", f"{d}/input_files/grid_l{level}.txt")
        over = dict(over)
        over["geodesic grid level"] = str(level)
        # first pass: learn totalIter for this time-step target so that the loop bound gives nsteps
        open(d + "/input.in", "w").write(input_text(dict(over, **{"simulation end time": "0"})))
        subprocess.run([binary, "--no-run"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        total = int(read_records(d + "/DATA/ref_tables.bin")["totalIter"][0])
        if nsteps > 0:
            over["simulation end time"] = repr((nsteps - 0.5) / total)
        else:                                   # whole orbits as configured: HDF5 gets int(endTime)*outputTime+1 rows
            nsteps = int(np.ceil(total * float(over["simulation end time"])))
        if every_step_dumps:
            over["output time"] = str(total)
        if init_state is not None:
            os.makedirs(d + "/InitialConditions")
            over["initial conditions"] = "LOAD"
            with open(d + "/InitialConditions/vel_init.txt", "w") as f:
                for i in range(init_state["v"].shape[0]):
                    f.write("%.17g, %.17g, %.17g, %.17g\n" % (init_state["v"][i], *init_state["dvdt"][i]))
            with open(d + "/InitialConditions/pres_init.txt", "w") as f:
                for i in range(init_state["eta"].shape[0]):
                    f.write("%.17g, %.17g, %.17g, %.17g\n" % (init_state["eta"][i], *init_state["detadt"][i]))
        start = None
        if capture_start:              # the state getInitialConditions built, before any step: a run whose loop bound is 0
            open(d + "/input.in", "w").write(input_text(dict(over, **{"simulation end time": "0"})))
            subprocess.run([binary, "--quiet-restart"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            start = read_records(d + "/DATA/ref_final.bin")
        text = input_text(over)
        open(d + "/input.in", "w").write(text)
        subprocess.run([binary], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tab = read_records(d + "/DATA/ref_tables.bin")
        fin = read_records(d + "/DATA/ref_final.bin")
        dumps = read_records(d + "/DATA/ref_dumps.bin")
        h5 = read_h5shim(d + "/DATA")
        out_txt = open(d + "/DATA/OUTPUT.txt").read()
    lat, lon, fr, cen = parse_grid_degrees(f"{REF_GRIDS}/grid_l{level}.txt")
    out = {"input_in": np.array(text), "level": np.array(level), "nsteps": np.array(nsteps),
           "grid_lat_deg": lat, "grid_lon_deg": lon, "grid_friends": fr, "grid_centroid_deg": cen,
           "output_txt": np.array(out_txt)}
    for s in SCALARS:
        out["scalar_" + s] = tab[s]
    for t in TABLES:
        out["sha256_" + t] = np.array(digest(tab[t]))
        if full_tables:
            out["table_" + t] = tab[t]
    for op in ("operatorGradient", "operatorDivergence", "operatorCoriolis", "operatorLinearDrag"):
        for part in ("indptr", "indices", "data"):
            out[f"sha256_{op}.{part}"] = np.array(digest(tab[f"{op}.{part}"]))
            if full_tables:
                out[f"table_{op}.{part}"] = tab[f"{op}.{part}"]
    if nl_tables:                       # what the nonlinear branch (advection; true) reads besides the linear tables
        for t in NL_TABLES:
            out["nl_" + t] = tab[t]
    for k, v in fin.items():
        out["final_" + k] = v
    # per-dump FP64 arrays: "<slice>:<tag>"
    diss, eta_d, vel_d, slices = [], [], [], []
    for k, v in dumps.items():
        sl, tag = k.split(":")
        if tag == "dissipation avg output":
            diss.append(v[0]); slices.append(int(sl))
        elif tag == "displacement output" and (not every_step_dumps or int(sl) in (1, nsteps + 1)) and len(eta_d) < 16:
            eta_d.append(v)
        elif tag == "velocity output" and (not every_step_dumps or int(sl) in (1, nsteps + 1)) and len(vel_d) < 16:
            vel_d.append(v.reshape(-1, 2))
    out["dump_slices"] = np.array(slices)
    out["dump_dissipation_avg"] = np.array(diss)
    out["dump_displacement"] = np.array(eta_d)
    out["dump_velocity_en"] = np.array(vel_d)
    if not every_step_dumps:
        for k, v in h5.items():
            out["h5_" + k] = v
    if init_state is not None:
        for k, v in init_state.items():
            out["init_" + k] = v
    if start is not None:
        for k, v in start.items():
            out["start_" + k] = v
    path = os.path.join(HERE, f"case_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: level {level}, {nsteps} steps, totalIter {total}, dumps {len(slices)} -> {os.path.getsize(path) / 1e6:.2f} MB")


def run_shipped_l6_checkpoints(name: str = "l6_shipped_verbatim", checkpoints=(10, 100, 1000, 2900)):
    """BASELINE config 0 at its real size: /root/reference/input.in UNCHANGED except for the end time (1 orbit = 2,900 steps instead of
    150 orbits) on the shipped grid_l6.txt (10,242 cells; advection true, velocity cartesian output true). One reference run per
    checkpoint (the loop bound is the only thing that changes), FP64 state kept at each: v and eta everywhere, both AB3 histories at
    step 1000. The grid itself is the one case_l6_obliqwest_earth.npz already holds, so it is not stored twice.
    NOTE (found by running it): the reference's own solver DIVERGES on this configuration — its dissipation grows orbit-resonantly and
    the state is NaN from some step between 1,160 and 1,450 on (its OUTPUT.txt prints "AVG DISS: -nan" from the 6th dump on). The
    checkpoint at 2,900 therefore records non-finite arrays; the comparison there is "non-finite in the same entries"."""
    level = 6
    binary = os.path.join(ROOT, "oracle", "_ref", f"odis_ref_l{level}")
    verbatim = []
    for line in open("/root/reference/input.in"):
        parts = [x.strip() for x in line.split(";")]
        if len(parts) >= 2 and parts[0] == "simulation end time":
            continue
        verbatim.append(line)
    out = {"level": np.array(level), "checkpoints": np.array(checkpoints), "grid_case": np.array("l6_obliqwest_earth")}
    total = None
    for n in checkpoints:
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(d + "/input_files"); os.makedirs(d + "/DATA")
            os.symlink(f"{REF_GRIDS}/grid_l{level}.txt", f"{d}/input_files/grid_l{level}.txt")
            if total is None:
                open(d + "/input.in", "w").write("".join(verbatim) + "simulation end time; 0; endTime;\n")
                subprocess.run([binary, "--no-run"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                tab = read_records(d + "/DATA/ref_tables.bin")
                total = int(tab["totalIter"][0])
                for s in SCALARS:
                    out["scalar_" + s] = tab[s]
                for t in TABLES + NL_TABLES:
                    out["sha256_" + t] = np.array(digest(tab[t]))
            end = "1" if n == total else repr((n - 0.5) / total)
            text = "".join(verbatim) + f"simulation end time; {end}; endTime;\n"
            open(d + "/input.in", "w").write(text)
            subprocess.run([binary, "--quiet-restart"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            fin = read_records(d + "/DATA/ref_final.bin")
            out[f"step{n}_v"], out[f"step{n}_eta"] = fin["v"], fin["eta"]
            if n == 1000:
                out[f"step{n}_dvdt"], out[f"step{n}_detadt"] = fin["dvdt"], fin["detadt"]
            if n == total:             # the whole orbit as ./ODIS would run it: progress lines and the dissipation row of data.h5
                out["input_in"] = np.array(text)
                out["output_txt"] = np.array(open(d + "/DATA/OUTPUT.txt").read())
                dumps = read_records(d + "/DATA/ref_dumps.bin")
                diss = [(int(k.split(":")[0]), v[0]) for k, v in dumps.items() if k.split(":")[1] == "dissipation avg output"]
                out["dump_slices"] = np.array([s for s, _ in diss])
                out["dump_dissipation_avg"] = np.array([v for _, v in diss])
                eta_d = [v for k, v in dumps.items() if k.split(":")[1] == "displacement output"]
                out["dump_displacement_first5"] = np.array(eta_d[:5])          # steps 0, 290, 580, 870, 1160: all finite
    out["nsteps"] = np.array(total)
    path = os.path.join(HERE, f"case_{name}.npz")
    np.savez_compressed(path, **out)
    finite = {n: bool(np.isfinite(out[f"step{n}_v"]).all() and np.isfinite(out[f"step{n}_eta"]).all()) for n in checkpoints}
    print(f"{name}: level {level}, totalIter {total}, checkpoints finite: {finite} -> {os.path.getsize(path) / 1e6:.2f} MB")


def random_state(level: int, seed: int):
    n = 10 * 4 ** (level - 1) + 2
    f = 3 * n - 6
    rng = np.random.default_rng(seed)
    return {"v": rng.uniform(-1, 1, f) * 0.05, "dvdt": rng.uniform(-1, 1, (f, 3)) * 1e-5,
            "eta": rng.uniform(-1, 1, n) * 2.0, "detadt": rng.uniform(-1, 1, (n, 3)) * 1e-3}


if __name__ == "__main__":
    built = build_reference((3, 4, 5, 6))
    assert len(built) == 4, built
    only = sys.argv[1:]                  # case names to (re)generate; none = all
    _run = run_case

    def run_case(name, *a, **k):         # noqa: F811
        if not only or name in only:
            _run(name, *a, **k)

    earth = {"radius": "6.37122e6", "angular velocity": "7.292e-5", "surface gravity": "9.80616", "semimajor axis": "671100000.0",
             "eccentricity": "0.01", "obliquity": "-2.0", "ocean thickness": "8e3", "time step": "30", "potential": "OBLIQ_WEST"}
    # (1) shipped default physics (Earth-like, OBLIQ_WEST) on the shipped L3 grid, linear path, full tables
    run_case("l3_obliqwest_earth", 3, earth, 200, every_step_dumps=True, full_tables=True)
    # (2) Enceladus free surface, ECC, shipped L4 grid, dumps every 29 steps incl. float32 HDF5 content
    run_case("l4_ecc_enceladus", 4, {"time step": "50", "output time": "100"}, 290, every_step_dumps=False, full_tables=False)
    # (3) shipped default on the shipped L6 grid (config 0 of BASELINE.json, advection off)
    run_case("l6_obliqwest_earth", 6, earth, 100, every_step_dumps=False, full_tables=False)
    # (4) restart path: AB3 from a loaded random state with history (3-level formula from step 0), FULL potential
    run_case("l3_full_loaded", 3, {"potential": "FULL", "time step": "80"}, 60, every_step_dumps=True, full_tables=False,
             init_state=random_state(3, 11))
    # (5) OBLIQ + quadratic-drag diagnostic
    run_case("l3_obliq_quadratic", 3, {"potential": "OBLIQ", "friction type": "QUADRATIC", "time step": "80"}, 50,
             every_step_dumps=True, full_tables=False)
    # (6) FULL2 under an ice shell (LID_LOVE: radius reduced by the shell, forcing at the outer radius)
    run_case("l4_full2_lidlove", 4, {"potential": "FULL2", "surface type": "LID_LOVE", "shell thickness": "23e3",
                                     "love reduction factor": "0.9", "time step": "50"}, 40, every_step_dumps=True, full_tables=False)
    # (8) one whole orbit on L3 with the HDF5 rows the reference wrote (11 slices) and dissipation output on
    run_case("l3_ecc_full_orbit", 3, {"time step": "100", "output time": "10", "simulation end time": "1", "dissipation output": "true"},
             0, every_step_dumps=False, full_tables=False)
    # (9) nonlinear branch: the shipped input.in physics WITH advection true (as shipped) on L3, and a loaded random state on L4
    shipped = dict({"radius": "6.37122e6", "angular velocity": "7.292e-5", "surface gravity": "9.80616", "semimajor axis": "671100000.0",
                    "eccentricity": "0.01", "obliquity": "-2.0", "ocean thickness": "8e3", "time step": "30", "potential": "OBLIQ_WEST"},
                   advection="true")
    run_case("l3_advection_shipped", 3, shipped, 150, every_step_dumps=True, full_tables=True, nl_tables=True)
    run_case("l4_advection_loaded", 4, {"advection": "true", "potential": "FULL", "time step": "40", "ocean thickness": "2e3"}, 40,
             every_step_dumps=True, full_tables=False, init_state=random_state(4, 21), nl_tables=True)
    # (11) Beuthe membrane shell (LID_MEMBR): g, radius and the tidal prefactor are replaced by membraneNuBeta's values
    run_case("l3_ecc_lidmembr", 3, {"surface type": "LID_MEMBR", "shell thickness": "10e3", "sh degree": "4", "time step": "60",
                                    "eccentricity": "0.0047"}, 40, every_step_dumps=True, full_tables=False)
    # (12) FREE_LOADING: loading Love numbers change the tidal prefactor (boundaryConditions.cpp:29-77)
    run_case("l3_obliq_freeloading", 3, {"surface type": "FREE_LOADING", "potential": "OBLIQ", "sh degree": "3", "time step": "70"}, 45,
             every_step_dumps=True, full_tables=False)
    # (13) `initial conditions; ANALYTICAL`: the reference's analytical OBLIQ_WEST response as the start state (initialConditions.cpp:146-208),
    #      captured before any step ("start_*") and after 80 steps
    run_case("l3_obliqwest_analytical", 3, dict(earth, **{"initial conditions": "ANALYTICAL", "friction coefficient": "1e-6"}), 80,
             every_step_dumps=True, full_tables=False, capture_start=True)
    # (14) PLANET forcing (moon-moon tides: a companion on the inner 2:1 orbit, tidalPotentials.cpp:176-225) on a Europa-like ocean
    run_case("l3_planet_europa", 3, {"potential": "PLANET", "radius": "1560.8e3", "angular velocity": "2.0478e-5", "surface gravity": "1.315",
                                     "semimajor axis": "671100000.0", "ocean thickness": "100e3", "orbital period": "306822.0", "time step": "200",
                                     "friction coefficient": "1e-6"}, 90, every_step_dumps=True, full_tables=False)
    # (10) the shipped input.in VERBATIM (advection true, velocity cartesian output true, ...) except for the grid level (3) and the
    #      end time (1 orbit = 48,100 steps at the shipped 30 s step): the whole-run drop-in check, HDF5 rows included
    verbatim = {}
    for line in open("/root/reference/input.in"):
        parts = [x.strip() for x in line.split(";")]
        if len(parts) >= 2 and parts[0] in BASE:
            verbatim[parts[0]] = parts[1]
    verbatim["simulation end time"] = "1"
    run_case("l3_shipped_verbatim", 3, verbatim, 0, every_step_dumps=False, full_tables=False)
    run_case("l5_advection_ecc", 5, {"advection": "true", "potential": "ECC", "time step": "20", "ocean thickness": "1e3", "eccentricity": "0.05",
                                     "friction type": "QUADRATIC", "friction coefficient": "1e-3"}, 60, every_step_dumps=False, full_tables=False,
             init_state=random_state(5, 31), nl_tables=True)
    # (15) BASELINE config 0 at its real size: the shipped input.in on the shipped L6 grid, FP64 checkpoints at steps 10 / 100 / 1000 / 2900
    if not only or "l6_shipped_verbatim" in only:
        run_shipped_l6_checkpoints()
    # (7) no forcing, decaying loaded state on L5
    run_case("l5_none_loaded", 5, {"potential": "NONE", "time step": "20"}, 30, every_step_dumps=False, full_tables=False,
             init_state=random_state(5, 5))
