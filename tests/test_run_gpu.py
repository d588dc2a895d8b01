"""The whole-run entry point (odis_run = `./ODIS` in a run directory) against the reference's own run of the
same directory: data.h5 content, the OUTPUT.txt progress lines and the restart files."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case, make_run_dir
from h5lite_reader import read_h5

pytestmark = pytest.mark.gpu


def dumping_lines(text: str):
    return [l for l in text.splitlines() if l.startswith("DUMPING DATA AT")]


def test_full_orbit_run_matches_reference_outputs(odis, tmp_path):
    case = load_case("l3_ecc_full_orbit")
    d = make_run_dir(tmp_path, case)
    res = odis.run(d)
    assert res["steps"] == int(case["nsteps"]) == 1200 and res["dumps"] == 11 and res["interrupted"] == 0
    assert res["steps_per_period"] == int(case["scalar_totalIter"][0]) and res["dt"] == float(case["scalar_timeStep"][0])
    assert res["kernel_launches"] >= 1200
    h5 = read_h5(os.path.join(d, "DATA", "data.h5"))
    ref = {k[3:]: case[k] for k in case if k.startswith("h5_")}
    assert sorted(h5) == sorted(ref)                                   # same dataset names (src/outFiles.cpp:250-338,500,509)
    for name, r in ref.items():
        assert h5[name].dtype == np.float32 and h5[name].shape == r.shape, name
    # eta is bit-identical in FP64, so its float32 rows are too; grid positions likewise
    for name in ("displacement", "face longitude", "face latitude"):
        assert np.array_equal(h5[name], ref[name]), name
    # v_avg / dissipation use a different 10-point summation order and a parallel sum: float32 round-off at most
    for name in ("east velocity", "north velocity", "dissipated energy", "dissipation avg output"):
        # components are sums with cancellation, so the bound is relative to the field's magnitude (float32 eps = 6e-8)
        assert np.abs(h5[name] - ref[name]).max() <= 2e-7 * np.abs(ref[name]).max(), name
    # the progress lines are the de-facto status API (parsed by python_scripts/dissipation_progress.py)
    out = open(os.path.join(d, "DATA", "OUTPUT.txt")).read()
    assert dumping_lines(out) == dumping_lines(str(case["output_txt"]))
    assert "Calculations appear to have finished!" in out
    # restart files: "%1.6E" text of the final state (src/initialConditions.cpp:209-276)
    vel = np.array([[float(x) for x in re.split(r",\s*", l.strip())] for l in open(os.path.join(d, "InitialConditions", "vel_init.txt"))])
    pres = np.array([[float(x) for x in re.split(r",\s*", l.strip())] for l in open(os.path.join(d, "InitialConditions", "pres_init.txt"))])
    assert vel.shape == (480, 4) and pres.shape == (162, 4)
    fmt = lambda a: np.array([float("%1.6E" % x) for x in a.ravel()]).reshape(a.shape)
    assert np.array_equal(vel[:, 0], fmt(case["final_v"])) and np.array_equal(vel[:, 1:], fmt(case["final_dvdt"]))
    assert np.array_equal(pres[:, 0], fmt(case["final_eta"])) and np.array_equal(pres[:, 1:], fmt(case["final_detadt"]))


def test_shipped_input_in_runs_verbatim(odis, tmp_path):
    """BASELINE config 0: the reference's shipped input.in as it is (advection true -> nonlinear branch, velocity cartesian
    output true, OBLIQ_WEST, Earth-like) on the shipped L3 grid for one orbit, against the reference's own run of it."""
    case = load_case("l3_shipped_verbatim")
    assert "advection; \t true;" in str(case["input_in"]) and "velocity cartesian output; \t true;" in str(case["input_in"])
    d = make_run_dir(tmp_path, case)
    res = odis.run(d)
    assert res["steps"] == int(case["nsteps"]) == 2900 and res["dumps"] == 11 and res["interrupted"] == 0
    h5 = read_h5(os.path.join(d, "DATA", "data.h5"))
    ref = {k[3:]: case[k] for k in case if k.startswith("h5_")}
    assert sorted(h5) == sorted(ref) and "x velocity" in h5
    for name, r in ref.items():
        assert h5[name].dtype == np.float32 and h5[name].shape == r.shape, name
    # eta and v are bit-identical in FP64 on the nonlinear branch too; the Cartesian velocity is operatorRBFinterp * v in the
    # reference's own order
    for name in ("displacement", "x velocity", "y velocity", "z velocity", "face longitude", "face latitude"):
        assert np.array_equal(h5[name], ref[name]), name
    for name in ("east velocity", "north velocity", "dissipation avg output"):
        assert np.abs(h5[name] - ref[name]).max() <= 2e-7 * np.abs(ref[name]).max(), name
    out = open(os.path.join(d, "DATA", "OUTPUT.txt")).read()
    assert dumping_lines(out) == dumping_lines(str(case["output_txt"]))
    vel = np.array([[float(x) for x in re.split(r",\s*", l.strip())] for l in open(os.path.join(d, "InitialConditions", "vel_init.txt"))])
    pres = np.array([[float(x) for x in re.split(r",\s*", l.strip())] for l in open(os.path.join(d, "InitialConditions", "pres_init.txt"))])
    fmt = lambda a: np.array([float("%1.6E" % x) for x in a.ravel()]).reshape(a.shape)
    assert np.array_equal(vel[:, 0], fmt(case["final_v"])) and np.array_equal(vel[:, 1:], fmt(case["final_dvdt"]))
    assert np.array_equal(pres[:, 0], fmt(case["final_eta"])) and np.array_equal(pres[:, 1:], fmt(case["final_detadt"]))


def test_restart_run_continues_from_files(odis, tmp_path):
    """initial conditions; LOAD reads InitialConditions/*.txt and uses the 3-level AB3 formula from step 0."""
    case = load_case("l3_full_loaded")
    d = make_run_dir(tmp_path, case)
    os.makedirs(os.path.join(d, "InitialConditions"))
    with open(os.path.join(d, "InitialConditions", "vel_init.txt"), "w") as f:
        for i in range(480):
            f.write("%.17g, %.17g, %.17g, %.17g\n" % (case["init_v"][i], *case["init_dvdt"][i]))
    with open(os.path.join(d, "InitialConditions", "pres_init.txt"), "w") as f:
        for i in range(162):
            f.write("%.17g, %.17g, %.17g, %.17g\n" % (case["init_eta"][i], *case["init_detadt"][i]))
    res = odis.run(d)
    assert res["steps"] == int(case["nsteps"])
    out = open(os.path.join(d, "DATA", "OUTPUT.txt")).read()
    ref_lines = dumping_lines(str(case["output_txt"]))
    assert dumping_lines(out) == ref_lines and len(ref_lines) == 61
    h5 = read_h5(os.path.join(d, "DATA", "data.h5"))
    assert h5["displacement"].shape == (1, 162)                        # int(endTime)*outputTime+1 rows (src/outFiles.cpp:179)
    assert np.array_equal(h5["displacement"][0], case["init_eta"].astype(np.float32))


def test_unsupported_configurations_fail_loudly(odis, tmp_path):
    case = load_case("l3_obliqwest_earth")
    d = make_run_dir(tmp_path, case)
    text = open(os.path.join(d, "input.in")).read().replace("solver type; \t AB3;", "solver type; \t RK4;")
    open(os.path.join(d, "input.in"), "w").write(text)
    with pytest.raises(odis.OdisError) as e:
        odis.run(d)
    assert e.value.code == -6 and "AB3" in str(e.value)
    assert "TERMINATING ODIS." in open(os.path.join(d, "DATA", "ERROR.txt")).read()


def test_cli_binary(odis, tmp_path):
    case = load_case("l3_obliq_quadratic")
    d = make_run_dir(tmp_path, case)
    exe = os.path.join(ROOT, "geodesicodis_b200", "bin", "ODIS")
    r = subprocess.run([exe, "--quiet"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(os.path.join(d, "DATA", "OUTPUT.txt")).read()
    assert dumping_lines(out) == dumping_lines(str(case["output_txt"]))
    r = subprocess.run([exe, "--quiet", "--dir", os.path.join(d, "nowhere")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ODIS HAS FOUND AN ERROR" in r.stdout          # the reference exits 0 from TerminateODIS


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_whole_run_writes_the_same_files(odis, tmp_path, world):
    """`ODIS --gpus N` (odis_run with n_gpus: one partitioned solver per GPU, driven from the one process): the same run directory gives the
    same data.h5, progress lines and restart files as on one GPU — bit for bit, the halo exchange changes where a value lives, never its
    arithmetic (the energy sum associates per rank: float32 round-off at most). Reference: src/main.cpp:46-64 -> solveODIS -> ab3Explicit."""
    from test_multigpu import _device_count
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    case = load_case("l4_ecc_enceladus")
    d1 = make_run_dir(tmp_path / "one", case)
    dn = make_run_dir(tmp_path / "many", case)
    r1 = odis.run(d1)
    rn = odis.run(dn, n_gpus=world)
    assert rn["steps"] == r1["steps"] == int(case["nsteps"]) and rn["dumps"] == r1["dumps"] and rn["dt"] == r1["dt"]
    h1, hn = read_h5(os.path.join(d1, "DATA", "data.h5")), read_h5(os.path.join(dn, "DATA", "data.h5"))
    assert sorted(h1) == sorted(hn)
    for name in h1:
        if name in ("dissipation avg output",):
            assert np.abs(hn[name] - h1[name]).max() <= 2e-7 * np.abs(h1[name]).max(), name
        else:
            assert np.array_equal(hn[name], h1[name]), name
    out1, outn = open(os.path.join(d1, "DATA", "OUTPUT.txt")).read(), open(os.path.join(dn, "DATA", "OUTPUT.txt")).read()
    assert dumping_lines(outn) == dumping_lines(out1) and "grid partitioned over %d GPUs" % world in outn
    for f in ("vel_init.txt", "pres_init.txt"):
        assert open(os.path.join(dn, "InitialConditions", f)).read() == open(os.path.join(d1, "InitialConditions", f)).read()
    # against the reference's own float32 rows too
    ref = {k[3:]: case[k] for k in case if k.startswith("h5_")}
    assert np.array_equal(hn["displacement"], ref["displacement"])
    # ... and with the dumps overlapped with stepping (`ODIS --gpus N --overlap-output`): every rank's own entries come back compact
    # through its snapshot slots and are placed by its partition map — byte for byte the files of the synchronous partitioned run
    do = make_run_dir(tmp_path / "many_overlapped", case)
    ro = odis.run(do, n_gpus=world, overlap_output=True)
    assert ro["steps"] == rn["steps"] and ro["dumps"] == rn["dumps"]
    ho = read_h5(os.path.join(do, "DATA", "data.h5"))
    for name in hn:
        assert np.array_equal(ho[name], hn[name]), name
    assert dumping_lines(open(os.path.join(do, "DATA", "OUTPUT.txt")).read()) == dumping_lines(outn)
    for f in ("vel_init.txt", "pres_init.txt"):
        assert open(os.path.join(do, "InitialConditions", f)).read() == open(os.path.join(dn, "InitialConditions", f)).read()
