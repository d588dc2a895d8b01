"""Drop-in proof at the reference's own call surface: the UNMODIFIED reference program (its main sequence, Globals, Mesh,
OutFiles, getInitialConditions, restart writer — compiled from /root/reference by oracle/ref_build/Makefile) linked against
the reference-side bindings of integration/ and libodis_b200.so:

  oracle/_ref/odis_hybrid_l<L>      integration/timeIntegrator_b200.cpp replaces src/timeIntegrator.cpp
                                    (ab3Explicit -> odis_set_state / odis_step / odis_get_field)
  oracle/_ref/odis_hybridops_l<L>   the reference's ab3Explicit kept; updateMomentum, updateEta, forcing, integrateAB3scalar,
                                    interpolateVelocity, updateEnergy supplied by integration/operators_b200.cpp (odis_op_*)

Each is run in a directory reproducing a golden case and must leave what the all-CPU reference left: v, eta and both AB3
histories bit for bit; the displacement dumps bit for bit; velocity components to 1e-13 and the dissipation series to 1e-12
(summation order of the 10-point stencil / tree sum, as in tests/test_step_parity_gpu.py)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case, make_run_dir
from oracle.refio import read_h5shim, read_records

pytestmark = pytest.mark.gpu


def run_binary(tmp_path, case, exe, env=None):
    path = os.path.join(ROOT, "oracle", "_ref", f"{exe}_l{int(case['level'])}")
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the reference tree at build time)")
    d = make_run_dir(tmp_path, case)
    r = subprocess.run([path, "--quiet-restart"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    err = open(os.path.join(d, "DATA", "ERROR.txt")).read() if os.path.exists(os.path.join(d, "DATA", "ERROR.txt")) else ""
    assert "TERMINATING" not in err, err[-2000:]
    assert r.returncode == 0, r.stdout[-2000:]
    data = os.path.join(d, "DATA")
    return read_records(os.path.join(data, "ref_final.bin")), read_records(os.path.join(data, "ref_dumps.bin")), data


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_against_case(case, fin, dumps, velocity_rtol):
    assert np.array_equal(fin["v"], case["final_v"])
    assert np.array_equal(fin["eta"], case["final_eta"])
    assert np.array_equal(fin["dvdt"], case["final_dvdt"])
    assert np.array_equal(fin["detadt"], case["final_detadt"])
    slices = [int(k.split(":")[0]) for k in dumps if k.endswith(":dissipation avg output")]
    assert slices == list(case["dump_slices"])
    diss = np.array([dumps[f"{sl}:dissipation avg output"][0] for sl in slices])
    assert np.allclose(diss, case["dump_dissipation_avg"], rtol=1e-12, atol=0.0)
    last = slices[-1]
    assert np.array_equal(dumps[f"{last}:displacement output"], case["dump_displacement"][-1])
    ven = dumps[f"{last}:velocity output"].reshape(-1, 2)
    assert rel_err(ven, case["dump_velocity_en"][-1]) <= velocity_rtol


@pytest.mark.parametrize("name", ["l3_obliqwest_earth", "l4_ecc_enceladus", "l3_obliq_quadratic", "l4_full2_lidlove", "l3_ecc_lidmembr",
                                  "l3_advection_shipped"])
def test_reference_program_with_the_device_time_loop(tmp_path, name):
    case = load_case(name)
    fin, dumps, data = run_binary(tmp_path, case, "odis_hybrid")
    check_against_case(case, fin, dumps, 1e-13)
    log = open(os.path.join(data, "OUTPUT.txt")).read()
    assert log.count("DUMPING DATA AT") == len(case["dump_slices"])
    if "h5_displacement" in case:                    # the float32 rows the reference's DumpData handed to HDF5
        h5 = read_h5shim(data)
        assert np.array_equal(h5["displacement"], case["h5_displacement"])
        assert np.allclose(h5["east velocity"], case["h5_east velocity"], rtol=1e-6, atol=1e-30)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", ["l4_ecc_enceladus", "l3_obliq_quadratic"])
def test_reference_program_with_the_device_time_loop_on_several_gpus(tmp_path, name, world):
    """The same unmodified reference program with ODIS_B200_GPUS=N: integration/odis_b200_bridge.cpp cuts the reference's own Mesh into N
    parts and drives N partitioned solvers from the reference's one process (what src/main.cpp:46-64 can reach). Same bars as on one GPU:
    state and restart arrays bit for bit, dissipation to 1e-12."""
    from test_multigpu import _device_count
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    case = load_case(name)
    fin, dumps, data = run_binary(tmp_path, case, "odis_hybrid", env={"ODIS_B200_GPUS": str(world)})
    check_against_case(case, fin, dumps, 1e-13)
    log = open(os.path.join(data, "OUTPUT.txt")).read()
    assert log.count("DUMPING DATA AT") == len(case["dump_slices"]) and f"grid partitioned over {world} GPUs" in log


@pytest.mark.parametrize("name", ["l3_obliq_quadratic", "l4_full2_lidlove", "l3_ecc_lidmembr"])
def test_reference_loop_with_device_operators(tmp_path, name):
    case = load_case(name)
    fin, dumps, _ = run_binary(tmp_path, case, "odis_hybridops")
    check_against_case(case, fin, dumps, 1e-13)


def test_device_operators_refuse_the_nonlinear_branch(tmp_path):
    """No CPU fallback behind the wrappers: `advection; true` ends through the reference's own fatal path."""
    case = load_case("l3_advection_shipped")
    path = os.path.join(ROOT, "oracle", "_ref", "odis_hybridops_l3")
    if not os.path.exists(path):
        pytest.skip("hybrid binary not built")
    d = make_run_dir(tmp_path, case)
    subprocess.run([path, "--quiet-restart"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    err = open(os.path.join(d, "DATA", "ERROR.txt")).read()
    assert "linear branch" in err and "TERMINATING ODIS" in err
    assert not os.path.exists(os.path.join(d, "DATA", "ref_final.bin"))
