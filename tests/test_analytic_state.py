"""`initial conditions; ANALYTICAL` (odis_analytical_state, csrc/odis_analytic.cpp) against the state the reference's own
getInitialConditions -> analyticalInitialConditions built (golden case l3_obliqwest_analytical, keys start_*), bit for bit;
and the oracle stepped from it against the reference's final state."""
import os

import numpy as np
import pytest

from conftest import case_params, load_case, make_run_dir
from oracle.lte_oracle import LteOracle


def mesh_and_params(odis, tmp_path, case):
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l%d.txt" % int(case["level"])), float(case["scalar_radius"][0]))
    return mesh, case_params(case, init_load=0)


def test_analytical_state_is_the_references(odis, tmp_path):
    case = load_case("l3_obliqwest_analytical")
    mesh, prm = mesh_and_params(odis, tmp_path, case)
    v, dv, eta, de = odis.analytical_state(mesh, dict(prm, semimajor_axis=0.0, reorder=1))
    assert np.array_equal(v, case["start_v"]) and np.array_equal(dv, case["start_dvdt"])
    assert np.array_equal(eta, case["start_eta"]) and np.array_equal(de, case["start_detadt"])
    assert np.abs(eta).max() > 1.0 and np.abs(v).max() > 1.0            # a real tide, not zeros
    # the tendencies are time derivatives of a rotating pattern: level 0 and level 1 differ by O(omega dt)
    assert 0 < np.abs(de[:, 0] - de[:, 1]).max() < 0.05 * np.abs(de[:, 0]).max()


def test_oracle_from_the_analytical_state_reproduces_the_reference_run(odis, tmp_path):
    case = load_case("l3_obliqwest_analytical")
    mesh, prm = mesh_and_params(odis, tmp_path, case)
    o = LteOracle(mesh.tables, prm)                                      # init_load = 0: Euler / two-level start-up (temporalOperators.cpp:36)
    o.set_state(*[case[k] for k in ("start_v", "start_eta", "start_dvdt", "start_detadt")])
    e0 = o.dissipation_avg()
    series = o.step(int(case["nsteps"]))
    for fid, key in ((0, "final_v"), (1, "final_eta"), (2, "final_dvdt"), (3, "final_detadt")):
        assert np.array_equal(o.field(fid), case[key]), key
    assert np.array_equal(np.concatenate([[e0], series]), case["dump_dissipation_avg"])


def test_analytical_state_is_close_to_the_numerical_steady_state(odis, tmp_path):
    """Known-answer check ('approximate'): started on the analytical solution, the discrete solution stays near it —
    after 80 steps eta differs from the analytical pattern advanced in time by a few per cent of its amplitude at most."""
    case = load_case("l3_obliqwest_analytical")
    mesh, prm = mesh_and_params(odis, tmp_path, case)
    eta0, eta1 = case["start_eta"], case["final_eta"]
    # the response is a pattern rotating westward at the spin rate: eta(lon, t) = eta(lon + omega t, 0) -> compare amplitudes
    assert abs(np.abs(eta1).max() / np.abs(eta0).max() - 1.0) < 0.05
    assert abs(np.sqrt((eta1 ** 2).mean()) / np.sqrt((eta0 ** 2).mean()) - 1.0) < 0.05


def test_analytical_state_errors(odis, tmp_path):
    case = load_case("l3_obliqwest_analytical")
    mesh, prm = mesh_and_params(odis, tmp_path, case)
    with pytest.raises(odis.OdisError) as e:
        odis.analytical_state(mesh, dict(prm, potential=5, semimajor_axis=0.0, reorder=1))      # ECC: the reference has no solution
    assert e.value.code == -6
