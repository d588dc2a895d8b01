"""`initial conditions; ANALYTICAL` on the device: the solver started from odis_analytical_state and the whole-run entry point
(`odis_run`, which the reference's getInitialConditions branch corresponds to) against the reference's own run."""
import os

import numpy as np
import pytest

from conftest import case_params, load_case, make_run_dir

pytestmark = pytest.mark.gpu


def test_device_run_from_the_analytical_state(odis, tmp_path):
    case = load_case("l3_obliqwest_analytical")
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l3.txt"), float(case["scalar_radius"][0]))
    prm = dict(case_params(case, init_load=0), semimajor_axis=0.0, reorder=1)
    v, dv, eta, de = odis.analytical_state(mesh, prm)
    s = odis.Solver(mesh, prm)
    s.set_state(v, eta, dv, de)
    s.step(int(case["nsteps"]))
    for fid, key in ((odis.FIELD_VELOCITY, "final_v"), (odis.FIELD_ETA, "final_eta"), (odis.FIELD_DVDT, "final_dvdt"),
                     (odis.FIELD_DETADT, "final_detadt")):
        assert np.array_equal(s.field(fid), case[key]), key
    assert np.allclose(s.dissipation_series(), case["dump_dissipation_avg"], rtol=1e-12, atol=0.0)


def test_whole_run_with_analytical_initial_conditions(odis, tmp_path):
    case = load_case("l3_obliqwest_analytical")
    d = make_run_dir(tmp_path, case)
    res = odis.run(d)
    assert res["steps"] == int(case["nsteps"]) and res["dumps"] == len(case["dump_slices"])
    assert abs(res["last_dissipation_avg"] - case["dump_dissipation_avg"][-1]) <= 1e-12 * case["dump_dissipation_avg"][-1]
    ours = [l for l in open(os.path.join(d, "DATA", "OUTPUT.txt")).read().splitlines() if l.startswith("DUMPING DATA AT")]
    ref = [l for l in str(case["output_txt"]).splitlines() if l.startswith("DUMPING DATA AT")]
    assert ours == ref
