"""The nonlinear step in 6 gather launches (the baseline selection, kernel_select=1; the default since round 2 is the 4-launch step —
vertex potential vorticity + cell kinetic energy in one grid, thickness flux inside the edge update — measured +11 % on a B200).
Same arithmetic, so the reference's nonlinear solver output must be reproduced bit for bit by both (tests/test_nonlinear_gpu.py
runs the default)."""
import os

import numpy as np
import pytest

from conftest import NL_CASES, case_params, load_case, make_run_dir, nonlinear_tables

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("reorder", [1, 0])
@pytest.mark.parametrize("name", NL_CASES)
def test_six_launch_nonlinear_step_matches_reference_solver(odis, tmp_path, name, reorder):
    case = load_case(name)
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l%d.txt" % int(case["level"])), float(case["scalar_radius"][0]))
    loaded = "init_v" in case
    prm = case_params(case, init_load=int(loaded))
    s = odis.Solver(mesh, dict(prm, reorder=reorder, semimajor_axis=0.0, kernel_select=1))
    s.enable_advection(nonlinear_tables(case))
    if loaded:
        s.set_state(case["init_v"], case["init_eta"], case["init_dvdt"], case["init_detadt"])
    n = int(case["nsteps"])
    l0 = s.launches
    s.step(n // 3); s.step(n - n // 3)
    assert s.launches - l0 == 8 * n                            # diagnostics + 6 gather launches + potential pass
    for fid, key in ((odis.FIELD_VELOCITY, "final_v"), (odis.FIELD_ETA, "final_eta"), (odis.FIELD_DVDT, "final_dvdt"), (odis.FIELD_DETADT, "final_detadt")):
        assert np.array_equal(s.field(fid), case[key]), key
    total, out_time = int(case["scalar_totalIter"][0]), int(case["scalar_outputTime"][0])
    out_freq = total // out_time
    assert np.allclose(s.dissipation_series()[(case["dump_slices"] - 1) * out_freq], case["dump_dissipation_avg"], rtol=1e-12, atol=0.0)
