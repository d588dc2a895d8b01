"""`potential; PLANET` (moon-moon tides: a companion on the inner 2:1 orbit, src/tidalPotentials.cpp:176-225). The step kernels
leave U = 0 for this type and one more pass over the cells writes it (launch_planet_potential), so the default kernels are the
ones that were profiled. Checked against the reference's own run (golden case l3_planet_europa) and the oracle."""
import os

import numpy as np
import pytest

from conftest import case_params, load_case, make_run_dir
from oracle.lte_oracle import LteOracle

pytestmark = pytest.mark.gpu


def setup(odis, tmp_path, **over):
    case = load_case("l3_planet_europa")
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l3.txt"), float(case["scalar_radius"][0]))
    prm = case_params(case)
    assert prm["potential"] == 13 and prm["semimajor_axis"] == 671100000.0
    return case, d, mesh, prm


@pytest.mark.parametrize("select", [0, 1, 8, 128])             # default; direct-load baseline kernels; no graphs; 32-bit stencil ids
@pytest.mark.parametrize("reorder", [1, 0])
def test_planet_forcing_matches_reference_solver(odis, tmp_path, reorder, select):
    case, d, mesh, prm = setup(odis, tmp_path)
    s = odis.Solver(mesh, dict(prm, reorder=reorder, kernel_select=select))
    n = int(case["nsteps"])
    l0 = s.launches
    s.step(n // 2); s.step(n - n // 2)
    assert s.launches - l0 == 3 * n                            # edge, cell, PLANET pass
    for fid, key in ((odis.FIELD_VELOCITY, "final_v"), (odis.FIELD_ETA, "final_eta"), (odis.FIELD_DVDT, "final_dvdt"), (odis.FIELD_DETADT, "final_detadt")):
        assert np.array_equal(s.field(fid), case[key]), key
    assert np.allclose(s.dissipation_series(), case["dump_dissipation_avg"], rtol=1e-12, atol=0.0)
    assert np.abs(s.field(odis.FIELD_ETA)).max() > 0.1          # the companion raises a real tide


def test_planet_forcing_operator_and_errors(odis, tmp_path):
    case, d, mesh, prm = setup(odis, tmp_path)
    s, o = odis.Solver(mesh, dict(prm, reorder=1)), LteOracle(mesh.tables, prm)
    for t in (prm["dt"], 777 * prm["dt"]):
        assert np.array_equal(s.forcing(t), o.forcing(t))
    with pytest.raises(odis.OdisError):
        odis.Solver(mesh, dict(prm, semimajor_axis=0.0))        # PLANET needs the companion's orbit


def test_whole_run_with_planet_forcing(odis, tmp_path):
    case, d, mesh, prm = setup(odis, tmp_path)
    res = odis.run(d)
    assert res["steps"] == int(case["nsteps"]) and res["dumps"] == len(case["dump_slices"])
    ours = [l for l in open(os.path.join(d, "DATA", "OUTPUT.txt")).read().splitlines() if l.startswith("DUMPING DATA AT")]
    ref = [l for l in str(case["output_txt"]).splitlines() if l.startswith("DUMPING DATA AT")]
    assert ours == ref
