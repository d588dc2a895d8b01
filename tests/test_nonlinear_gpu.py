"""Nonlinear branch (`advection; true`) on the GPU against the reference's own solver output (golden fixtures
written by the unmodified reference run with advection on) and against the CPU oracle.

The operators only this branch uses (operatorCurl, operatorRBFinterp, operatorDirectionalSecondDeriv) come from the fixture, as
the reference built them; the kernels follow the reference's operation order, so v, eta and both AB3 histories must be
BIT-IDENTICAL. The dissipation sum is a parallel tree instead of the serial loop: 1e-12 relative."""
import os

import numpy as np
import pytest

from conftest import NL_CASES, case_params, load_case, make_run_dir, nonlinear_tables

pytestmark = pytest.mark.gpu


def solver_for(odis, tmp_path, case, reorder=1):
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l%d.txt" % int(case["level"])), float(case["scalar_radius"][0]))
    loaded = "init_v" in case
    prm = case_params(case, init_load=int(loaded))
    s = odis.Solver(mesh, dict(prm, reorder=reorder, semimajor_axis=0.0))
    s.enable_advection(nonlinear_tables(case))
    if loaded:
        s.set_state(case["init_v"], case["init_eta"], case["init_dvdt"], case["init_detadt"])
    return mesh, prm, s


@pytest.mark.parametrize("reorder", [1, 0])
@pytest.mark.parametrize("name", NL_CASES)
def test_nonlinear_step_matches_reference_solver(odis, tmp_path, name, reorder):
    case = load_case(name)
    assert int(case["scalar_advection"][0]) == 1
    mesh, prm, s = solver_for(odis, tmp_path, case, reorder)
    n = int(case["nsteps"])
    s.step(n // 3); s.step(n - n // 3)
    for fid, key in ((odis.FIELD_VELOCITY, "final_v"), (odis.FIELD_ETA, "final_eta"), (odis.FIELD_DVDT, "final_dvdt"), (odis.FIELD_DETADT, "final_detadt")):
        got = s.field(fid)
        rel = float(np.abs(got - case[key]).max() / max(np.abs(case[key]).max(), 1e-300))
        assert rel <= 1e-10, (key, rel)                       # BASELINE.json's bar
        assert np.array_equal(got, case[key]), (key, rel)     # what the kernels are built for
    total, out_time = int(case["scalar_totalIter"][0]), int(case["scalar_outputTime"][0])
    out_freq = total // out_time
    series = s.dissipation_series()
    assert series.shape == (n + 1,)
    assert np.allclose(series[(case["dump_slices"] - 1) * out_freq], case["dump_dissipation_avg"], rtol=1e-12, atol=0.0)


def test_nonlinear_matches_oracle_and_differs_from_linear(odis, tmp_path):
    """Random loaded state: GPU vs the CPU oracle step by step in two chunks, and the nonlinear terms are not a no-op."""
    from oracle.lte_oracle import LteOracle
    case = load_case("l4_advection_loaded")
    mesh, prm, s = solver_for(odis, tmp_path, case)
    o = LteOracle(mesh.tables, prm)
    o.set_nonlinear(nonlinear_tables(case))
    st = (case["init_v"], case["init_eta"], case["init_dvdt"], case["init_detadt"])
    o.set_state(*st, iter=11)
    s.set_state(*st, iter=11)
    o.step(25)
    s.step(10); s.step(15)
    for fid in range(4):
        assert np.array_equal(s.field(fid), o.field(fid)), fid
    lin = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    lin.set_state(*st, iter=11)
    lin.step(25)
    assert np.abs(lin.field(odis.FIELD_ETA) - s.field(odis.FIELD_ETA)).max() > 1e-6 * np.abs(s.field(odis.FIELD_ETA)).max()


def test_enable_advection_errors(odis, tmp_path):
    case = load_case("l3_advection_shipped")
    mesh, prm, s = solver_for(odis, tmp_path, case)
    with pytest.raises(odis.OdisError):
        s.enable_advection(nonlinear_tables(case))            # already on
    other = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    bad = dict(nonlinear_tables(case))
    bad["operatorCurl.indptr"] = bad["operatorCurl.indptr"][:-1]
    with pytest.raises(ValueError):
        other.enable_advection(bad)
    late = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    late.enable_self_gravity(2, [0.0, 0.0, 0.5])             # the folded harmonic analysis belongs to the linear cell update
    with pytest.raises(odis.OdisError):
        late.enable_advection(nonlinear_tables(case))
