"""The operator surface (odis_op_*: the free functions the reference's loop calls, src/timeIntegrator.cpp:205-313, one call
each) against the CPU oracle's routines of the same names, and composed by hand into whole time steps against the
reference's own final state.

Bar: bit-identical for updateMomentum / updateEta / forcing / integrateAB3scalar (same kernels and operation order as the
fused step); 1e-13 relative for interpolateVelocity / updateEnergy, whose 10-point sums run in the CSR column order of the
step kernels instead of the reference's table order (the state does not depend on them)."""
import os

import numpy as np
import pytest

from conftest import case_params, load_case, make_run_dir
from oracle.lte_oracle import LteOracle

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def pair_for_case(odis, tmp_path, case, reorder=1):
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l%d.txt" % int(case["level"])), float(case["scalar_radius"][0]))
    prm = case_params(case, init_load=int("init_v" in case))
    return mesh, odis.Solver(mesh, dict(prm, reorder=reorder, semimajor_axis=0.0)), LteOracle(mesh.tables, prm), prm


@pytest.mark.parametrize("potential,friction", [(5, 0), (0, 1), (1, 0), (8, 0), (9, 1), (16, 0)])
def test_each_operator_matches_the_oracle(odis, potential, friction):
    pos, fr, cen = odis.generate_grid(5)                                  # 2,562 cells
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=30.0, radius=r, omega=5.307e-5, love_reduct=0.95, ecc=0.0047, obl=0.002,
               shell_thickness=0.0, potential=potential, friction=friction, surface=0, init_load=0)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    o = LteOracle(mesh.tables, prm)
    rng = np.random.default_rng(4321)
    v, eta = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    assert np.array_equal(s.updateMomentum(v, eta), o.updateMomentum(v, eta))
    assert np.array_equal(s.updateEta(v), o.updateEta(v))
    for t in (30.0, 12345.0 * 30.0 + 30.0):
        assert np.array_equal(s.forcing(t), o.forcing(t))
    for n, it in ((mesh.n_edges, 0), (mesh.n_cells, 1), (mesh.n_edges, 2), (7, 5)):
        sol, hist = rng.uniform(-1, 1, n), rng.uniform(-1, 1, (n, 3)) * 1e-3
        a, b = s.integrateAB3scalar(sol, hist, it), o.integrateAB3scalar(sol, hist, it)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    va_s, va_o = s.interpolateVelocity(v), o.interpolateVelocity(v)
    assert rel_err(va_s, va_o) <= 1e-13
    area = mesh.tables["face_area"]
    (ef_s, avg_s), (ef_o, avg_o) = s.updateEnergy(va_o, area), o.updateEnergy(va_o, area)
    assert np.array_equal(ef_s, ef_o)                                    # per-edge flux: same expression on the same input
    assert abs(avg_s - avg_o) <= 1e-12 * abs(avg_o)                      # tree sum vs serial sum
    # the solver's state was the scratch space: stepping needs a fresh odis_set_state
    with pytest.raises(odis.OdisError) as e:
        s.step(1)
    assert e.value.code == -7
    s.set_state(v, eta)
    o.set_state(v, eta)
    s.step(3)
    o.step(3)
    assert np.array_equal(s.field(odis.FIELD_VELOCITY), o.field(0)) and np.array_equal(s.field(odis.FIELD_ETA), o.field(1))


@pytest.mark.parametrize("reorder", [1, 0])
@pytest.mark.parametrize("name", ["l3_obliqwest_earth", "l3_full_loaded"])
def test_operator_calls_compose_to_the_reference_state(odis, tmp_path, name, reorder):
    """The reference's loop written out with one device call per function (what integration/operators_b200.cpp gives a
    maintainer who keeps ab3Explicit): final v, eta and both histories equal the reference solver's, bit for bit. The
    drag/forcing-gradient product of timeIntegrator.cpp:219 has no function of its own in the reference; it stays with
    the caller (here: the oracle's restatement)."""
    case = load_case(name)
    mesh, s, o, prm = pair_for_case(odis, tmp_path, case, reorder)
    if "init_v" in case:
        v, eta, dv, de = (np.array(case[k]) for k in ("init_v", "init_eta", "init_dvdt", "init_detadt"))
    else:
        v, eta = np.zeros(mesh.n_edges), np.zeros(mesh.n_cells)
        dv, de = np.zeros((mesh.n_edges, 3)), np.zeros((mesh.n_cells, 3))
    dt = prm["dt"]
    for it in range(int(case["nsteps"])):
        dv[:, 0] = s.updateMomentum(v, eta)
        drag = o.dragForcing(v, s.forcing(dt * it + dt))
        v, dv = s.integrateAB3scalar(v, dv, it)
        v = v + dt * drag
        de[:, 0] = s.updateEta(v)
        eta, de = s.integrateAB3scalar(eta, de, it)
    assert np.array_equal(v, case["final_v"]) and np.array_equal(eta, case["final_eta"])
    assert np.array_equal(dv, case["final_dvdt"]) and np.array_equal(de, case["final_detadt"])
    _, avg = s.updateEnergy(s.interpolateVelocity(v), mesh.tables["face_area"])
    assert abs(avg - case["dump_dissipation_avg"][-1]) <= 1e-12 * abs(case["dump_dissipation_avg"][-1])


def test_operator_call_errors(odis, tmp_path):
    case = load_case("l3_advection_shipped")
    from conftest import nonlinear_tables
    mesh, s, o, prm = pair_for_case(odis, tmp_path, case)
    with pytest.raises(ValueError):
        s.updateEta(np.zeros(3))
    with pytest.raises(odis.OdisError):
        s.integrateAB3scalar(np.zeros(mesh.n_edges + 1), np.zeros((mesh.n_edges + 1, 3)), 0)     # longer than any field
    s.enable_advection(nonlinear_tables(case))
    with pytest.raises(odis.OdisError) as e:
        s.updateMomentum(np.zeros(mesh.n_edges), np.zeros(mesh.n_cells))                           # linear branch only
    assert e.value.code == -6
