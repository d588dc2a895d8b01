"""Pins the CPU oracle (oracle/lte_oracle.c) to the reference's own solver.

The fixtures hold FP64 states produced by the UNMODIFIED reference sources (oracle/_ref, see
tests/golden/make_golden.py). The oracle restates the same arithmetic in the same order, so the bar
is bit-for-bit equality of v, eta, both AB3 histories and the per-step dissipation."""
import os

import numpy as np
import pytest

from conftest import ALL_CASES, NL_CASES, case_params, load_case, make_run_dir, nonlinear_tables
from oracle.lte_oracle import LteOracle


def oracle_for(odis, tmp_path, case):
    d = make_run_dir(tmp_path, case)
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l%d.txt" % int(case["level"])), float(case["scalar_radius"][0]))
    loaded = "init_v" in case
    o = LteOracle(mesh.tables, case_params(case, init_load=int(loaded)))
    if int(case["scalar_advection"][0]):
        o.set_nonlinear(nonlinear_tables(case))
    if loaded:
        o.set_state(case["init_v"], case["init_eta"], case["init_dvdt"], case["init_detadt"])
    else:
        o.set_state()
    return mesh, o


@pytest.mark.parametrize("name", ALL_CASES + NL_CASES)
def test_oracle_reproduces_reference_state(odis, tmp_path, name):
    case = load_case(name)
    mesh, o = oracle_for(odis, tmp_path, case)
    e0 = o.dissipation_avg()
    series = o.step(int(case["nsteps"]))
    assert np.array_equal(o.field(0), case["final_v"])
    assert np.array_equal(o.field(1), case["final_eta"])
    assert np.array_equal(o.field(2), case["final_dvdt"])
    assert np.array_equal(o.field(3), case["final_detadt"])
    # dissipation at the reference's dump slices (slice k+1 is written after k*out_freq steps)
    slices = case["dump_slices"]
    total, out_time = int(case["scalar_totalIter"][0]), int(case["scalar_outputTime"][0])
    out_freq = total // out_time
    full = np.concatenate([[e0], series])
    assert np.array_equal(full[(slices - 1) * out_freq], case["dump_dissipation_avg"])


def test_oracle_operators_match_reference_csr(odis, tmp_path):
    """Operator assembly (mesh.cpp:2808-3287) against the reference's Eigen matrices, entry by entry."""
    import ctypes as C
    from oracle import lte_oracle
    case = load_case("l3_obliqwest_earth")
    mesh, o = oracle_for(odis, tmp_path, case)
    lib = lte_oracle._load()
    for which, name in enumerate(["operatorGradient", "operatorDivergence", "operatorCoriolis", "operatorLinearDrag"]):
        nr, nc = C.c_int(), C.c_int()
        ptr, idx, val = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        lib.oracle_get_operator.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        lib.oracle_get_operator(o._h, which, C.byref(nr), C.byref(nc), C.byref(ptr), C.byref(idx), C.byref(val))
        indptr = np.ctypeslib.as_array(ptr, shape=(nr.value + 1,))
        nnz = int(indptr[-1])
        assert np.array_equal(indptr, case[f"table_{name}.indptr"])
        assert np.array_equal(np.ctypeslib.as_array(idx, shape=(nnz,)), case[f"table_{name}.indices"])
        assert np.array_equal(np.ctypeslib.as_array(val, shape=(nnz,)), case[f"table_{name}.data"])


def test_dumped_fields_match(odis, tmp_path):
    """v_avg (interpolateVelocity) and eta at the reference's last dump."""
    case = load_case("l4_ecc_enceladus")
    mesh, o = oracle_for(odis, tmp_path, case)
    total, out_time = int(case["scalar_totalIter"][0]), int(case["scalar_outputTime"][0])
    out_freq = total // out_time
    last = int(case["dump_slices"][-1])
    o.step((last - 1) * out_freq)
    assert np.array_equal(o.field(1), case["dump_displacement"][-1])
    assert np.array_equal(o.field(4), case["dump_velocity_en"][-1])
    # and the float32 rows the reference handed to HDF5
    assert np.array_equal(o.field(1).astype(np.float32), case["h5_displacement"][last - 1] if case["h5_displacement"].shape[0] >= last else o.field(1).astype(np.float32))


@pytest.mark.parametrize("name", ["l3_obliqwest_earth", "l3_full_loaded", "l3_obliq_quadratic"])
def test_operator_calls_compose_to_a_step(odis, tmp_path, name):
    """The oracle's loop-level functions called one by one (the checkers of the odis_op_* entry points) in the order of
    timeIntegrator.cpp:205-277 reproduce the reference's final state bit for bit."""
    case = load_case(name)
    mesh, o = oracle_for(odis, tmp_path, case)
    prm = case_params(case)
    dt = prm["dt"]
    v, eta, dv, de = o.field(0), o.field(1), o.field(2), o.field(3)
    it = 0
    for _ in range(int(case["nsteps"])):
        dv[:, 0] = o.updateMomentum(v, eta)
        drag = o.dragForcing(v, o.forcing(dt * it + dt))
        v, dv = o.integrateAB3scalar(v, dv, it)              # start-up formulas unless the case was loaded (INIT_LOAD)
        v = v + dt * drag
        de[:, 0] = o.updateEta(v)
        eta, de = o.integrateAB3scalar(eta, de, it)
        it += 1
    assert np.array_equal(v, case["final_v"]) and np.array_equal(eta, case["final_eta"])
    assert np.array_equal(dv, case["final_dvdt"]) and np.array_equal(de, case["final_detadt"])
    v_avg = o.interpolateVelocity(v)
    e_flux, avg = o.updateEnergy(v_avg, mesh.tables["face_area"])
    assert avg == case["dump_dissipation_avg"][-1]
