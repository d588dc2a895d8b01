"""Overlapped output (odis_run_options.overlap_output / `ODIS --overlap-output`; odis_snapshot_begin / _wait): the next output
interval is computed while the previous dump is copied out and written. It must leave the same output as the synchronous
path — every data.h5 dataset, the progress lines and the restart files, bit for bit — and the snapshot calls must return what odis_get_field returns."""
import filecmp
import os

import numpy as np
import pytest

from conftest import load_case, make_run_dir
from h5lite_reader import read_h5

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["l3_ecc_full_orbit", "l3_shipped_verbatim", "l4_ecc_enceladus"])
def test_overlapped_run_writes_the_same_files(odis, tmp_path, name):
    case = load_case(name)
    a, b = make_run_dir(tmp_path / "sync", case), make_run_dir(tmp_path / "overlap", case)
    ra, rb = odis.run(a), odis.run(b, overlap_output=True)
    assert ra["steps"] == rb["steps"] and ra["dumps"] == rb["dumps"] and ra["last_dissipation_avg"] == rb["last_dissipation_avg"]
    ha, hb = read_h5(os.path.join(a, "DATA", "data.h5")), read_h5(os.path.join(b, "DATA", "data.h5"))
    assert sorted(ha) == sorted(hb) and len(ha) >= 5
    for name in ha:                                            # every dataset, every row, to the bit (object headers carry a timestamp)
        assert ha[name].dtype == hb[name].dtype and np.array_equal(ha[name], hb[name]), name
    lines = lambda d: [l for l in open(os.path.join(d, "DATA", "OUTPUT.txt")) if l.startswith("DUMPING DATA AT")]
    assert lines(a) == lines(b) and len(lines(a)) == ra["dumps"]
    for f in ("vel_init.txt", "pres_init.txt"):
        assert filecmp.cmp(os.path.join(a, "InitialConditions", f), os.path.join(b, "InitialConditions", f), shallow=False)


def test_snapshots_equal_field_reads_while_stepping_continues(odis):
    pos, fr, cen = odis.generate_grid(5)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=30.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=5, friction=0, surface=0, init_load=0, reorder=1)
    s, ref = odis.Solver(mesh, prm), odis.Solver(mesh, prm)
    every = s.SNAP_ETA | s.SNAP_VELOCITY_EN | s.SNAP_DISSIPATION | s.SNAP_VELOCITY
    expected = []
    for k in range(4):                                        # reference: synchronous reads every 30 steps
        ref.step(30)
        expected.append({"eta": ref.field(odis.FIELD_ETA), "velocity_en": ref.field(odis.FIELD_VELOCITY_EN),
                         "dissipation": ref.field(odis.FIELD_DISSIPATION), "velocity": ref.field(odis.FIELD_VELOCITY),
                         "dissipation_avg": ref.dissipation_avg(), "iter": 30 * (k + 1)})
    s.step(30)
    for k in range(4):
        s.snapshot_begin(k & 1, every)
        s.step(30)                                            # runs while the snapshot is copied out
        got = s.snapshot_wait(k & 1)
        for key, val in expected[k].items():
            assert np.array_equal(got[key], val), (k, key)
    with pytest.raises(odis.OdisError):
        s.snapshot_wait(2)
    with pytest.raises(odis.OdisError):
        odis.Solver(mesh, prm).snapshot_wait(0)               # nothing begun on this slot


def test_staged_states_and_snapshots_pipeline_equals_the_synchronous_loop(odis):
    """The input side of the pipeline (odis_stage_state / odis_commit_state) together with the snapshots: interval k+1's state travels to
    the device while interval k is stepping, interval k's fields travel back while interval k+1 is stepping. Five intervals with five
    different states; every result must equal the synchronous set_state / step / field loop to the bit, and the rules of the calls hold."""
    pos, fr, cen = odis.generate_grid(5)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=30.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=5, friction=0, surface=0, init_load=0, reorder=1)
    rng = np.random.default_rng(21)
    F, N = mesh.n_edges, mesh.n_cells
    states = [(rng.uniform(-1, 1, F) * 1e-2, rng.uniform(-1, 1, N), rng.uniform(-1, 1, 3 * F) * 1e-6, rng.uniform(-1, 1, 3 * N) * 1e-4)
              for _ in range(5)]
    states[3] = (states[3][0], states[3][1], None, None)       # arrays that are not given are zero, as in set_state
    S = 20
    ref, expected = odis.Solver(mesh, prm), []
    for k, st in enumerate(states):
        ref.set_state(*st, iter=7 * k)
        ref.step(S)
        expected.append((ref.field(odis.FIELD_ETA), ref.field(odis.FIELD_VELOCITY), ref.dissipation_avg()))
    s = odis.Solver(mesh, prm)
    with pytest.raises(odis.OdisError):
        s.commit_state(0)                                      # nothing staged
    fields = s.SNAP_ETA | s.SNAP_VELOCITY
    s.stage_state(*states[0])
    with pytest.raises(odis.OdisError):
        s.stage_state(*states[1])                              # one staged state at a time
    got = []
    for k in range(len(states)):
        s.commit_state(iter=7 * k)
        if k + 1 < len(states):
            s.stage_state(*states[k + 1])                      # copies while interval k runs
        s.step(S)
        s.snapshot_begin(k & 1, fields)
        if k > 0:
            got.append(s.snapshot_wait((k - 1) & 1))
    got.append(s.snapshot_wait((len(states) - 1) & 1))
    for k, (eta, v, avg) in enumerate(expected):
        assert np.array_equal(got[k]["eta"], eta) and np.array_equal(got[k]["velocity"], v), k
        assert got[k]["dissipation_avg"] == avg and got[k]["iter"] == 7 * k + S
    # and with the self-gravity term (its first potential is part of the commit)
    factor = np.array([0.0, 0.0, 0.4])
    a, b = odis.Solver(mesh, prm), odis.Solver(mesh, prm)
    for x in (a, b):
        x.enable_self_gravity(2, factor)
    a.set_state(*states[1], iter=3)
    b.stage_state(*states[1])
    b.commit_state(iter=3)
    a.step(S); b.step(S)
    assert np.array_equal(a.field(odis.FIELD_ETA), b.field(odis.FIELD_ETA)) and np.array_equal(a.field(odis.FIELD_POTENTIAL), b.field(odis.FIELD_POTENTIAL))
