"""Parity at BASELINE.json's own sizes ("Next round" item 1).

(a) 655,362 cells (BASELINE 'L8') and 163,842 cells ('L7'): the default kernels against the CPU oracle (oracle/lte_oracle.c, pinned
    bit-for-bit to the reference's own solver) on a seeded random state, 60 steps: v, eta and both AB3 histories must be BIT-IDENTICAL.
    At these sizes every CTA of the staged edge kernel owns many tiles (15,360 tiles over 296 CTAs at 655,362 cells), so the shared-memory
    stages are recycled and the `empty`-mbarrier / parity-flip path of the pipeline is compared with the oracle, which the 10,242-cell
    fixtures cannot do (one tile per CTA). Staged (kernel_select 0) against direct-load (1) kernels on the same sizes as well.
(b) the same with the spherical-harmonic self-gravity term: within 1e-10 relative of the oracle (the term has no reference arithmetic
    to follow: tree sums on the device, FMA in the recurrences — `parity unpinned`).
(c) BASELINE config 0 at its real size: /root/reference/input.in UNCHANGED (advection true -> nonlinear branch, OBLIQ_WEST, Earth-like)
    on the shipped grid_l6.txt (10,242 cells) against FP64 checkpoints of the reference's own run at steps 10 / 100 / 1000 / 2900
    (tests/golden/case_l6_shipped_verbatim.npz, written by tests/golden/make_golden.py).

Tolerances: BASELINE.json asks for <= 1e-10 relative on eta and v after N steps and <= 1e-8 on the dissipation; the kernels follow the
reference's FP64 operation order with FMA contraction off, so bit equality is asserted (and 1e-10 beside it, so a failure says how far)."""
import os

import numpy as np
import pytest

from conftest import case_params, load_case, make_run_dir

pytestmark = pytest.mark.gpu

FIELD_RTOL = 1e-10
DISS_RTOL = 1e-12

ENC = dict(g=0.113, h=38e3, alpha=1e-7, radius=252.1e3, omega=5.307e-5, love_reduct=0.9, ecc=0.0047, obl=0.0,
           shell_thickness=0.0, potential=5, friction=0, surface=0)
_mesh_cache = {}


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def mesh_for(odis, level):
    if level not in _mesh_cache:
        pos, fr, cen = odis.generate_grid(level)
        _mesh_cache[level] = (odis.Mesh.from_arrays(pos, fr, cen, ENC["radius"]), pos)
    return _mesh_cache[level]


def random_state(mesh, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells),
            rng.uniform(-1, 1, (mesh.n_edges, 3)) * 1e-6, rng.uniform(-1, 1, (mesh.n_cells, 3)) * 1e-4)


def params_for(mesh, **over):
    dmin = float(mesh.tables["face_node_dist"].min())
    return dict(ENC, dt=0.2 * dmin / np.sqrt(ENC["g"] * ENC["h"]), init_load=0, **over)


@pytest.mark.parametrize("level,cells", [(9, 655362), (8, 163842)])
def test_default_kernels_match_oracle_at_baseline_size(odis, level, cells):
    """60 steps from a seeded random state; the first two are the AB3 start-up steps (iter 0, 1), the rest the 3-level formula, so the
    run goes through the one-by-one launches AND the captured graphs (12 steps each) of the default selection."""
    from oracle.lte_oracle import LteOracle
    mesh, _ = mesh_for(odis, level)
    assert mesh.n_cells == cells
    prm = params_for(mesh)
    v0, e0, dv, de = random_state(mesh, 20 + level)
    o = LteOracle(mesh.tables, prm)
    o.set_state(v0, e0, dv, de, iter=0)
    e_init = o.dissipation_avg()
    so = o.step(60)
    fields = {}
    for sel in (0, 1):                       # staged (default) and direct-load edge kernel
        s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=sel))
        s.set_state(v0, e0, dv, de, iter=0)
        s.step(17); s.step(43)               # 2 start-up + 15 single launches, then 36 graph-replayed + 7 single
        fields[sel] = [s.field(f) for f in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT)]
        series = s.dissipation_series()
        s.close()
        for k, fid in enumerate((0, 1, 2, 3)):
            ref = o.field(fid)
            assert rel_err(fields[sel][k], ref) <= FIELD_RTOL, (sel, fid, rel_err(fields[sel][k], ref))
            assert np.array_equal(fields[sel][k], ref), (sel, fid, rel_err(fields[sel][k], ref))
        assert np.allclose(series, np.concatenate([[e_init], so]), rtol=DISS_RTOL, atol=0.0)
    for a, b in zip(fields[0], fields[1]):
        assert np.array_equal(a, b)


def test_self_gravity_matches_oracle_at_baseline_size(odis):
    """655,362 cells with the degree-2 self-gravity / shell-pressure term (the bench's headline workload): <= 1e-10 of the oracle."""
    from oracle.lte_oracle import LteOracle
    from oracle import sh_oracle
    mesh, pos = mesh_for(odis, 9)
    prm = params_for(mesh, surface=2, shell_thickness=23e3)
    factor = np.array([0.0, 0.0, 1.0 - 2.970754525850653494e+01])
    v0, e0, dv, de = random_state(mesh, 77)
    o = LteOracle(mesh.tables, prm)
    Y = sh_oracle.basis(pos, 2)
    o.set_self_gravity(Y, sh_oracle.apply_operator(Y, factor))
    o.set_state(v0, e0, dv, de, iter=0)
    o.step(40)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    s.enable_self_gravity(2, factor)
    s.set_state(v0, e0, dv, de, iter=0)
    s.step(40)
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA):
        assert rel_err(s.field(fid), o.field(fid)) <= FIELD_RTOL, (fid, rel_err(s.field(fid), o.field(fid)))
    s.close()


# ---- (c) the shipped input.in on the shipped L6 grid -----------------------------------------------------------------------------------
def shipped_l6(odis, tmp_path):
    case = load_case("l6_shipped_verbatim")
    grid = load_case(str(case["grid_case"]))
    merged = dict(grid)
    merged.update({k: case[k] for k in case})
    d = make_run_dir(tmp_path, merged)
    return case, merged, d


def test_shipped_input_on_shipped_l6_grid_matches_reference_checkpoints(odis, tmp_path):
    case, merged, d = shipped_l6(odis, tmp_path)
    text = str(case["input_in"])
    # the reference's file as shipped (only the end time differs: 1 orbit instead of 150)
    assert "advection;                     true;" in text and "geodesic grid level; \t \t \t6;" in text and "OBLIQ_WEST" in text
    mesh = odis.Mesh.from_file(os.path.join(d, "input_files", "grid_l6.txt"), float(case["scalar_radius"][0]))
    assert mesh.n_cells == 10242 and int(case["scalar_advection"][0]) == 1
    nl = odis.nonlinear_tables(mesh, 0.5)
    # the operators only this branch reads are the reference's, entry for entry (digests of the reference's own tables)
    import hashlib
    for key in ("operatorCurl", "operatorRBFinterp", "operatorDirectionalSecondDeriv"):
        for part in ("indptr", "indices", "data"):
            got = hashlib.sha256(np.ascontiguousarray(nl[f"{key}.{part}"]).tobytes()).hexdigest()
            assert got == str(case[f"sha256_{key}.{part}"]), (key, part)
    prm = case_params(merged)
    s = odis.Solver(mesh, dict(prm, reorder=1))
    s.enable_advection(nl)
    done = 0
    for n in (int(x) for x in case["checkpoints"]):
        s.step(n - done)
        done = n
        v, eta = s.field(odis.FIELD_VELOCITY), s.field(odis.FIELD_ETA)
        rv, re = case[f"step{n}_v"], case[f"step{n}_eta"]
        if np.isfinite(rv).all() and np.isfinite(re).all():
            # tolerance of SURVEY §8d-1 / BASELINE.json: 1e-10 relative; the kernels are built for bit equality
            assert rel_err(v, rv) <= FIELD_RTOL and rel_err(eta, re) <= FIELD_RTOL, (n, rel_err(v, rv), rel_err(eta, re))
            assert np.array_equal(v, rv) and np.array_equal(eta, re), (n, rel_err(v, rv), rel_err(eta, re))
            if f"step{n}_dvdt" in case:
                assert np.array_equal(s.field(odis.FIELD_DVDT), case[f"step{n}_dvdt"])
                assert np.array_equal(s.field(odis.FIELD_DETADT), case[f"step{n}_detadt"])
        else:
            # the reference's own run of its shipped configuration diverges on this grid (NaN from some step in 1,160..1,450 on): the
            # same arithmetic diverges the same way — non-finite in exactly the entries where the reference is
            assert n == 2900
            assert np.array_equal(np.isfinite(v), np.isfinite(rv)) and np.array_equal(np.isfinite(eta), np.isfinite(re))
    # dissipation at the reference's dump slices while it is finite (steps 0, 290, ..., 1160)
    series = s.dissipation_series()
    ref = case["dump_dissipation_avg"]
    ok = np.isfinite(ref)
    assert ok[:5].all() and not ok[5:].any()
    ours = series[(case["dump_slices"] - 1) * 290]
    assert np.allclose(ours[ok], ref[ok], rtol=DISS_RTOL, atol=0.0)
    s.close()


def test_shipped_input_on_shipped_l6_grid_through_odis_run(odis, tmp_path):
    """The same through the drop-in boundary (`./ODIS` in the run directory = odis_run): data.h5 rows and the progress lines of the part
    of the orbit in which the reference's run is finite."""
    from h5lite_reader import read_h5
    case, merged, d = shipped_l6(odis, tmp_path)
    res = odis.run(d, max_steps=1160)                       # 4 output intervals: dumps at steps 0, 290, 580, 870, 1160
    assert res["steps"] == 1160 and res["steps_per_period"] == 2900 and res["dt"] == float(case["scalar_timeStep"][0])
    h5 = read_h5(os.path.join(d, "DATA", "data.h5"))
    assert "x velocity" in h5 and h5["displacement"].shape == (11, 10242)
    ref_eta = case["dump_displacement_first5"].astype(np.float32)
    assert np.array_equal(h5["displacement"][:5], ref_eta)
    ref_d = case["dump_dissipation_avg"][:5].astype(np.float32)
    got_d = h5["dissipation avg output"].ravel()[:5]
    assert np.abs(got_d - ref_d).max() <= 2e-7 * np.abs(ref_d).max()
    lines = lambda t: [l for l in t.splitlines() if l.startswith("DUMPING DATA AT")]
    assert lines(open(os.path.join(d, "DATA", "OUTPUT.txt")).read())[:5] == lines(str(case["output_txt"]))[:5]
