"""Batched parameter-sweep ensembles (BASELINE config 5): every member must be bit-identical to a run of its own —
against the CPU oracle on a small grid, and against the single-run CUDA solver on a larger one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = dict(g=0.113, h=38e3, alpha=1e-7, dt=30.0, radius=252.1e3, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.002,
            shell_thickness=0.0, semimajor_axis=0.0, friction=0, surface=0, init_load=0, reorder=1)


def members(n, potential, friction=0, init_load=0):
    hs = np.logspace(np.log10(5e3), np.log10(60e3), n)
    al = np.logspace(-9, -6, n)[::-1]
    return [dict(BASE, potential=potential, friction=friction, init_load=init_load, h=float(hs[m]), alpha=float(al[m]),
                 love_reduct=1.0 - 0.02 * m, g=0.113 * (1 + 0.01 * m), ecc=0.0047 * (1 + 0.1 * m), obl=0.002 * (1 + m))
            for m in range(n)]


@pytest.mark.parametrize("potential,n_members", [(5, 5), (9, 4), (8, 3), (1, 2), (16, 1)])
def test_members_match_oracle(odis, potential, n_members):
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(4)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, BASE["radius"])
    plist = members(n_members, potential, init_load=1)
    rng = np.random.default_rng(11)
    ens = odis.Ensemble(mesh, plist)
    states = []
    for m in range(n_members):
        st = (rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells), rng.uniform(-1, 1, (mesh.n_edges, 3)) * 1e-6,
              rng.uniform(-1, 1, (mesh.n_cells, 3)) * 1e-4)
        states.append(st)
        ens.set_state(m, *st, iter=3)
    for m in range(n_members):
        assert np.array_equal(ens.field(m, odis.FIELD_DVDT), states[m][2]) and np.array_equal(ens.field(m, odis.FIELD_DETADT), states[m][3])
    ens.step(17); ens.step(8)
    keys = ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl", "shell_thickness", "potential", "friction", "surface", "init_load")
    for m in range(n_members):
        o = LteOracle(mesh.tables, {k: plist[m][k] for k in keys})
        o.set_state(*states[m], iter=3)
        e0 = o.dissipation_avg()
        so = o.step(25)
        for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT):
            assert np.array_equal(ens.field(m, fid), o.field(fid)), (m, fid)
        assert np.allclose(ens.dissipation_series(m), np.concatenate([[e0], so]), rtol=1e-12, atol=0.0), m


def test_members_match_single_solver_startup(odis):
    """Zero initial state, AB3 start-up steps included, odd member count (one pad slot), 40,962 cells."""
    pos, fr, cen = odis.generate_grid(7)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, BASE["radius"])
    plist = members(3, 5, friction=1)
    ens = odis.Ensemble(mesh, plist)
    ens.step(30)
    assert ens.info()["n_members"] == 3 and ens.iter == 30
    for m in (0, 2):
        s = odis.Solver(mesh, plist[m])
        s.step(30)
        assert np.array_equal(ens.field(m, odis.FIELD_VELOCITY), s.field(odis.FIELD_VELOCITY))
        assert np.array_equal(ens.field(m, odis.FIELD_ETA), s.field(odis.FIELD_ETA))
        assert np.allclose(ens.dissipation_series(m), s.dissipation_series(), rtol=1e-12, atol=0.0)
        s.close()


def test_ensemble_rejects_mixed_time_steps(odis):
    pos, fr, cen = odis.generate_grid(3)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, BASE["radius"])
    plist = members(2, 5)
    plist[1]["dt"] = 31.0
    with pytest.raises(odis.OdisError) as e:
        odis.Ensemble(mesh, plist)
    assert e.value.code == -1


@pytest.mark.parametrize("level,n_members,l_max", [(4, 5, 2), (5, 35, 4), (6, 12, 8), (5, 3, 10)])
def test_ensemble_self_gravity_matches_oracle(odis, level, n_members, l_max):
    """Self-gravity term of all members as FP64 tensor-core GEMMs (analysis: basis x member-innermost eta; synthesis:
    basis^T x coefficients): every member against the CPU oracle with the term on, within 1e-10 (DMMA accumulates with
    FMAs in a different order than the oracle's serial long-double sums)."""
    from oracle.lte_oracle import LteOracle
    from oracle import sh_oracle as so
    pos, fr, cen = odis.generate_grid(level)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, BASE["radius"])
    plist = members(n_members, 5, init_load=1)
    for p in plist:
        p["dt"] = 15.0
    factor = 0.6 / (1.0 + 0.3 * np.arange(l_max + 1))
    rng = np.random.default_rng(5)
    ens = odis.Ensemble(mesh, plist)
    ens.enable_self_gravity(l_max, factor)
    states = []
    for m in range(n_members):
        st = (rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells), rng.uniform(-1, 1, (mesh.n_edges, 3)) * 1e-6,
              rng.uniform(-1, 1, (mesh.n_cells, 3)) * 1e-4)
        states.append(st)
        ens.set_state(m, *st, iter=3)
    Y = so.basis(pos, l_max)
    T = so.apply_operator(Y, factor)
    check = sorted(set([0, n_members // 2, n_members - 1]))
    for m in check:
        assert np.abs(ens.sh_coefficients(m) - so.lsq_coefficients(Y, states[m][1])).max() <= 1e-11
    ens.step(20)
    keys = ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl", "shell_thickness", "potential", "friction", "surface", "init_load")
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    for m in check:
        o = LteOracle(mesh.tables, {k: plist[m][k] for k in keys})
        o.set_self_gravity(Y, T)
        o.set_state(*states[m], iter=3)
        o.step(20)
        for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA):
            assert rel(ens.field(m, fid), o.field(fid)) <= 1e-10, (m, fid)
        plain = LteOracle(mesh.tables, {k: plist[m][k] for k in keys})
        plain.set_state(*states[m], iter=3)
        plain.step(20)
        assert rel(plain.field(1), o.field(1)) > 1e-7          # the term matters in this set-up


def test_ensemble_self_gravity_matches_single_solver(odis):
    """40,962 cells, 32 members (one full member block): members against single runs with the matrix-free kernels."""
    pos, fr, cen = odis.generate_grid(7)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, BASE["radius"])
    plist = members(32, 5)
    for p in plist:
        p["dt"] = 10.0
    factor = np.array([0.0, 0.0, 0.5, 0.4, 0.3])
    ens = odis.Ensemble(mesh, plist)
    ens.enable_self_gravity(4, factor)
    ens.step(30)
    for m in (0, 17, 31):
        s = odis.Solver(mesh, plist[m])
        s.enable_self_gravity(4, factor)
        s.step(30)
        for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA):
            a, b = ens.field(m, fid), s.field(fid)
            assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max(), (m, fid)
    with pytest.raises(odis.OdisError):
        ens.enable_self_gravity(4, factor)
    with pytest.raises(odis.OdisError):
        odis.Ensemble(mesh, plist[:2]).enable_self_gravity(12, np.zeros(13))
