"""Spherical-harmonic self-gravity / shell-pressure term on the GPU (odis_enable_self_gravity) against the CPU oracle
(oracle/lte_oracle.c with the term of oracle/sh_oracle.py switched on).

Tolerance: the device sums the dense products in a tree, the oracle serially in long double, so fields agree to
rounding, not bit for bit. BASELINE.json's bar (eta, v within 1e-10 relative after N steps) is asserted; the
measured differences are ~1e-14. The term itself has no reference run to pin against (dead code at HEAD)."""
import numpy as np
import pytest

from oracle import sh_oracle as so

pytestmark = pytest.mark.gpu

FIELD_RTOL = 1e-10


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def setup(odis, level, l_max, seed=11):
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(level)
    r = 252.1e3 - 23e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=20.0, radius=r, omega=5.307e-5, love_reduct=0.95, ecc=0.0047, obl=0.002,
               shell_thickness=23e3, potential=5, friction=0, surface=2, init_load=1)
    factor = 0.6 / (1.0 + 0.3 * np.arange(l_max + 1))           # stands in for 1 - beta_l
    rng = np.random.default_rng(seed)
    state = (rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells),
             rng.uniform(-1, 1, (mesh.n_edges, 3)) * 1e-6, rng.uniform(-1, 1, (mesh.n_cells, 3)) * 1e-4)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    o = LteOracle(mesh.tables, prm)
    Y = so.basis(pos, l_max)
    o.set_self_gravity(Y, so.apply_operator(Y, factor))
    return mesh, pos, prm, factor, state, s, o, Y


@pytest.mark.parametrize("stored", [False, True])           # matrix-free (default) and stored-basis GEMV kernels
@pytest.mark.parametrize("level,l_max", [(4, 2), (5, 2), (5, 8), (6, 4), (5, 12)])     # l_max 12: 169 rows, the separate solve launch
def test_time_steps_match_oracle(odis, level, l_max, stored):
    mesh, pos, prm, factor, state, s, o, Y = setup(odis, level, l_max)
    s.enable_self_gravity(l_max, factor, stored_basis=stored)
    s.set_state(*state, iter=5)
    o.set_state(*state, iter=5)
    # potential of the pending step = tide + g * sum factor_l c_lm Y_lm of the eta just loaded
    tide_only = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    tide_only.set_state(*state, iter=5)
    extra = s.field(odis.FIELD_POTENTIAL) - tide_only.field(odis.FIELD_POTENTIAL)
    ref_extra = so.self_gravity_potential(Y, factor, prm["g"], state[1])
    assert np.abs(extra - ref_extra).max() <= 1e-11 * np.abs(ref_extra).max()
    assert np.abs(s.sh_coefficients() - so.lsq_coefficients(Y, state[1])).max() <= 1e-12
    n = 60
    series_o = o.step(n)
    s.step(25); s.step(n - 25)                                  # graph replay + single launches
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT):
        assert rel_err(s.field(fid), o.field(fid)) <= FIELD_RTOL, fid
    assert np.allclose(s.dissipation_series()[1:], series_o, rtol=1e-10, atol=0.0)
    # the term matters in this set-up: switching it off changes eta far beyond the tolerance
    tide_only.step(n)
    assert rel_err(tide_only.field(odis.FIELD_ETA), o.field(1)) > 1e-6


@pytest.mark.parametrize("stored", [False, True])
def test_band_limited_eta_is_recovered_on_device(odis, stored):
    """Size-independent property at a larger grid (163,842 cells): eta synthesised from known coefficients ->
    the device analysis returns them, and the potential gets exactly g * factor_l * c_lm Y_lm."""
    l_max = 6
    pos, fr, cen = odis.generate_grid(8)
    r = 1.0e6
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=1.3, h=1.0e3, alpha=1e-6, dt=5.0, radius=r, omega=2e-5, love_reduct=1.0, ecc=0.0, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=16, friction=0, surface=0, init_load=0, reorder=1)
    s = odis.Solver(mesh, prm)
    factor = np.linspace(1.0, 0.2, l_max + 1)
    s.enable_self_gravity(l_max, factor, stored_basis=stored)
    Y = odis.sh_basis(pos, l_max)
    rng = np.random.default_rng(2)
    c = rng.uniform(-1, 1, Y.shape[0])
    s.set_state(eta=Y.T @ c)
    assert np.abs(s.sh_coefficients() - c).max() <= 1e-11
    f = factor[so.row_degree(l_max)].copy(); f[:4] = 0.0
    u = s.field(odis.FIELD_POTENTIAL)                          # potential "NONE": the self-gravity term alone
    ref = prm["g"] * (Y.T @ (f * c))
    assert np.abs(u - ref).max() <= 1e-11 * np.abs(ref).max()


def test_high_degree_matrix_free(odis):
    """l_max = 30 (961 rows, the dynamic shared-memory path) on a 40,962-cell grid: band-limited recovery."""
    l_max = 30
    pos, fr, cen = odis.generate_grid(7)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
    prm = dict(g=1.3, h=1.0e3, alpha=1e-6, dt=5.0, radius=1.0e6, omega=2e-5, love_reduct=1.0, ecc=0.0, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=16, friction=0, surface=0, init_load=0, reorder=1)
    s = odis.Solver(mesh, prm)
    factor = 1.0 / (1.0 + np.arange(l_max + 1))
    s.enable_self_gravity(l_max, factor)
    Y = odis.sh_basis(pos, l_max)
    c = np.random.default_rng(4).uniform(-1, 1, Y.shape[0])
    s.set_state(eta=Y.T @ c)
    assert np.abs(s.sh_coefficients() - c).max() <= 1e-10
    f = factor[so.row_degree(l_max)].copy(); f[:4] = 0.0
    ref = prm["g"] * (Y.T @ (f * c))
    assert np.abs(s.field(odis.FIELD_POTENTIAL) - ref).max() <= 1e-10 * np.abs(ref).max()


def test_enable_errors(odis):
    pos, fr, cen = odis.generate_grid(3)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
    prm = dict(g=1.3, h=1.0e3, alpha=1e-6, dt=5.0, radius=1.0e6, omega=2e-5, love_reduct=1.0, ecc=0.0, obl=0.0,
               shell_thickness=0.0, semimajor_axis=0.0, potential=5, friction=0, surface=0, init_load=0, reorder=1)
    s = odis.Solver(mesh, prm)
    with pytest.raises(odis.OdisError):
        s.enable_self_gravity(1, [0.0, 0.0])
    with pytest.raises(odis.OdisError):
        s.enable_self_gravity(40, np.zeros(41))
    s.enable_self_gravity(2, [0.0, 0.0, 0.5])
    with pytest.raises(odis.OdisError):
        s.enable_self_gravity(2, [0.0, 0.0, 0.5])               # already on
