"""Opt-in kernel selection bit 4 (kernel_select=16): self-gravity term in 3 launches per step — harmonic analysis folded into
the cell update (cell_step_sg_kernel), reduce + solve folded into the synthesis (sh_solve_synthesis_mf_kernel) — against the
CPU oracle (1e-10, BASELINE.json's bar) and against the default 5-launch path (same sums, different association)."""
import numpy as np
import pytest

from test_self_gravity_gpu import rel_err, setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("select", [16, 80])                       # 80: the same with the cell kernel's registers capped at 64
@pytest.mark.parametrize("level,l_max", [(4, 2), (5, 2), (6, 2), (5, 3), (6, 4)])
def test_three_launch_variant_matches_oracle_and_default_path(odis, level, l_max, select):
    mesh, pos, prm, factor, state, s_default, o, Y = setup(odis, level, l_max)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=select))
    for solver in (s, s_default):
        solver.enable_self_gravity(l_max, factor)
        solver.set_state(*state, iter=5)
    o.set_state(*state, iter=5)
    n = 60
    series_o = o.step(n)
    l0 = s.launches
    s.step(25); s.step(n - 25)                                  # graph replay + single launches
    assert s.launches - l0 == 3 * n
    s_default.step(n)
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT):
        assert rel_err(s.field(fid), o.field(fid)) <= 1e-10, fid
        assert rel_err(s.field(fid), s_default.field(fid)) <= 1e-11, fid
    # the potential held for the NEXT step (tide + the term of the newest eta; the oracle keeps the previous step's)
    assert rel_err(s.field(odis.FIELD_POTENTIAL), s_default.field(odis.FIELD_POTENTIAL)) <= 1e-11
    assert np.allclose(s.dissipation_series()[1:], series_o, rtol=1e-10, atol=0.0)
    assert np.abs(s.sh_coefficients() - s_default.sh_coefficients()).max() <= 1e-11 * max(1.0, np.abs(s_default.sh_coefficients()).max())
    # repeatable to the bit: the sums do not depend on the order in which CTAs finish
    again = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=select))
    again.enable_self_gravity(l_max, factor)
    again.set_state(*state, iter=5)
    again.step(n)
    assert np.array_equal(again.field(odis.FIELD_ETA), s.field(odis.FIELD_ETA))


def test_three_launch_variant_rejects_what_it_does_not_cover(odis):
    mesh, pos, prm, factor, state, s_default, o, Y = setup(odis, 4, 8)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=16))
    with pytest.raises(odis.OdisError) as e:
        s.enable_self_gravity(8, factor)                         # degree > 4
    assert e.value.code == -6


@pytest.mark.parametrize("l_max", [2, 4])
@pytest.mark.parametrize("world", [2, 4])
def test_three_launch_variant_on_a_partitioned_grid(odis, world, l_max):
    """Partitioned solvers with bit 4: edge update (halo push), cell update + analysis + publish of the rank's harmonic sums
    (cell_step_sgx_kernel), all-reduce + solve + synthesis (sh_allsolve_synthesis_mf_kernel) — 3 launches per step instead of 6.
    Fields against the single-device default path (1e-10; the sums group differently), identical coefficients on every rank."""
    from test_multigpu import _device_count
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=40.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001, shell_thickness=0.0,
               semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
    factor = 0.5 / (1.0 + 0.2 * np.arange(l_max + 1))
    rng = np.random.default_rng(3)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, prm, device=0)
    ref.enable_self_gravity(l_max, factor)
    ref.set_state(v0, e0)
    ref.step(60)
    parts = [odis.Solver(mesh, dict(prm, kernel_select=16), device=k, rank=k, world=world) for k in range(world)]
    blobs = [p.halo_blob() for p in parts]
    for p in parts:
        p.halo_connect(blobs)
    for p in parts:
        p.enable_self_gravity(l_max, factor)
    for p in parts:
        p.set_state(v0, e0)
    l0 = [p.launches for p in parts]
    for n in (25, 35):                                   # graph replay + single launches, every rank the same steps in turn
        for p in parts:
            p.step(n)
    assert all(p.launches - a == 3 * 60 for p, a in zip(parts, l0))
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_POTENTIAL):
        total = sum(p.field(fid) for p in parts)
        assert rel_err(total, ref.field(fid)) <= 1e-10, fid
    coeffs = [p.sh_coefficients() for p in parts]
    for c in coeffs[1:]:
        assert np.array_equal(c, coeffs[0])
    assert np.abs(coeffs[0] - ref.sh_coefficients()).max() <= 1e-11 * np.abs(ref.sh_coefficients()).max()
    for p in parts:
        p.synchronize()
