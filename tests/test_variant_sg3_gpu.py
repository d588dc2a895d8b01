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
