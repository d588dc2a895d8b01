"""The default self-gravity step (degrees 2..4, matrix-free basis): edge update, then the staged cell update with the harmonic
analysis b = Y eta^{n+1} accumulated per thread on the way (cell_step_pipe_kernel, odis_kernels_pipe.cu) and — on hardware — a grid-wide
barrier behind which every CTA solves and adds the term to the potential of its own tiles (2 launches per step; cooperative launch). With
ODIS_B200_MERGED_SYNTH=0, and under the host emulation (where the CTAs of a launch run one after another), solve + synthesis are a third
launch (sh_bsolve_synthesis_mf_kernel, odis_sh.cu); both forms give the same bits — against the CPU oracle (1e-10, BASELINE.json's bar; the term has no reference arithmetic
to follow) and against the baseline selection (kernel_select=1: direct-load kernels, separate analysis / reduce-solve /
synthesis launches; same sums, different association)."""
import os

import numpy as np
import pytest

from test_self_gravity_gpu import rel_err, setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level,l_max", [(4, 2), (5, 2), (6, 2), (5, 3), (6, 4)])
def test_three_launch_step_matches_oracle_and_baseline_kernels(odis, level, l_max):
    mesh, pos, prm, factor, state, _, o, Y = setup(odis, level, l_max)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    s_base = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=1))
    for solver in (s, s_base):
        solver.enable_self_gravity(l_max, factor)
        solver.set_state(*state, iter=5)
    o.set_state(*state, iter=5)
    n = 60
    series_o = o.step(n)
    l0, b0 = s.launches, s_base.launches
    s.step(25); s.step(n - 25)                                  # graph replay + single launches
    assert s.launches - l0 in (2 * n, 3 * n)                   # merged kernel / separate solve + synthesis launch
    if os.environ.get("ODIS_TEST_EXPECT_MERGED"):              # set by the emulated run that must exercise the grid-barrier kernel
        assert s.launches - l0 == 2 * n
    s_base.step(n)
    assert s_base.launches - b0 == 5 * n
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT):
        assert rel_err(s.field(fid), o.field(fid)) <= 1e-10, fid
        assert rel_err(s.field(fid), s_base.field(fid)) <= 1e-11, fid
    # the potential held for the NEXT step (tide + the term of the newest eta; the oracle keeps the previous step's)
    assert rel_err(s.field(odis.FIELD_POTENTIAL), s_base.field(odis.FIELD_POTENTIAL)) <= 1e-11
    assert np.allclose(s.dissipation_series()[1:], series_o, rtol=1e-10, atol=0.0)
    assert np.abs(s.sh_coefficients() - s_base.sh_coefficients()).max() <= 1e-11 * max(1.0, np.abs(s_base.sh_coefficients()).max())
    # repeatable to the bit: the sums do not depend on the order in which CTAs run or finish
    again = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    again.enable_self_gravity(l_max, factor)
    again.set_state(*state, iter=5)
    again.step(n)
    assert np.array_equal(again.field(odis.FIELD_ETA), s.field(odis.FIELD_ETA))


def test_high_degree_keeps_the_separate_launches(odis):
    """Degree > 4: the sums do not fit a thread's registers; the staged cell update runs without them and the analysis / reduce-solve /
    synthesis launches follow (5 launches per step)."""
    mesh, pos, prm, factor, state, _, o, Y = setup(odis, 4, 8)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0))
    s.enable_self_gravity(8, factor)
    s.set_state(*state, iter=5)
    o.set_state(*state, iter=5)
    o.step(30)
    l0 = s.launches
    s.step(30)
    assert s.launches - l0 == 5 * 30
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA):
        assert rel_err(s.field(fid), o.field(fid)) <= 1e-10, fid


@pytest.mark.parametrize("l_max", [2, 4])
@pytest.mark.parametrize("world", [2, 4])
def test_three_launch_step_on_a_partitioned_grid(odis, world, l_max):
    """Partitioned solvers: edge update (halo push), cell update + analysis + publish of the rank's harmonic sums, all-reduce through peer
    memory + solve + synthesis — 3 launches per step. Fields against the single-device run (1e-10; the sums group differently),
    identical coefficients on every rank."""
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
    parts = [odis.Solver(mesh, prm, device=k, rank=k, world=world) for k in range(world)]
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
    assert all(p.launches - a in (2 * 60, 3 * 60) for p, a in zip(parts, l0))
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_POTENTIAL):
        total = sum(p.field(fid) for p in parts)
        assert rel_err(total, ref.field(fid)) <= 1e-10, fid
    coeffs = [p.sh_coefficients() for p in parts]
    for c in coeffs[1:]:
        assert np.array_equal(c, coeffs[0])
    assert np.abs(coeffs[0] - ref.sh_coefficients()).max() <= 1e-11 * np.abs(ref.sh_coefficients()).max()
    for p in parts:
        p.synchronize()
