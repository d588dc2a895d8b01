"""Opt-in L2 prefetch in the per-step cell update (odis_params.reserved[0] bit 9, `kernel_select=512`; with bit 6 also in the
register-capped kernel): one thread per CTA issues cp.async.bulk.prefetch.L2 for the streamed rows of the tile one GPU-full of CTAs
ahead. No arithmetic changes, so the fields must equal the oracle's to the bit; the prefetched ranges must stay inside the arrays
(the emulation touches both ends of every range under the address sanitizer)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRM = dict(g=0.113, h=38e3, alpha=1e-6, dt=30.0, radius=252.1e3, omega=5.307e-5, love_reduct=0.95, ecc=0.0047, obl=0.002,
           shell_thickness=0.0, friction=1, surface=0, init_load=0)


@pytest.mark.parametrize("level", [3, 5, 6])
@pytest.mark.parametrize("potential", [5, 8])
@pytest.mark.parametrize("kernel_select", [512, 512 + 64, 512 + 8])
def test_prefetching_cell_update_matches_oracle(odis, level, potential, kernel_select):
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(level)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    prm = dict(PRM, potential=potential)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, kernel_select=kernel_select))
    o = LteOracle(mesh.tables, prm)
    o.set_state()
    series = o.step(40)
    s.step(40)
    assert np.array_equal(s.field(odis.FIELD_VELOCITY), o.field(0)) and np.array_equal(s.field(odis.FIELD_ETA), o.field(1))
    assert np.array_equal(s.field(odis.FIELD_DETADT), o.field(3))
    assert np.allclose(s.dissipation_series()[1:], series, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("world", [2])
def test_prefetching_cell_update_on_a_partitioned_grid(odis, world):
    from test_multigpu import _device_count
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    prm = dict(PRM, potential=8, reorder=1, semimajor_axis=0.0, friction=0)
    rng = np.random.default_rng(5)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, prm, device=0)
    ref.set_state(v0, e0)
    ref.step(50)
    parts = [odis.Solver(mesh, dict(prm, kernel_select=512), device=k, rank=k, world=world) for k in range(world)]
    blobs = [p.halo_blob() for p in parts]
    for p in parts:
        p.halo_connect(blobs)
    for p in parts:
        p.set_state(v0, e0)
    for n in (20, 30):
        for p in parts:
            p.step(n)
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT):
        assert np.array_equal(sum(p.field(fid) for p in parts), ref.field(fid)), fid


@pytest.mark.parametrize("world", [1, 2])
def test_prefetch_with_the_three_launch_self_gravity_step(odis, world):
    """Bits 4 + 9: the cell update that also carries the harmonic analysis prefetches too (single solver: same bits as without the
    prefetch; partitioned: against the single-device default, 1e-10)."""
    from test_multigpu import _device_count
    if world > 1 and _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    prm = dict(PRM, potential=5, reorder=1, semimajor_axis=0.0, friction=0)
    factor = np.array([0.0, 0.0, 0.4])
    rng = np.random.default_rng(9)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, dict(prm, kernel_select=16 if world == 1 else 0), device=0)
    ref.enable_self_gravity(2, factor)
    ref.set_state(v0, e0)
    ref.step(40)
    parts = [odis.Solver(mesh, dict(prm, kernel_select=16 + 512), device=k, rank=k, world=world) for k in range(world)]
    if world > 1:
        blobs = [p.halo_blob() for p in parts]
        for p in parts:
            p.halo_connect(blobs)
    for p in parts:
        p.enable_self_gravity(2, factor)
    for p in parts:
        p.set_state(v0, e0)
    for n in (15, 25):
        for p in parts:
            p.step(n)
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_POTENTIAL):
        total, want = sum(p.field(fid) for p in parts), ref.field(fid)
        if world == 1:
            assert np.array_equal(total, want), fid
        else:
            assert float(np.abs(total - want).max() / np.abs(want).max()) <= 1e-10, fid
    for p in parts:
        p.synchronize()
