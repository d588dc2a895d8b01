"""Domain-decomposed stepping on 2+ GPUs of one box: fields must be bit-identical to the single-GPU run
(the halo exchange changes where a value lives, never its arithmetic). Skipped when fewer than 2 devices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_count():
    import os
    emulated = int(os.environ.get("ODIS_B200_EMULATED_DEVICES", "0"))      # set by tests/test_emulated_kernels.py only: every stream of the
    if emulated:                                                            # emulation is a "device" of its own
        return emulated
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_run_matches_single_gpu(odis, world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=40.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001, shell_thickness=0.0,
               semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
    rng = np.random.default_rng(3)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, prm, device=0)
    ref.set_state(v0, e0)
    ref.step(60)
    parts = [odis.Solver(mesh, prm, device=k, rank=k, world=world) for k in range(world)]
    blobs = [p.halo_blob() for p in parts]
    for p in parts:
        p.halo_connect(blobs)
    for p in parts:
        p.set_state(v0, e0)
    for n in (25, 35):                                   # all ranks enqueue the same number of steps, in turn
        for p in parts:
            p.step(n)
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DETADT, odis.FIELD_VELOCITY_EN, odis.FIELD_DISSIPATION):
        total = sum(p.field(fid) for p in parts)         # each rank fills its own entries, zeros elsewhere
        assert np.array_equal(total, ref.field(fid)), fid
    series = sum(p.dissipation_series() for p in parts)
    assert np.allclose(series, ref.dissipation_series(), rtol=1e-12, atol=0.0)
    info = [p.partition() for p in parts]
    assert sum(i["own_cells"] for i in info) == mesh.n_cells and all(i["n_peers"] >= 1 for i in info)


@pytest.mark.parametrize("stored", [False, True])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_self_gravity_matches_single_gpu(odis, world, stored):
    """Self-gravity term on a partitioned grid: the harmonic sums are all-reduced through peer memory inside the kernels.
    The sums group differently than on one GPU, so fields agree to rounding (1e-10 asserted, BASELINE.json's bar), and every
    rank must hold bit-identical coefficients."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=40.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001, shell_thickness=0.0,
               semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
    l_max = 4
    factor = 0.5 / (1.0 + 0.2 * np.arange(l_max + 1))
    rng = np.random.default_rng(3)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, prm, device=0)
    ref.enable_self_gravity(l_max, factor, stored_basis=stored)
    ref.set_state(v0, e0)
    ref.step(60)
    parts = [odis.Solver(mesh, prm, device=k, rank=k, world=world) for k in range(world)]
    blobs = [p.halo_blob() for p in parts]
    for p in parts:
        p.halo_connect(blobs)
    for p in parts:
        p.enable_self_gravity(l_max, factor, stored_basis=stored)
    for p in parts:
        p.set_state(v0, e0)
    for n in (25, 35):
        for p in parts:
            p.step(n)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_POTENTIAL):
        total = sum(p.field(fid) for p in parts)
        assert rel(total, ref.field(fid)) <= 1e-10, fid
    coeffs = [p.sh_coefficients() for p in parts]
    for c in coeffs[1:]:
        assert np.array_equal(c, coeffs[0])
    assert np.abs(coeffs[0] - ref.sh_coefficients()).max() <= 1e-11 * np.abs(ref.sh_coefficients()).max()
    for p in parts:
        p.synchronize()
