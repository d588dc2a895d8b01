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


@pytest.mark.parametrize("world", [2, 4, 8])
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
@pytest.mark.parametrize("world", [2, 4, 8])
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


@pytest.mark.parametrize("with_sg", [False, True])
def test_partitioned_pipelined_io_matches_synchronous_calls(odis, with_sg):
    """odis_stage_state / odis_commit_state / odis_snapshot_begin / _wait on a partitioned grid (2 ranks): every rank packs and uploads its
    own share of the NEXT interval's state while the current interval steps, and its own entries come back compact through the snapshot
    slots. Three intervals from three different states: the fields of every interval must equal, bit for bit, what the synchronous calls
    (odis_set_state / odis_step / odis_get_field) give on the same partitioned solvers — and, without the self-gravity term, the
    single-GPU run."""
    world = 2
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(5)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=40.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001, shell_thickness=0.0,
               semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
    rng = np.random.default_rng(11)
    N, F, S = mesh.n_cells, mesh.n_edges, 14
    states = [(rng.uniform(-1, 1, F) * 1e-2, rng.uniform(-1, 1, N), rng.uniform(-1, 1, (F, 3)).ravel() * 1e-6, rng.uniform(-1, 1, (N, 3)).ravel() * 1e-4)
              for _ in range(3)]
    l_max, factor = 2, np.array([0.0, 0.0, 0.3])

    def make():
        parts = [odis.Solver(mesh, prm, device=k, rank=k, world=world) for k in range(world)]
        blobs = [p.halo_blob() for p in parts]
        for p in parts:
            p.halo_connect(blobs)
        if with_sg:
            for p in parts:
                p.enable_self_gravity(l_max, factor)
        return parts

    # synchronous calls
    parts = make()
    want = []
    for k, st in enumerate(states):
        for p in parts:
            p.set_state(*st, iter=5 * k)
        for p in parts:
            p.step(S)
        want.append({f: sum(p.field(fid) for p in parts) for f, fid in (("eta", odis.FIELD_ETA), ("velocity", odis.FIELD_VELOCITY))})
        want[-1]["dissipation_avg"] = sum(p.dissipation_avg() for p in parts)
    for p in parts:
        p.synchronize()
        p.close()

    # pipelined: stage k+1 while interval k steps, snapshot k while interval k+1 steps
    parts = make()
    maps = [p.partition_map() for p in parts]
    fields = odis.Solver.SNAP_ETA | odis.Solver.SNAP_VELOCITY
    got = []

    def collect(slot):
        snaps = [p.snapshot_wait(slot) for p in parts]
        eta, vel = np.zeros(N), np.zeros(F)
        for (cm, em), sn in zip(maps, snaps):
            assert sn["eta"].shape == (cm.size,) and sn["velocity"].shape == (em.size,)
            eta[cm] = sn["eta"]
            vel[em] = sn["velocity"]
        got.append({"eta": eta, "velocity": vel, "dissipation_avg": sum(sn["dissipation_avg"] for sn in snaps)})

    for p in parts:
        p.stage_state(*states[0])
    for k in range(len(states)):
        for p in parts:
            p.commit_state(iter=5 * k)
        if k + 1 < len(states):
            for p in parts:
                p.stage_state(*states[k + 1])
        for p in parts:
            p.step(S)
        for p in parts:
            p.snapshot_begin(k & 1, fields)
        if k > 0:
            collect((k - 1) & 1)
    collect((len(states) - 1) & 1)
    for p in parts:
        p.synchronize()
    for k in range(len(states)):
        assert np.array_equal(got[k]["eta"], want[k]["eta"]), k
        assert np.array_equal(got[k]["velocity"], want[k]["velocity"]), k
        assert got[k]["dissipation_avg"] == want[k]["dissipation_avg"], k
    assert sorted(np.concatenate([m[0] for m in maps]).tolist()) == list(range(N))         # the own cells / edges partition the grid
    assert sorted(np.concatenate([m[1] for m in maps]).tolist()) == list(range(F))
    for k, (cm, em) in enumerate(maps):                                                     # ... exactly as the host-only plan says
        plan = odis.partition_plan(mesh, k, world)
        assert np.array_equal(cm, plan["local_cell_ref"][:plan["own_cells"]]) and np.array_equal(em, plan["local_edge_ref"][:plan["own_edges"]])
    if not with_sg:
        ref = odis.Solver(mesh, prm, device=0)
        ref.set_state(*states[-1], iter=5 * (len(states) - 1))
        ref.step(S)
        assert np.array_equal(got[-1]["eta"], ref.field(odis.FIELD_ETA)) and np.array_equal(got[-1]["velocity"], ref.field(odis.FIELD_VELOCITY))
    for p in parts:
        p.close()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("level", [3, 4])
def test_partitioned_small_grids_stay_in_step(odis, level, world):
    """Grids so small that every CTA of the staged kernels is through its one tile within a few microseconds: the exchange protocol must
    not depend on a kernel lasting longer than a system-scope fence (round 2: the halo warp published an epoch one too high when the last
    CTA had already counted the step — invisible at 40,962 cells, a stale ghost read at 2,562). 400 steps, graph replay included; fields
    bit-identical to the single-GPU run, every entry of the dissipation series within 1e-12."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(level)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    dmin = float(mesh.tables["face_node_dist"].min())
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001,
               shell_thickness=0.0, semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
    rng = np.random.default_rng(5)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, prm, device=0)
    ref.set_state(v0, e0)
    parts = [odis.Solver(mesh, prm, device=k, rank=k, world=world) for k in range(world)]
    blobs = [p.halo_blob() for p in parts]
    for p in parts:
        p.halo_connect(blobs)
    for p in parts:
        p.set_state(v0, e0)
    for n in (7, 193, 200):
        ref.step(n)
        for p in parts:
            p.step(n)
        for fid in (odis.FIELD_VELOCITY, odis.FIELD_ETA):
            assert np.array_equal(sum(p.field(fid) for p in parts), ref.field(fid)), (fid, n)
        assert np.isclose(sum(p.dissipation_avg() for p in parts), ref.dissipation_avg(), rtol=1e-12, atol=0.0), n
    series = sum(p.dissipation_series() for p in parts)
    assert np.allclose(series, ref.dissipation_series(), rtol=1e-12, atol=0.0)
    for p in parts:
        p.synchronize()
