"""Narrow stencil ids of the staged edge kernel (the default since round 2: measured -2.5 us per step at 655,362 cells; odis_params.reserved[0]
bit 7, `kernel_select=128`, switches them off): the ten stencil ids of an edge travel as 16-bit offsets from the edge's own id, one 2560-byte
bulk copy per 128-edge tile; tiles in which an offset does not fit are flagged "wide" and read from the int rows. The id only addresses the
gather, so fields must be bit-identical to the oracle / the 32-bit selection. Bit 8 (tests only) narrows the range to +-1023 so that small
grids have wide AND narrow tiles in one launch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRM = dict(g=0.113, h=38e3, alpha=1e-6, dt=30.0, radius=252.1e3, omega=5.307e-5, love_reduct=0.95, ecc=0.0047, obl=0.002,
           shell_thickness=0.0, potential=8, friction=1, surface=0, init_load=0)


@pytest.mark.parametrize("level", [3, 5, 6])
@pytest.mark.parametrize("kernel_select", [0, 256, 8])       # narrow; narrow + forced wide tiles; narrow without graph replay
def test_narrow_ids_match_oracle(odis, level, kernel_select):
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(level)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    s = odis.Solver(mesh, dict(PRM, reorder=1, semimajor_axis=0.0, kernel_select=kernel_select))
    o = LteOracle(mesh.tables, PRM)
    o.set_state()
    series = o.step(40)
    s.step(40)
    assert np.array_equal(s.field(odis.FIELD_VELOCITY), o.field(0)) and np.array_equal(s.field(odis.FIELD_ETA), o.field(1))
    assert np.array_equal(s.field(odis.FIELD_DVDT), o.field(2)) and np.array_equal(s.field(odis.FIELD_DETADT), o.field(3))
    assert np.allclose(s.dissipation_series()[1:], series, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("kernel_select", [0, 256])
def test_narrow_ids_random_state_equals_wide_ids(odis, kernel_select):
    """Loaded random state (every stencil slot contributes), 3 x 37 steps so that stage reuse and the AB3 history roles rotate."""
    pos, fr, cen = odis.generate_grid(6)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    rng = np.random.default_rng(11)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    out = []
    for sel in (128, kernel_select):
        s = odis.Solver(mesh, dict(PRM, reorder=1, semimajor_axis=0.0, kernel_select=sel))
        s.set_state(v0, e0)
        for _ in range(3):
            s.step(37)
        out.append([s.field(f) for f in (odis.FIELD_VELOCITY, odis.FIELD_ETA, odis.FIELD_DVDT, odis.FIELD_DISSIPATION)] + [s.dissipation_series()])
        s.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("world", [2, 4])
def test_narrow_ids_on_a_partitioned_grid(odis, world):
    """Ghost edges sit behind the own edges in a rank's numbering, so boundary tiles are wide by construction: both paths and the
    in-kernel halo push in one run. Bit-identical to the single-device default."""
    from test_multigpu import _device_count
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    pos, fr, cen = odis.generate_grid(6)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, PRM["radius"])
    prm = dict(PRM, reorder=1, semimajor_axis=0.0, friction=0)
    rng = np.random.default_rng(5)
    v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
    ref = odis.Solver(mesh, dict(prm, kernel_select=128), device=0)
    ref.set_state(v0, e0)
    ref.step(50)
    parts = [odis.Solver(mesh, dict(prm, kernel_select=256), device=k, rank=k, world=world) for k in range(world)]
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
