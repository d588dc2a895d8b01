"""odis_params.block_threads (128 default, 256, 512) with the direct-load kernels: same fields to the bit, energy sum within the
tolerance of the parallel tree, on grids whose edge count is not a multiple of the block size (the last block's warps beyond the
edge range must find room in the partial-sum buffer)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [3, 4, 5])
@pytest.mark.parametrize("block_threads", [256, 512])
@pytest.mark.parametrize("kernel_select", [1, 9])          # direct-load edge kernel, with and without graph replay; the default
                                                               # kernels with the register-capped (50 % occupancy) cell update
def test_block_sizes_give_identical_fields(odis, level, block_threads, kernel_select):
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(level)
    r = 252.1e3
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=30.0, radius=r, omega=5.307e-5, love_reduct=0.95, ecc=0.0047, obl=0.002,
               shell_thickness=0.0, potential=8, friction=1, surface=0, init_load=0)
    s = odis.Solver(mesh, dict(prm, reorder=1, semimajor_axis=0.0, block_threads=block_threads, kernel_select=kernel_select))
    o = LteOracle(mesh.tables, prm)
    o.set_state()
    series = o.step(40)
    s.step(40)
    assert np.array_equal(s.field(odis.FIELD_VELOCITY), o.field(0)) and np.array_equal(s.field(odis.FIELD_ETA), o.field(1))
    assert np.allclose(s.dissipation_series()[1:], series, rtol=1e-12, atol=0.0)
    assert np.allclose(s.field(odis.FIELD_DISSIPATION), o.field(5), rtol=1e-13, atol=0.0)
