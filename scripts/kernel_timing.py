"""Per-kernel timings of the headline workload: python scripts/kernel_timing.py [level] [block_threads ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis
from oracle.lte_oracle import LteOracle

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
blocks = [int(a) for a in sys.argv[2:]] or [256]
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
if level <= 6:   # correctness guard on small grids
    o = LteOracle(mesh.tables, {k: v for k, v in prm.items() if k not in ("semimajor_axis", "reorder")}); o.set_state(); o.step(20)
for bt in blocks:
    # block size 0: fused one-launch step; -1: default (two-launch staged edge + direct cell); -2: two-launch both staged;
    # > 0: two-launch both direct with that block size
    # -3: default kernels without CUDA graph replay
    sel = 1 if bt > 0 else (2 if bt == -2 else (0 if bt == -1 else (8 if bt == -3 else 4)))
    s = odis.Solver(mesh, dict(prm, block_threads=max(bt, 0), kernel_select=sel))
    if level <= 6:
        s.step(20)
        print("bit-identical:", np.array_equal(s.field(0), o.field(0)) and np.array_equal(s.field(1), o.field(1)))
        s.set_state()
    s.step(50)
    ms = s.step_timed(500)
    e, c = s.step_profiled(300)
    F, N = mesh.n_edges, mesh.n_cells
    print(f"lib={os.environ.get('ODIS_B200_LIB','default')} level={level} block={bt}: {500/ms*1e3:.0f} steps/s  step {ms/500*1e3:.1f} us | edge {e/300*1e3:.1f} us "
          f"({200*F/(e/300*1e-3)/1e9:.0f} GB/s alg) cell {c/300*1e3:.1f} us ({128*N/(c/300*1e-3)/1e9:.0f} GB/s alg)", flush=True)
    s.close()
