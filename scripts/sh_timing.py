"""Per-launch timing of a step with the self-gravity term: python scripts/sh_timing.py [level] [l_max ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
lmaxes = [int(a) for a in sys.argv[2:]] or [2, 8]
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
N = mesh.n_cells
for l_max, stored in [(0, False)] + [(l, st) for l in lmaxes for st in (False, True)]:
    s = odis.Solver(mesh, prm)
    if l_max:
        s.enable_self_gravity(l_max, 0.1 * np.ones(l_max + 1), stored_basis=stored)
    s.step(100)
    ms = s.step_timed(1200) / 1200
    e, c, g = s.step_profiled_sh(200)
    rows = (l_max + 1) ** 2 if l_max else 0
    sh_bytes = (8 * rows * N + 8 * max(rows - 4, 0) * N + 40 * N if stored else 104 * N) if l_max else 0
    _, alg = s.footprint()
    print(f"level {level} l_max {l_max} {'stored' if stored else 'matrix-free'}: step {ms * 1e3:.1f} us ({alg / ms / 1e6:.0f} GB/s alg) | edge {e / 200 * 1e3:.1f} cell {c / 200 * 1e3:.1f} "
          f"sh {g / 200 * 1e3:.1f} us" + (f" (sh alone {sh_bytes / (g / 200 * 1e-3) / 1e9:.0f} GB/s)" if l_max and g > 0 else ""), flush=True)
    s.close()
