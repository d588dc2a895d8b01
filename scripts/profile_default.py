"""A few steps of the headline workload with the default kernel selection, launched one by one (no graph replay), for ncu:
python scripts/profile_default.py [level] [l_max]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
l_max = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1, kernel_select=8)
s = odis.Solver(mesh, prm)
if l_max >= 2:
    f = 0.1 * np.ones(l_max + 1); f[:2] = 0.0
    s.enable_self_gravity(l_max, f)
s.step(8)
s.synchronize()
print("done", mesh.n_cells, s.launches)
