"""Per-step time of the nonlinear branch vs the linear one: python scripts/nonlinear_timing.py [level]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pos, fr, cen = odis.generate_grid(level)
r = 6.37122e6
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
t0 = time.time()
nl = odis.nonlinear_tables(mesh, 0.5)
t_tab = time.time() - t0
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=9.80616, h=8e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(9.80616 * 8e3), radius=r, omega=7.292e-5, love_reduct=1.0, ecc=0.01,
           obl=np.deg2rad(-2.0), shell_thickness=0.0, semimajor_axis=0.0, potential=1, friction=0, surface=0, init_load=0, reorder=1)
out = []
for adv in (False, True, "4-launch variant (kernel_select=32)"):
    s = odis.Solver(mesh, dict(prm, kernel_select=32) if isinstance(adv, str) else prm)
    if adv:
        s.enable_advection(nl)
    s.step(50)
    ms = s.step_timed(400) / 400
    out.append(ms)
    eta = s.field(odis.FIELD_ETA)
    print(f"level {level} ({mesh.n_cells} cells) advection={adv}: {ms * 1e3:.1f} us/step, max|eta| {np.abs(eta).max():.4e}, finite {bool(np.isfinite(eta).all())}", flush=True)
    s.close()
print(f"nonlinear tables built on the host in {t_tab:.2f} s; nonlinear / linear step time = {out[1] / out[0]:.2f} (4-launch variant: {out[2] / out[0]:.2f})")
