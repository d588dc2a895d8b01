"""Per-step time of the nonlinear branch vs the linear one: python scripts/nonlinear_timing.py [level]
Variants: linear; nonlinear with the baseline selection (six gather kernels + diagnostics + potential pass = 8 launches per step); the
default (4 launches). (The intermediate form — 4 gather launches with the two passes of their own — measured 100.5 us per step at
163,842 cells against 113.8 and 80.0, profiles/r02/gpurun_r02h, and is gone.)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = sys.argv[2] if len(sys.argv) > 2 else ""          # substring of a variant's name: run only those (profiling)
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3                                  # the headline physics (Enceladus ocean, ECC tide, linear drag), free surface: stays finite
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
t0 = time.time()
nl = odis.nonlinear_tables(mesh, 0.5)
t_tab = time.time() - t0
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047,
           obl=0.0, shell_thickness=0.0, semimajor_axis=0.0, potential=5, friction=0, surface=0, init_load=0, reorder=1)
out, fields = {}, {}
for name, adv, sel in (("linear", False, 0), ("nonlinear 8-launch baseline selection", True, 1), ("nonlinear default (folded, 4 launches)", True, 0)):
    if only and only not in name:
        continue
    s = odis.Solver(mesh, dict(prm, kernel_select=sel))
    if adv:
        s.enable_advection(nl)
    l0 = s.launches
    s.step(40)
    per_step = (s.launches - l0) / 40
    # every timed chunk starts from the zero state again: at this resolution the reference's nonlinear scheme (no viscosity) goes
    # unstable after a few hundred steps with any physics, and a timing over NaN fields would prove little
    times = []
    for _ in range(3):
        s.set_state()
        s.step(20)
        times.append(s.step_timed(80) / 80)
    ms = min(times)
    out[name] = ms
    eta = s.field(odis.FIELD_ETA)
    fields[name] = (eta, s.field(odis.FIELD_VELOCITY), s.dissipation_series())
    print(f"level {level} ({mesh.n_cells} cells) {name}: {ms * 1e3:.1f} us/step, {per_step:.0f} launches/step, max|eta| {np.abs(eta).max():.4e}, "
          f"finite {bool(np.isfinite(eta).all())}", flush=True)
    s.close()
names = [n for n in out if n.startswith("nonlinear")]
if len(names) < 2 or "linear" not in out:
    sys.exit(0)
same = all(np.array_equal(fields[n][0], fields[names[0]][0], equal_nan=True) and np.array_equal(fields[n][1], fields[names[0]][1], equal_nan=True) for n in names[1:])
ser = max(float(np.nanmax(np.abs(fields[n][2] - fields[names[0]][2])) / max(np.nanmax(np.abs(fields[names[0]][2])), 1e-300)) for n in names[1:])
print(f"nonlinear variants: fields bit-identical {same}; dissipation series max rel diff {ser:.2e}")
print(f"nonlinear tables built on the host in {t_tab:.2f} s; nonlinear default / linear step time = {out[names[-1]] / out['linear']:.2f}")
