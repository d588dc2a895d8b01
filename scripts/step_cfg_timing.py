"""One configuration of the headline step per process (the experiment switches are environment variables read once per process):

    [ODIS_B200_MERGED_SYNTH=0] python scripts/step_cfg_timing.py [level] [l_max] [kernel_select]

Prints one line: device time per step under graph replay (CUDA events), launches per step, the per-launch split from
odis_step_profiled_sh (events between launches, no graph) and the difference of eta to the baseline selection after 120 steps.
Nothing here is a bench value (bench.py is)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
l_max = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
factor = 0.1 * np.ones(max(l_max, 2) + 1)
factor[:2] = 0.0


def make(select):
    s = odis.Solver(mesh, dict(prm, kernel_select=select))
    if l_max >= 2:
        s.enable_self_gravity(l_max, factor)
    return s


s = make(sel)
s.step(120)
eta = s.field(odis.FIELD_ETA)
l0 = s.launches
best = min(s.step_timed(1200) / 1200 for _ in range(3))
per_step = (s.launches - l0) / 3600
e, c, g = s.step_profiled_sh(200)
s.close()
ref = make(1)                 # direct-load baseline kernels
ref.step(120)
d = float(np.abs(eta - ref.field(odis.FIELD_ETA)).max() / np.abs(eta).max())
ref.close()
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("ODIS_B200_") and k != "ODIS_B200_LIB")
print(f"level {level} l_max {l_max} select {sel} [{env}]: {best * 1e3:.2f} us/step ({1e3 / best:.0f} steps/s), {per_step:.0f} launches/step | "
      f"edge {e / 200 * 1e3:.1f} cell {c / 200 * 1e3:.1f} sh {g / 200 * 1e3:.1f} us | eta vs baseline kernels: {d:.2e}", flush=True)
