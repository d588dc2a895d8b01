"""Round-2 first GPU call: validate and time the opt-in kernel selections against the default path.

    python scripts/variants_timing.py [level] [l_max]

For each selection (odis_params.reserved[0] / `kernel_select`): fields after 120 steps against the default selection
(relative difference), then device time per step (CUDA events, graph replay) and the per-launch split. Prints one line
per selection; nothing here is a bench value (bench.py is)."""
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
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
factor = 0.1 * np.ones(l_max + 1)
factor[:2] = 0.0
SELECTIONS = [("default (5 launches/step)", 0), ("3-launch self-gravity (bit 4)", 16), ("3-launch + direct edge kernel (bits 0,4)", 17),
              ("16-bit stencil ids in the staged edge kernel (bit 7)", 128), ("16-bit ids + 3-launch self-gravity (bits 4,7)", 144),
              ("16-bit ids + 3-launch + 64-register cell update (bits 4,6,7)", 208),
              ("L2 prefetch in the cell update (bit 9)", 512), ("L2 prefetch + 64 registers (bits 6,9)", 576),
              ("16-bit ids + L2 prefetch (bits 7,9)", 640), ("3-launch + 16-bit ids + L2 prefetch (bits 4,7,9)", 656)]
ref = None
for name, sel in SELECTIONS:
    s = odis.Solver(mesh, dict(prm, kernel_select=sel))
    s.enable_self_gravity(l_max, factor)
    s.step(120)
    eta, v = s.field(odis.FIELD_ETA), s.field(odis.FIELD_VELOCITY)
    if ref is None:
        ref = (eta, v)
    d_eta = float(np.abs(eta - ref[0]).max() / np.abs(ref[0]).max())
    d_v = float(np.abs(v - ref[1]).max() / np.abs(ref[1]).max())
    l0 = s.launches
    ms = s.step_timed(1200) / 1200
    per_step = (s.launches - l0) / 1200
    e, c, g = s.step_profiled_sh(200)
    print(f"level {level} l_max {l_max} {name}: {ms * 1e3:.1f} us/step, {per_step:.0f} launches/step | edge {e / 200 * 1e3:.1f} cell {c / 200 * 1e3:.1f} "
          f"sh {g / 200 * 1e3:.1f} us | vs default: eta {d_eta:.2e} v {d_v:.2e}", flush=True)
    s.close()
