"""Batched ensemble vs one-member-at-a-time: python scripts/ensemble_timing.py [level] [members ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
counts = [int(a) for a in sys.argv[2:] if not a.startswith("l")] or [32]
lmax = ([int(a[1:]) for a in sys.argv[2:] if a.startswith("l")] or [0])[0]      # e.g. l8: self-gravity to degree 8
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
base = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 100e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
            obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
single = odis.Solver(mesh, base)
single.step(50)
ms1 = single.step_timed(600) / 600
single.close()
for M in counts:
    hs = np.logspace(3, 5, M)
    plist = [dict(base, h=float(hs[m]), alpha=float(10 ** (-11 + 5 * ((m * 7) % M) / max(M - 1, 1)))) for m in range(M)]
    ens = odis.Ensemble(mesh, plist)
    if lmax:
        ens.enable_self_gravity(lmax, 0.1 * np.ones(lmax + 1))
    ens.step(20)
    ms = ens.step_timed(200) / 200
    info = ens.info()
    print(f"level {level} ({mesh.n_cells} cells) M={M} sh degree {lmax}: batched step {ms * 1e3:.1f} us = {ms * 1e3 / M:.2f} us per member-step "
          f"({M / ms * 1e3:.0f} member-steps/s; {info['algorithmic_bytes_per_step'] / (ms * 1e-3) / 1e9:.0f} GB/s alg) | one member alone "
          f"{ms1 * 1e3:.1f} us/step -> batching gain x{ms1 * M / ms:.1f}; device {info['device_bytes'] / 1e6:.0f} MB", flush=True)
    ens.close()
