"""Quick GPU sanity run: CUDA path vs the CPU oracle on generated grids (developer aid, not a test)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis
from oracle.lte_oracle import LteOracle

def run(level, nsteps, potential=5, reorder=1, seed=0):
    pos, fr, cen = odis.generate_grid(level)
    r = 252.1e3
    t0 = time.time(); mesh = odis.Mesh.from_arrays(pos, fr, cen, r); t1 = time.time()
    prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=40.0, radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001,
               shell_thickness=0.0, potential=potential, friction=0, surface=0, init_load=0)
    dmin = mesh.tables['face_node_dist'].min(); prm['dt'] = 0.25 * dmin / np.sqrt(prm['g'] * prm['h'])
    s = odis.Solver(mesh, dict(prm, reorder=reorder, semimajor_axis=0.0))
    rng = np.random.default_rng(seed)
    v0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2; e0 = rng.uniform(-1, 1, mesh.n_cells)
    s.set_state(v0, e0); s.step(nsteps); v = s.field(0); eta = s.field(1); ser = s.dissipation_series()
    out = dict(level=level, N=mesh.n_cells, mesh_s=round(t1 - t0, 2))
    if mesh.n_cells <= 200000:
        o = LteOracle(mesh.tables, prm); o.set_state(v0, e0); t2 = time.time(); so = o.step(nsteps); t3 = time.time()
        vo, eo = o.field(0), o.field(1)
        out.update(v_maxabs=float(np.abs(v - vo).max()), v_scale=float(np.abs(vo).max()), eta_maxabs=float(np.abs(eta - eo).max()),
                   eta_scale=float(np.abs(eo).max()), diss_rel=float(np.abs(ser[1:] - so).max() / np.abs(so).max()),
                   vavg_maxabs=float(np.abs(s.field(4) - o.field(4)).max()), ediss_rel=float(np.abs(s.field(5) - o.field(5)).max() / np.abs(o.field(5)).max()),
                   dvdt_maxabs=float(np.abs(s.field(2) - o.field(2)).max()), detadt_maxabs=float(np.abs(s.field(3) - o.field(3)).max()),
                   U_maxabs=float(np.abs(s.field(6) - o.field(6)).max()) if False else None,
                   oracle_steps_per_s=round(nsteps / (t3 - t2), 2))
    ms = s.step_timed(200); out['gpu_steps_per_s'] = round(200 / ms * 1e3, 1)
    dev, alg = s.footprint(); out['alg_GBps'] = round(alg * 200 / ms * 1e3 / 1e9, 1)
    print(out, flush=True)

if __name__ == '__main__':
    for pot in (5, 0, 1, 8, 9, 16):
        run(4, 30, potential=pot)
    run(5, 100); run(5, 100, reorder=0)
    run(7, 50); run(8, 20)
    run(9, 10); run(9, 10, reorder=0)
