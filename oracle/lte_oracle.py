"""TEST INFRASTRUCTURE — ctypes front end of oracle/lte_oracle.c (the CPU restatement of the reference's
per-step arithmetic). Mesh tables are passed as a dict of numpy arrays with the reference's names."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .build_oracle import build_oracle

_MESH_FIELDS = [
    ("node_friends", np.int32), ("faces", np.int32), ("node_face_dir", np.int32), ("face_nodes", np.int32),
    ("face_interp_friends", np.int32), ("face_interp_weights", np.float64), ("face_len", np.float64),
    ("face_node_dist", np.float64), ("face_centre_m", np.float64), ("face_centre_pos_sph", np.float64),
    ("face_area", np.float64), ("face_normal_vec_map", np.float64), ("control_volume_surf_area_map", np.float64),
    ("node_pos_sph", np.float64),
]


class _Mesh(C.Structure):
    _fields_ = [("n_cells", C.c_int), ("n_edges", C.c_int)] + [(n, C.c_void_p) for n, _ in _MESH_FIELDS]


class _Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl", "shell_thickness",
                                          "semimajor_axis")] + \
               [(n, C.c_int) for n in ("potential", "friction", "surface", "init_load")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_oracle())
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params)]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        lib.oracle_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_get_dissipation_avg.restype = C.c_double
        lib.oracle_get_dissipation_avg.argtypes = [C.c_void_p]
        lib.oracle_set_sh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.oracle_get_sh_b.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_set_nonlinear.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 14
        lib.oracle_get_iter.restype = C.c_long
        lib.oracle_get_iter.argtypes = [C.c_void_p]
        lib.oracle_op_update_momentum.argtypes = [C.c_void_p] * 4
        lib.oracle_op_update_eta.argtypes = [C.c_void_p] * 3
        lib.oracle_op_forcing.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        lib.oracle_op_drag_forcing.argtypes = [C.c_void_p] * 4
        lib.oracle_op_integrate_ab3_scalar.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        lib.oracle_op_interpolate_velocity.argtypes = [C.c_void_p] * 3
        lib.oracle_op_update_energy.restype = C.c_double
        lib.oracle_op_update_energy.argtypes = [C.c_void_p] * 4
        _lib = lib
    return _lib


FIELD_SHAPES = {0: ("F",), 1: ("N",), 2: ("F", 3), 3: ("N", 3), 4: ("F", 2), 5: ("F",), 6: ("N",)}


class LteOracle:
    """oracle = LteOracle(tables, params); oracle.set_state(...); oracle.step(n); oracle.field(0)"""

    def __init__(self, tables: dict, params: dict):
        lib = _load()
        self._keep = {}
        m = _Mesh()
        m.n_cells = int(tables["node_friends"].shape[0])
        m.n_edges = int(tables["face_nodes"].shape[0])
        for name, dt in _MESH_FIELDS:
            a = np.ascontiguousarray(tables[name], dtype=dt)
            self._keep[name] = a
            setattr(m, name, a.ctypes.data)
        p = _Params()
        for n, _ in _Params._fields_:
            setattr(p, n, params.get(n, 0))
        self.N, self.F = m.n_cells, m.n_edges
        self._h = lib.oracle_create(C.byref(m), C.byref(p))

    def __del__(self):
        if getattr(self, "_h", None):
            _load().oracle_destroy(self._h)
            self._h = None

    @staticmethod
    def _ptr(a, n):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == n, (a.size, n)
        return a, a.ctypes.data

    def set_state(self, v=None, eta=None, dvdt=None, detadt=None, iter: int = 0):
        k = [self._ptr(v, self.F), self._ptr(eta, self.N), self._ptr(dvdt, self.F * 3), self._ptr(detadt, self.N * 3)]
        _load().oracle_set_state(self._h, k[0][1], k[1][1], k[2][1], k[3][1], iter)

    def set_nonlinear(self, nl: dict) -> None:
        """Nonlinear branch (advection; true). nl: the reference's operators and vertex tables, keys as the reference names them:
        operatorCurl / operatorRBFinterp / operatorDirectionalSecondDeriv + '.indptr' '.indices' '.data', vertex_sinlat,
        vertex_area, vertex_R, vertex_nodes, face_vertexes."""
        a = []
        for op in ("operatorCurl", "operatorRBFinterp", "operatorDirectionalSecondDeriv"):
            for part, dt in ((".indptr", np.int32), (".indices", np.int32), (".data", np.float64)):
                a.append(np.ascontiguousarray(nl[op + part], dtype=dt))
        for name, dt in (("vertex_sinlat", np.float64), ("vertex_area", np.float64), ("vertex_R", np.float64), ("vertex_nodes", np.int32),
                         ("face_vertexes", np.int32)):
            a.append(np.ascontiguousarray(nl[name], dtype=dt))
        self._keep["nl"] = a
        _load().oracle_set_nonlinear(self._h, int(a[9].shape[0]), *[x.ctypes.data for x in a])

    def set_self_gravity(self, Y, T) -> None:
        """Y [rows][N] basis, T [rows][rows] = factor_l * (Y Y^T)^-1 (oracle/sh_oracle.py)."""
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        T = np.ascontiguousarray(T, dtype=np.float64)
        assert Y.shape[1] == self.N and T.shape == (Y.shape[0], Y.shape[0])
        self._keep["sh_Y"], self._keep["sh_T"] = Y, T
        _load().oracle_set_sh(self._h, Y.shape[0], Y.ctypes.data, T.ctypes.data)

    def step(self, nsteps: int) -> np.ndarray:
        series = np.zeros(nsteps, dtype=np.float64)
        _load().oracle_step(self._h, nsteps, series.ctypes.data)
        return series

    def field(self, fid: int) -> np.ndarray:
        shape = tuple(self.F if d == "F" else self.N if d == "N" else d for d in FIELD_SHAPES[fid])
        out = np.zeros(shape, dtype=np.float64)
        _load().oracle_get_field(self._h, fid, out.ctypes.data)
        return out

    def dissipation_avg(self) -> float:
        return float(_load().oracle_get_dissipation_avg(self._h))

    @property
    def iter(self) -> int:
        return int(_load().oracle_get_iter(self._h))

    # ---- the loop-level functions one at a time (checkers of the odis_op_* entry points) ----
    def updateMomentum(self, v, eta) -> np.ndarray:
        a, b = self._ptr(v, self.F), self._ptr(eta, self.N)
        out = np.zeros(self.F)
        _load().oracle_op_update_momentum(self._h, a[1], b[1], out.ctypes.data)
        return out

    def updateEta(self, v) -> np.ndarray:
        a = self._ptr(v, self.F)
        out = np.zeros(self.N)
        _load().oracle_op_update_eta(self._h, a[1], out.ctypes.data)
        return out

    def forcing(self, time: float) -> np.ndarray:
        out = np.zeros(self.N)
        _load().oracle_op_forcing(self._h, float(time), out.ctypes.data)
        return out

    def dragForcing(self, v, potential) -> np.ndarray:
        a, b = self._ptr(v, self.F), self._ptr(potential, self.N)
        out = np.zeros(self.F)
        _load().oracle_op_drag_forcing(self._h, a[1], b[1], out.ctypes.data)
        return out

    def integrateAB3scalar(self, solution, dsolution_dt, iter: int):
        sol = np.array(solution, dtype=np.float64, order="C").ravel()
        hist = np.array(dsolution_dt, dtype=np.float64, order="C")
        assert hist.size == 3 * sol.size
        _load().oracle_op_integrate_ab3_scalar(self._h, sol.ctypes.data, hist.ctypes.data, iter, sol.size)
        return sol, hist.reshape(sol.size, 3)

    def interpolateVelocity(self, v) -> np.ndarray:
        a = self._ptr(v, self.F)
        out = np.zeros((self.F, 2))
        _load().oracle_op_interpolate_velocity(self._h, a[1], out.ctypes.data)
        return out

    def updateEnergy(self, v_avg, areas):
        a, b = self._ptr(v_avg, self.F * 2), self._ptr(areas, self.F)
        out = np.zeros(self.F)
        avg = _load().oracle_op_update_energy(self._h, a[1], b[1], out.ctypes.data)
        return out, float(avg)
