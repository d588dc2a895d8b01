"""Host-side mirror of the reference's interface for the LTE path, over the C ABI.

  reference (C++)                                   here
  ------------------------------------------------  ---------------------------------------------
  new Globals(0)            src/globals.cpp:35      Globals.load(run_dir)
  globals->g.Value() ...    include/globalVar.h:91  Globals["surface gravity"] (input.in key strings)
  new Mesh(globals, ...)    src/mesh.cpp:32         Mesh.from_file(path, radius) / Mesh.from_globals(g)
  grid->face_nodes(i,k) ... include/mesh.h:154      Mesh.tables["face_nodes"][i, k]
  ab3Explicit(globals,grid) src/timeIntegrator.cpp:57   Solver(mesh, params).step(n)  (device resident)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import MESH_VIEW_ARRAYS, MeshView, Params, check

FIELD_VELOCITY, FIELD_ETA, FIELD_DVDT, FIELD_DETADT, FIELD_VELOCITY_EN, FIELD_DISSIPATION, FIELD_POTENTIAL = range(7)
_FIELD_SHAPES = {0: ("F",), 1: ("N",), 2: ("F", 3), 3: ("N", 3), 4: ("F", 2), 5: ("F",), 6: ("N",)}

# enum Potential / Surface / Friction names, include/globals.h:45-76
POTENTIALS = ["OBLIQ", "OBLIQ_WEST", "OBLIQ_EAST", "ECC_RAD", "ECC_LIB", "ECC", "ECC_WEST", "ECC_EAST", "FULL", "FULL2",
              "TOTAL", "ECC_W3", "OBLIQ_W3", "PLANET", "PLANET_OBL", "GENERAL", "NONE"]
SURFACES = ["FREE", "FREE_LOADING", "LID_LOVE", "LID_MEMBR", "LID_NUM", "LID_INF"]
FRICTIONS = ["LINEAR", "QUADRATIC"]


class Globals:
    """Parsed input.in (same keys as the reference, src/globals.cpp:58-203)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, run_dir: str) -> "Globals":
        h = C.c_void_p()
        check(_lib.load().odis_config_load(os.fsencode(run_dir), C.byref(h)))
        return cls(h)

    @classmethod
    def defaults(cls, **overrides) -> "Globals":
        """Titan defaults (Globals(1)) plus input.in-style overrides given as {key: text}; finalised."""
        h = C.c_void_p()
        check(_lib.load().odis_config_create(C.byref(h)))
        g = cls(h)
        for k, v in overrides.items():
            g.set(k, v)
        g.finalize()
        return g

    def set(self, key: str, value) -> None:
        text = ("true" if value else "false") if isinstance(value, bool) else repr(value) if isinstance(value, float) else str(value)
        check(_lib.load().odis_config_set(self._h, key.encode(), text.encode()))

    def finalize(self) -> None:
        check(_lib.load().odis_config_finalize(self._h))

    def __getitem__(self, key: str):
        lib = _lib.load()
        d = C.c_double()
        if lib.odis_config_get_double(self._h, key.encode(), C.byref(d)) == 0:
            return d.value
        i = C.c_int32()
        if lib.odis_config_get_int(self._h, key.encode(), C.byref(i)) == 0:
            return i.value
        if lib.odis_config_get_bool(self._h, key.encode(), C.byref(i)) == 0:
            return bool(i.value)
        buf = C.create_string_buffer(1024)
        check(lib.odis_config_get_string(self._h, key.encode(), buf, 1024))
        return buf.value.decode()

    def enum(self, which: int) -> int:
        i = C.c_int32()
        check(_lib.load().odis_config_get_enum(self._h, which, C.byref(i)))
        return i.value

    fric_type = property(lambda self: self.enum(0))
    surface_type = property(lambda self: self.enum(1))
    solver_type = property(lambda self: self.enum(2))
    tide_type = property(lambda self: self.enum(3))
    initial_condition = property(lambda self: self.enum(4))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "_lib", None) is not None:
            _lib._lib.odis_config_free(self._h)
            self._h = None


def quantise_time_step(period: float, target_dt: float) -> tuple[float, int]:
    """dt and steps per period as Mesh::CalcMaxTimeStep sets them (src/mesh.cpp:1601-1618)."""
    dt, n = C.c_double(), C.c_int32()
    check(_lib.load().odis_quantise_time_step(period, target_dt, C.byref(dt), C.byref(n)))
    return dt.value, n.value


class Mesh:
    """C-grid tables (reference names, include/mesh.h:77-187) as numpy views into library memory."""

    def __init__(self, handle, view=None, keep=None):
        self._h = handle
        self.view = MeshView() if view is None else view
        self._keep = keep                      # numpy arrays the view points into (from_tables)
        if view is None:
            check(_lib.load().odis_mesh_get_view(self._h, C.byref(self.view)))
        self.n_cells, self.n_edges, self.n_vertices = self.view.n_cells, self.view.n_edges, self.view.n_vertices
        self.radius = self.view.radius
        dims = {"N": self.n_cells, "F": self.n_edges, "V": self.n_vertices}
        self.tables = {}
        for name, code, (dim, cols) in MESH_VIEW_ARRAYS:
            n = dims[dim]
            ctype = C.c_double if code == "d" else C.c_int32
            arr = np.ctypeslib.as_array(C.cast(getattr(self.view, name), C.POINTER(ctype)), shape=(n * cols,))
            self.tables[name] = arr.reshape((n, cols)) if cols > 1 else arr
            self.tables[name].flags.writeable = False

    @classmethod
    def from_file(cls, grid_path: str, radius: float, threads: int = 0) -> "Mesh":
        h = C.c_void_p()
        check(_lib.load().odis_mesh_from_file(os.fsencode(grid_path), radius, threads, C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, node_pos_sph, node_friends, centroid_pos_sph, radius: float, threads: int = 0) -> "Mesh":
        pos = np.ascontiguousarray(node_pos_sph, dtype=np.float64)
        fr = np.ascontiguousarray(node_friends, dtype=np.int32)
        cen = np.ascontiguousarray(centroid_pos_sph, dtype=np.float64)
        h = C.c_void_p()
        check(_lib.load().odis_mesh_from_arrays(pos.shape[0], pos.ctypes.data, fr.ctypes.data, cen.ctypes.data, radius, threads, C.byref(h)))
        return cls(h)

    @classmethod
    def from_tables(cls, tables: dict, radius: float) -> "Mesh":
        """A mesh over caller-owned tables (reference names and layouts, as Mesh.tables holds them), e.g. tables built
        once and shared between the ranks of a multi-GPU run. Nothing is recomputed or checked here."""
        view = MeshView()
        keep = {}
        dims = {}
        for name, code, (dim, cols) in MESH_VIEW_ARRAYS:
            a = np.ascontiguousarray(tables[name], dtype=np.float64 if code == "d" else np.int32)
            if a.size % cols:
                raise ValueError(f"{name}: size {a.size} is not a multiple of {cols}")
            if dims.setdefault(dim, a.size // cols) != a.size // cols:
                raise ValueError(f"{name}: {a.size // cols} rows, expected {dims[dim]}")
            keep[name] = a
            setattr(view, name, a.ctypes.data)
        view.n_cells, view.n_edges, view.n_vertices, view.radius = dims["N"], dims["F"], dims["V"], radius
        return cls(None, view=view, keep=keep)

    def save(self, directory: str) -> None:
        """One .npy per table + radius.npy (np.load(..., mmap_mode='r') friendly)."""
        os.makedirs(directory, exist_ok=True)
        for name, a in self.tables.items():
            np.save(os.path.join(directory, name + ".npy"), a)
        np.save(os.path.join(directory, "radius.npy"), np.array([self.radius]))

    @classmethod
    def load(cls, directory: str, mmap: bool = True) -> "Mesh":
        tables = {name: np.load(os.path.join(directory, name + ".npy"), mmap_mode="r" if mmap else None) for name, _, _ in MESH_VIEW_ARRAYS}
        return cls.from_tables(tables, float(np.load(os.path.join(directory, "radius.npy"))[0]))

    @classmethod
    def from_globals(cls, g: Globals, run_dir: str, threads: int = 0) -> "Mesh":
        """input_files/grid_l<L>.txt under run_dir, radius after the surface BCs (src/mesh.cpp:4023)."""
        path = os.path.join(run_dir, "input_files", "grid_l%d.txt" % g["geodesic grid level"])
        return cls.from_file(path, g["radius"], threads)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "_lib", None) is not None:
            self.tables = {}
            _lib._lib.odis_mesh_free(self._h)
            self._h = None


def generate_grid(level: int):
    """Synthetic icosahedral-bisection grid, reference file-name level (10*4^(level-1)+2 cells).
    Returns (node_pos_sph [N,2], node_friends [N,6], centroid_pos_sph [N,6,2]) in radians."""
    lib = _lib.load()
    n = C.c_int32()
    pos, fr, cen = C.POINTER(C.c_double)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
    check(lib.odis_grid_generate(level, C.byref(n), C.byref(pos), C.byref(fr), C.byref(cen)))
    try:
        N = n.value
        a = np.ctypeslib.as_array(pos, shape=(N, 2)).copy()
        b = np.ctypeslib.as_array(fr, shape=(N, 6)).copy()
        c = np.ctypeslib.as_array(cen, shape=(N, 6, 2)).copy()
    finally:
        lib.odis_free(pos); lib.odis_free(fr); lib.odis_free(cen)
    return a, b, c


def write_grid_file(path: str, node_pos_sph, node_friends, centroid_pos_sph) -> None:
    pos = np.ascontiguousarray(node_pos_sph, dtype=np.float64)
    fr = np.ascontiguousarray(node_friends, dtype=np.int32)
    cen = np.ascontiguousarray(centroid_pos_sph, dtype=np.float64)
    check(_lib.load().odis_grid_write_file(os.fsencode(path), pos.shape[0], pos.ctypes.data, fr.ctypes.data, cen.ctypes.data))


def params_from_globals(g: Globals, dt: float, reorder: bool = True) -> dict:
    """The scalars ab3Explicit reads from Globals (src/timeIntegrator.cpp:116-128,198-201)."""
    return dict(g=g["surface gravity"], h=g["ocean thickness"], alpha=g["friction coefficient"], dt=dt, radius=g["radius"],
                omega=g["angular velocity"], love_reduct=g["love reduction factor"], ecc=g["eccentricity"], obl=g["obliquity"],
                shell_thickness=g["shell thickness"], semimajor_axis=g["semimajor axis"], potential=g.tide_type,
                friction=g.fric_type, surface=g.surface_type, init_load=int(g.initial_condition == 1), reorder=int(reorder))


def analytical_state(mesh: "Mesh", params: dict):
    """(v [F], dvdt [F][3], eta [N], detadt [N][3]) of `initial conditions; ANALYTICAL` (OBLIQ_WEST only), as the reference's
    analyticalInitialConditions builds them (src/initialConditions.cpp:146-208). Host only."""
    p = Params()
    for k, val in params.items():
        if k != "kernel_select":
            setattr(p, k, val)
    N, F = mesh.n_cells, mesh.n_edges
    v, dv, eta, de = np.empty(F), np.empty((F, 3)), np.empty(N), np.empty((N, 3))
    check(_lib.load().odis_analytical_state(C.byref(mesh.view), C.byref(p), v.ctypes.data, dv.ctypes.data, eta.ctypes.data, de.ctypes.data))
    return v, dv, eta, de


def sh_basis(pos_sph, l_max: int) -> np.ndarray:
    """Basis rows Y[(l_max+1)^2][n] at (lat, lon) [n][2] (radians): 4-pi normalised, Condon-Shortley phase, degree-major,
    per degree m = 0 then (cos, sin) for m = 1..l (host only)."""
    pos = np.ascontiguousarray(pos_sph, dtype=np.float64)
    out = np.empty(((l_max + 1) ** 2, pos.shape[0]), dtype=np.float64)
    check(_lib.load().odis_sh_basis(pos.shape[0], pos.ctypes.data, l_max, out.ctypes.data))
    return out


def sh_normal_inverse(pos_sph, l_max: int) -> np.ndarray:
    """(Y Y^T)^-1 of the least-squares fit over the given points (host only)."""
    pos = np.ascontiguousarray(pos_sph, dtype=np.float64)
    r = (l_max + 1) ** 2
    out = np.empty((r, r), dtype=np.float64)
    check(_lib.load().odis_sh_normal_inverse(pos.shape[0], pos.ctypes.data, l_max, out.ctypes.data))
    return out


def nonlinear_tables(mesh: Mesh, rbf_eps: float) -> dict:
    """The tables only the nonlinear branch (`advection; true`) reads, built from the mesh on the host: a dict keyed like the
    reference's members ('operatorCurl.indptr' ... 'vertex_sinlat', 'vertex_area'), ready for Solver.enable_advection."""
    if mesh._h is None:
        raise ValueError("nonlinear_tables needs a mesh built by the library (Mesh.from_file / from_arrays)")
    h = C.c_void_p()
    check(_lib.load().odis_nonlinear_create(mesh._h, rbf_eps, C.byref(h)))
    try:
        view = _lib.NonlinearView()
        check(_lib.load().odis_nonlinear_get_view(h, C.byref(view)))
        out = {}
        for name, csr in (("operatorCurl", view.curl), ("operatorRBFinterp", view.rbf_interp), ("operatorDirectionalSecondDeriv", view.directional_second_deriv)):
            ptr = np.ctypeslib.as_array(C.cast(csr.indptr, C.POINTER(C.c_int32)), shape=(csr.n_rows + 1,)).copy()
            nnz = int(ptr[-1])
            out[name + ".indptr"] = ptr
            out[name + ".indices"] = np.ctypeslib.as_array(C.cast(csr.indices, C.POINTER(C.c_int32)), shape=(nnz,)).copy()
            out[name + ".data"] = np.ctypeslib.as_array(C.cast(csr.data, C.POINTER(C.c_double)), shape=(nnz,)).copy()
        V = mesh.n_vertices
        out["vertex_sinlat"] = np.ctypeslib.as_array(C.cast(view.vertex_sinlat, C.POINTER(C.c_double)), shape=(V,)).copy()
        out["vertex_area"] = np.ctypeslib.as_array(C.cast(view.vertex_area, C.POINTER(C.c_double)), shape=(V,)).copy()
        return out
    finally:
        _lib.load().odis_nonlinear_free(h)


def partition_plan(mesh: Mesh, rank: int, world: int, reorder: bool = True) -> dict:
    """Host-only view of the domain decomposition rank `rank` of `world` would use (numpy copies)."""
    lib = _lib.load()
    plan = _lib.PartitionPlan()
    check(lib.odis_partition_plan(C.byref(mesh.view), int(reorder), rank, world, C.byref(plan)))
    try:
        arr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, dtype=np.int32)
        counts = arr(plan.peer_counts, plan.n_peers * 4).reshape(-1, 4)
        se, sc = int(counts[:, 0].sum()), int(counts[:, 1].sum())
        return dict(rank=plan.rank, world=plan.world, own_cells=plan.own_cells, own_edges=plan.own_edges,
                    local_cell_ref=arr(plan.local_cell_ref, plan.local_cells), local_edge_ref=arr(plan.local_edge_ref, plan.local_edges),
                    peer_rank=arr(plan.peer_rank, plan.n_peers), peer_counts=counts,
                    send_edge_ref=arr(plan.send_edge_ref, se), send_edge_slot=arr(plan.send_edge_slot, se),
                    send_cell_ref=arr(plan.send_cell_ref, sc), send_cell_slot=arr(plan.send_cell_slot, sc))
    finally:
        lib.odis_partition_plan_free(C.byref(plan))


class H5Writer:
    """DATA/data.h5 writer (float32, contiguous, fixed-shape datasets; src/outFiles.cpp:138-684)."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        check(_lib.load().odis_h5_create(os.fsencode(path), C.byref(self._h)))

    def add_dataset(self, name: str, shape) -> int:
        dims = (C.c_uint64 * len(shape))(*shape)
        i = C.c_int32()
        check(_lib.load().odis_h5_add_dataset(self._h, name.encode(), len(shape), dims, C.byref(i)))
        return i.value

    def write_rows(self, dataset: int, first_row: int, data) -> None:
        a = np.ascontiguousarray(data, dtype=np.float32)
        nrows = a.shape[0] if a.ndim >= 1 else 1
        check(_lib.load().odis_h5_write_rows(self._h, dataset, first_row, nrows, a.ctypes.data))

    def close(self) -> None:
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            check(_lib.load().odis_h5_close(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(run_dir: str, device: int = 0, reorder: bool = True, echo: bool = False, max_steps: int = 0, self_gravity: int = 0,
        overlap_output: bool = False, n_gpus: int = 1) -> dict:
    """`./ODIS` in run_dir: main -> solveODIS -> ab3Explicit (src/main.cpp:46-68), writing DATA/ and
    InitialConditions/ like the reference. Returns the run summary. self_gravity: 0 as reference HEAD; 1 the
    spherical-harmonic self-gravity / shell-pressure term with input.in's "sh degree" (2: stored-basis kernels).
    overlap_output: dumps are copied out and written while the next output interval is being computed.
    n_gpus > 1: the grid partitioned over that many GPUs of this process (`ODIS --gpus N`)."""
    opt = _lib.RunOptions(device, int(reorder), int(echo), int(self_gravity), max_steps, int(overlap_output), int(n_gpus))
    res = _lib.RunResult()
    check(_lib.load().odis_run(os.fsencode(run_dir), C.byref(opt), C.byref(res)))
    return {n: getattr(res, n) for n, _ in _lib.RunResult._fields_ if n != "reserved"}


class Solver:
    """Device-resident AB3 time stepper (replaces the body of ab3Explicit). Fields cross in reference numbering."""

    def __init__(self, mesh: Mesh, params: dict, device: int = 0, rank: int = 0, world: int = 1):
        """world > 1: this rank's part of a domain-decomposed run (one Solver per GPU). After creating all
        ranks' solvers, exchange halo_blob() among the ranks and call halo_connect() before stepping."""
        self.mesh = mesh
        p = Params()
        for k, v in params.items():
            if k == "kernel_select":             # odis_params.reserved[0], see include/odis_b200.h (0 = default kernels)
                p.reserved[0] = int(v)
            else:
                setattr(p, k, v)
        self.params = p
        self._h = C.c_void_p()
        self.rank, self.world = rank, world
        if world == 1:
            check(_lib.load().odis_create(C.byref(mesh.view), C.byref(p), device, C.byref(self._h)))
        else:
            check(_lib.load().odis_create_partitioned(C.byref(mesh.view), C.byref(p), device, rank, world, C.byref(self._h)))
        self.N, self.F = mesh.n_cells, mesh.n_edges
        self._iter0 = 0

    def halo_blob(self) -> bytes:
        n = _lib.load().odis_halo_blob_size()
        buf = C.create_string_buffer(n)
        check(_lib.load().odis_halo_export(self._h, buf))
        return buf.raw

    def halo_connect(self, blobs) -> None:
        """blobs: the halo_blob() of every rank, ordered by rank (list of bytes or one concatenated bytes)."""
        data = blobs if isinstance(blobs, (bytes, bytearray)) else b"".join(blobs)
        check(_lib.load().odis_halo_connect(self._h, C.c_char_p(bytes(data))))

    def partition(self) -> dict:
        v = [C.c_int32() for _ in range(7)]
        check(_lib.load().odis_get_partition(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("rank", "world", "own_cells", "own_edges", "ghost_cells", "ghost_edges", "n_peers"), (x.value for x in v)))

    def partition_map(self) -> tuple[np.ndarray, np.ndarray]:
        """(reference ids of the own cells, of the own edges) in the order of the solver's device arrays = the order of a partitioned
        solver's compact snapshot arrays."""
        p = self.partition()
        c, e = np.empty(p["own_cells"], dtype=np.int32), np.empty(p["own_edges"], dtype=np.int32)
        check(_lib.load().odis_get_partition_map(self._h, c.ctypes.data, e.ctypes.data))
        return c, e

    @property
    def steps_since_state(self) -> int:
        return self.iter - self._iter0

    @staticmethod
    def _ptr(a, n):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != n:
            raise ValueError(f"array has {a.size} elements, expected {n}")
        return a, a.ctypes.data

    def set_state(self, v=None, eta=None, dvdt=None, detadt=None, iter: int = 0) -> None:
        k = [self._ptr(v, self.F), self._ptr(eta, self.N), self._ptr(dvdt, self.F * 3), self._ptr(detadt, self.N * 3)]
        check(_lib.load().odis_set_state(self._h, k[0][1], k[1][1], k[2][1], k[3][1], iter))
        self._iter0 = iter

    def stage_state(self, v=None, eta=None, dvdt=None, detadt=None) -> None:
        """Starts the host -> device copies of the NEXT state on the second stream and returns; stepping goes on meanwhile. The arrays
        must be float64, contiguous (no conversion copy is made: they have to outlive the call) and unchanged until commit_state."""
        ptrs = []
        for a, n in ((v, self.F), (eta, self.N), (dvdt, self.F * 3), (detadt, self.N * 3)):
            if a is None:
                ptrs.append(None)
                continue
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.size == n):
                raise ValueError("stage_state needs contiguous float64 arrays of the state's sizes")
            ptrs.append(a.ctypes.data)
        self._staged = (v, eta, dvdt, detadt)                    # keep them alive until the commit
        check(_lib.load().odis_stage_state(self._h, *ptrs))

    def commit_state(self, iter: int = 0) -> None:
        """Makes the staged arrays the state, behind the steps enqueued so far; does not wait on the host."""
        check(_lib.load().odis_commit_state(self._h, iter))
        self._iter0 = iter

    def step(self, nsteps: int = 1) -> None:
        check(_lib.load().odis_step(self._h, nsteps))

    def step_timed(self, nsteps: int) -> float:
        ms = C.c_float()
        check(_lib.load().odis_step_timed(self._h, nsteps, C.byref(ms)))
        return ms.value

    def step_profiled(self, nsteps: int) -> tuple[float, float]:
        """(edge-kernel ms, cell-kernel ms) summed over nsteps, each launch timed with its own CUDA events."""
        a, b = C.c_float(), C.c_float()
        check(_lib.load().odis_step_profiled(self._h, nsteps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def step_profiled_sh(self, nsteps: int) -> tuple[float, float, float]:
        """(edge ms, cell ms, self-gravity ms) summed over nsteps."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        check(_lib.load().odis_step_profiled_sh(self._h, nsteps, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def enable_advection(self, nl: dict) -> None:
        """Nonlinear branch (`advection; true`). nl: the reference's operators as CSR and its vertex tables, keyed by the reference's
        names: 'operatorCurl' / 'operatorRBFinterp' / 'operatorDirectionalSecondDeriv' + '.indptr', '.indices', '.data'; 'vertex_sinlat';
        'vertex_area'."""
        keep = []
        view = _lib.NonlinearView()

        def csr(name, n_rows, n_cols):
            ptr = np.ascontiguousarray(nl[name + ".indptr"], dtype=np.int32)
            idx = np.ascontiguousarray(nl[name + ".indices"], dtype=np.int32)
            val = np.ascontiguousarray(nl[name + ".data"], dtype=np.float64)
            if ptr.size != n_rows + 1:
                raise ValueError(f"{name}: expected {n_rows} rows")
            keep.extend([ptr, idx, val])
            return _lib.CsrView(n_rows, n_cols, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data)

        V = self.mesh.n_vertices
        view.curl = csr("operatorCurl", V, self.F)
        view.rbf_interp = csr("operatorRBFinterp", 3 * self.N, self.F)
        view.directional_second_deriv = csr("operatorDirectionalSecondDeriv", 2 * self.F, self.N)
        vs = np.ascontiguousarray(nl["vertex_sinlat"], dtype=np.float64)
        va = np.ascontiguousarray(nl["vertex_area"], dtype=np.float64)
        keep.extend([vs, va])
        view.vertex_sinlat, view.vertex_area = vs.ctypes.data, va.ctypes.data
        check(_lib.load().odis_enable_advection(self._h, C.byref(self.mesh.view), C.byref(view)))

    def enable_self_gravity(self, l_max: int, factor, stored_basis: bool = False) -> None:
        """Spherical-harmonic self-gravity / shell-pressure term (pressureGradientSH): factor[l] for l = 0..l_max
        (globals->shell_factor_beta or loading_factor); degrees 0 and 1 are never applied. stored_basis: keep the basis
        matrix in HBM (two GEMVs per step) instead of rebuilding it per cell (matrix-free, the default)."""
        f = np.ascontiguousarray(factor, dtype=np.float64)
        if f.size != l_max + 1:
            raise ValueError("factor must have l_max + 1 entries")
        check(_lib.load().odis_enable_self_gravity(self._h, C.byref(self.mesh.view), l_max, f.ctypes.data, int(stored_basis)))
        self.sh_rows = (l_max + 1) ** 2

    def sh_coefficients(self) -> np.ndarray:
        """Least-squares harmonic coefficients of the eta the current potential was built from, (l_max+1)^2 values."""
        out = np.empty(self.sh_rows, dtype=np.float64)
        check(_lib.load().odis_get_sh_coefficients(self._h, out.ctypes.data))
        return out

    def field(self, fid: int, out: np.ndarray | None = None) -> np.ndarray:
        """Device -> host, reference numbering. `out`: a C-contiguous float64 array of the field's size to fill in place
        (e.g. page-locked memory, which the copy then reaches by DMA without a staging pass); default: a new array."""
        shape = tuple(self.F if d == "F" else self.N if d == "N" else d for d in _FIELD_SHAPES[fid])
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != int(np.prod(shape)):
            raise ValueError(f"out must be a C-contiguous float64 array of {int(np.prod(shape))} elements")
        check(_lib.load().odis_get_field(self._h, fid, out.ctypes.data))
        return out

    SNAP_ETA, SNAP_VELOCITY_EN, SNAP_DISSIPATION, SNAP_VELOCITY = 1, 2, 4, 8

    def snapshot_begin(self, slot: int, fields: int) -> None:
        """Enqueue the copy-out of `fields` (SNAP_* bits) behind the steps taken so far; returns at once (odis_snapshot_begin)."""
        check(_lib.load().odis_snapshot_begin(self._h, slot, fields))

    def snapshot_wait(self, slot: int, copy: bool = True) -> dict:
        """Block until the slot's copy has landed; arrays are COPIES of the library's page-locked buffers (copy=False: views of them,
        valid until the slot's next snapshot_begin)."""
        view = _lib.SnapshotView()
        check(_lib.load().odis_snapshot_wait(self._h, slot, C.byref(view)))
        out = {"dissipation_avg": view.dissipation_avg, "iter": view.iter}
        N, F = self.N, self.F
        if self.world > 1:                       # the rank's own entries, compact, in partition_map() order
            p = self.partition()
            N, F = p["own_cells"], p["own_edges"]
        for name, n, shape in (("eta", N, (N,)), ("velocity_en", 2 * F, (F, 2)), ("dissipation", F, (F,)), ("velocity", F, (F,))):
            ptr = getattr(view, name)
            if ptr:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,)).reshape(shape)
                out[name] = a.copy() if copy else a
        return out

    def dissipation_avg(self) -> float:
        d = C.c_double()
        check(_lib.load().odis_get_dissipation_avg(self._h, C.byref(d)))
        return d.value

    # ---- operator surface: the reference's loop-level free functions, one call each (odis_op_*). Host arrays in
    # reference numbering; the solver's state is scratch afterwards (call set_state before stepping again). ----
    def updateMomentum(self, v, eta) -> np.ndarray:
        """dvdt = -g G eta + C v  (updateMomentum, src/updateMomentum.cpp:42), [F]."""
        a, b = self._ptr(v, self.F), self._ptr(eta, self.N)
        out = np.empty(self.F, dtype=np.float64)
        check(_lib.load().odis_op_update_momentum(self._h, a[1], b[1], out.ctypes.data))
        return out

    def updateEta(self, v) -> np.ndarray:
        """detadt = h Div v  (updateEta, src/updateEta.cpp:39), [N]."""
        a = self._ptr(v, self.F)
        out = np.empty(self.N, dtype=np.float64)
        check(_lib.load().odis_op_update_eta(self._h, a[1], out.ctypes.data))
        return out

    def forcing(self, time: float) -> np.ndarray:
        """Tidal potential at `time`  (forcing, src/tidalPotentials.cpp:29-328), [N]."""
        out = np.empty(self.N, dtype=np.float64)
        check(_lib.load().odis_op_forcing(self._h, float(time), out.ctypes.data))
        return out

    def integrateAB3scalar(self, solution, dsolution_dt, iter: int):
        """(solution, dsolution_dt[n][3]) after one Adams-Bashforth update  (src/temporalOperators.cpp:17-68); copies."""
        sol = np.array(solution, dtype=np.float64, order="C").ravel()
        hist = np.array(dsolution_dt, dtype=np.float64, order="C")
        if hist.size != sol.size * 3:
            raise ValueError("dsolution_dt must be [n][3]")
        check(_lib.load().odis_op_integrate_ab3_scalar(self._h, sol.ctypes.data, hist.ctypes.data, iter, sol.size))
        return sol, hist.reshape(sol.size, 3)

    def interpolateVelocity(self, v) -> np.ndarray:
        """East/north velocity components at the edges  (src/interpolation.cpp:26-62), [F][2]."""
        a = self._ptr(v, self.F)
        out = np.empty((self.F, 2), dtype=np.float64)
        check(_lib.load().odis_op_interpolate_velocity(self._h, a[1], out.ctypes.data))
        return out

    def updateEnergy(self, v_avg, areas):
        """(e_flux [F], area-mean flux) from the east/north components and the edge areas  (src/energy.cpp:13-62)."""
        a, b = self._ptr(v_avg, self.F * 2), self._ptr(areas, self.F)
        out = np.empty(self.F, dtype=np.float64)
        avg = C.c_double()
        check(_lib.load().odis_op_update_energy(self._h, a[1], b[1], out.ctypes.data, C.byref(avg)))
        return out, avg.value

    def dissipation_series(self, first: int = 0, count: int | None = None) -> np.ndarray:
        if count is None:
            count = self.steps_since_state + 1 - first
        out = np.empty(count, dtype=np.float64)
        check(_lib.load().odis_get_dissipation_series(self._h, first, count, out.ctypes.data))
        return out

    def trim_dissipation_series(self) -> None:
        """Forget the per-step dissipation series before the current step (odis_trim_dissipation_series)."""
        check(_lib.load().odis_trim_dissipation_series(self._h))
        self._iter0 = self.iter

    @property
    def iter(self) -> int:
        i = C.c_int64()
        check(_lib.load().odis_get_iter(self._h, C.byref(i)))
        return i.value

    def footprint(self) -> tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        check(_lib.load().odis_get_footprint(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launches(self) -> int:
        i = C.c_int64()
        check(_lib.load().odis_get_launch_count(self._h, C.byref(i)))
        return i.value

    def synchronize(self) -> None:
        check(_lib.load().odis_synchronize(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "_lib", None) is not None:
            _lib._lib.odis_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Ensemble:
    """M independent runs on one grid advanced together (parameter sweeps; replaces M separate `./ODIS` runs).
    `params_list`: one dict per member, as for Solver; members may differ in g, h, alpha, love_reduct, ecc, obl."""

    def __init__(self, mesh: Mesh, params_list, device: int = 0):
        self.mesh = mesh
        self.M = len(params_list)
        arr = (Params * self.M)()
        for m, params in enumerate(params_list):
            for k, v in params.items():
                if k == "kernel_select":
                    arr[m].reserved[0] = int(v)
                else:
                    setattr(arr[m], k, v)
        self._h = C.c_void_p()
        check(_lib.load().odis_ensemble_create(C.byref(mesh.view), C.cast(arr, C.c_void_p), self.M, device, C.byref(self._h)))
        self.N, self.F = mesh.n_cells, mesh.n_edges
        self._iter0 = 0

    def set_state(self, member: int = -1, v=None, eta=None, dvdt=None, detadt=None, iter: int = 0) -> None:
        k = [Solver._ptr(v, self.F), Solver._ptr(eta, self.N), Solver._ptr(dvdt, self.F * 3), Solver._ptr(detadt, self.N * 3)]
        check(_lib.load().odis_ensemble_set_state(self._h, member, k[0][1], k[1][1], k[2][1], k[3][1], iter))
        self._iter0 = iter

    def enable_self_gravity(self, l_max: int, factor) -> None:
        """Self-gravity / shell-pressure term for every member (see Solver.enable_self_gravity); l_max <= 10. Analysis and
        synthesis of all members run as FP64 tensor-core GEMMs."""
        f = np.ascontiguousarray(factor, dtype=np.float64)
        if f.size != l_max + 1:
            raise ValueError("factor must have l_max + 1 entries")
        check(_lib.load().odis_ensemble_enable_self_gravity(self._h, C.byref(self.mesh.view), l_max, f.ctypes.data))
        self.sh_rows = (l_max + 1) ** 2

    def sh_coefficients(self, member: int) -> np.ndarray:
        out = np.empty(self.sh_rows, dtype=np.float64)
        check(_lib.load().odis_ensemble_get_sh_coefficients(self._h, member, out.ctypes.data))
        return out

    def step(self, nsteps: int = 1) -> None:
        check(_lib.load().odis_ensemble_step(self._h, nsteps))

    def step_timed(self, nsteps: int) -> float:
        ms = C.c_float()
        check(_lib.load().odis_ensemble_step_timed(self._h, nsteps, C.byref(ms)))
        return ms.value

    def field(self, member: int, fid: int) -> np.ndarray:
        shape = tuple(self.F if d == "F" else self.N if d == "N" else d for d in _FIELD_SHAPES[fid])
        out = np.empty(shape, dtype=np.float64)
        check(_lib.load().odis_ensemble_get_field(self._h, member, fid, out.ctypes.data))
        return out

    def info(self) -> dict:
        n, it, la, db, ab = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(_lib.load().odis_ensemble_get_info(self._h, C.byref(n), C.byref(it), C.byref(la), C.byref(db), C.byref(ab)))
        return dict(n_members=n.value, iter=it.value, launches=la.value, device_bytes=db.value, algorithmic_bytes_per_step=ab.value)

    @property
    def iter(self) -> int:
        return self.info()["iter"]

    def dissipation_series(self, member: int, first: int = 0, count: int | None = None) -> np.ndarray:
        if count is None:
            count = self.iter - self._iter0 + 1 - first
        out = np.empty(count, dtype=np.float64)
        check(_lib.load().odis_ensemble_get_dissipation_series(self._h, member, first, count, out.ctypes.data))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) and _lib is not None and getattr(_lib, "_lib", None) is not None:
            _lib._lib.odis_ensemble_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
