"""ctypes binding of the C ABI declared in include/odis_b200.h (libodis_b200.so, built in-tree by
geodesicodis_b200/build.py). There is no fallback: if the library is missing, loading raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ODIS_B200_LIB points at an alternative build of the same library (kernel tuning experiments)
# ODIS_B200_HOST_ONLY=1: the host-only part of the ABI (libodis_b200_host.so: grid generator, mesh tables, configuration; no CUDA, no
# solver) — what bench.py's reference arm uses to write the reference's input files without mapping the CUDA library
HOST_ONLY = os.environ.get("ODIS_B200_HOST_ONLY", "") not in ("", "0")
LIB_PATH = os.environ.get("ODIS_B200_LIB") or os.path.join(HERE, "libodis_b200_host.so" if HOST_ONLY else "libodis_b200.so")

c_i32, c_i64, c_f64 = C.c_int32, C.c_int64, C.c_double
P = C.POINTER

MESH_VIEW_ARRAYS = [  # name, dtype code, (dim symbol, columns)
    ("node_pos_sph", "d", ("N", 2)), ("node_friends", "i", ("N", 6)), ("centroid_pos_sph", "d", ("N", 12)),
    ("control_volume_surf_area_map", "d", ("N", 1)), ("faces", "i", ("N", 6)), ("node_face_dir", "i", ("N", 6)),
    ("vertexes", "i", ("N", 6)), ("face_nodes", "i", ("F", 2)), ("face_vertexes", "i", ("F", 2)),
    ("face_interp_friends", "i", ("F", 10)), ("face_interp_weights", "d", ("F", 10)), ("face_len", "d", ("F", 1)),
    ("face_node_dist", "d", ("F", 1)), ("face_centre_m", "d", ("F", 2)), ("face_centre_pos_sph", "d", ("F", 2)),
    ("face_intercept_pos_sph", "d", ("F", 2)), ("face_area", "d", ("F", 1)), ("face_normal_vec_map", "d", ("F", 2)),
    ("vertex_pos_sph", "d", ("V", 2)), ("vertex_nodes", "i", ("V", 3)), ("vertex_R", "d", ("V", 3)),
]


class MeshView(C.Structure):
    _fields_ = [("n_cells", c_i32), ("n_edges", c_i32), ("n_vertices", c_i32), ("radius", c_f64)] + \
               [(name, C.c_void_p) for name, _, _ in MESH_VIEW_ARRAYS]


class Params(C.Structure):
    _fields_ = [(n, c_f64) for n in ("g", "h", "alpha", "dt", "radius", "omega", "love_reduct", "ecc", "obl",
                                     "shell_thickness", "semimajor_axis")] + \
               [(n, c_i32) for n in ("potential", "friction", "surface", "init_load", "reorder", "block_threads")] + \
               [("reserved", c_i32 * 4)]


class CsrView(C.Structure):
    _fields_ = [("n_rows", c_i32), ("n_cols", c_i32), ("indptr", C.c_void_p), ("indices", C.c_void_p), ("data", C.c_void_p)]


class NonlinearView(C.Structure):
    _fields_ = [("curl", CsrView), ("rbf_interp", CsrView), ("directional_second_deriv", CsrView), ("vertex_sinlat", C.c_void_p),
                ("vertex_area", C.c_void_p)]


class PartitionPlan(C.Structure):
    _fields_ = [(n, c_i32) for n in ("rank", "world", "own_cells", "own_edges", "local_cells", "local_edges", "n_peers", "reserved")] + \
               [(n, P(c_i32)) for n in ("local_cell_ref", "local_edge_ref", "peer_rank", "peer_counts", "send_edge_ref", "send_edge_slot",
                                        "send_cell_ref", "send_cell_slot")]


class RunOptions(C.Structure):
    _fields_ = [("device", c_i32), ("reorder", c_i32), ("echo", c_i32), ("self_gravity", c_i32), ("max_steps", c_i64),
                ("overlap_output", c_i32), ("n_gpus", c_i32)]


class SnapshotView(C.Structure):
    _fields_ = [("eta", C.c_void_p), ("velocity_en", C.c_void_p), ("dissipation", C.c_void_p), ("velocity", C.c_void_p),
                ("dissipation_avg", c_f64), ("iter", c_i64)]


class RunResult(C.Structure):
    _fields_ = [("steps", c_i64), ("kernel_launches", c_i64), ("dumps", c_i32), ("interrupted", c_i32), ("n_cells", c_i32),
                ("n_edges", c_i32), ("steps_per_period", c_i32), ("reserved", c_i32), ("dt", c_f64), ("last_dissipation_avg", c_f64)]


# every symbol include/odis_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "odis_last_error": (C.c_char_p, []),
    "odis_version": (C.c_char_p, []),
    "odis_config_create": (C.c_int, [P(C.c_void_p)]),
    "odis_config_load": (C.c_int, [C.c_char_p, P(C.c_void_p)]),
    "odis_config_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "odis_config_finalize": (C.c_int, [C.c_void_p]),
    "odis_config_get_double": (C.c_int, [C.c_void_p, C.c_char_p, P(c_f64)]),
    "odis_config_get_int": (C.c_int, [C.c_void_p, C.c_char_p, P(c_i32)]),
    "odis_config_get_bool": (C.c_int, [C.c_void_p, C.c_char_p, P(c_i32)]),
    "odis_config_get_string": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, c_i32]),
    "odis_config_get_enum": (C.c_int, [C.c_void_p, c_i32, P(c_i32)]),
    "odis_config_free": (None, [C.c_void_p]),
    "odis_quantise_time_step": (C.c_int, [c_f64, c_f64, P(c_f64), P(c_i32)]),
    "odis_mesh_from_file": (C.c_int, [C.c_char_p, c_f64, c_i32, P(C.c_void_p)]),
    "odis_mesh_from_arrays": (C.c_int, [c_i32, C.c_void_p, C.c_void_p, C.c_void_p, c_f64, c_i32, P(C.c_void_p)]),
    "odis_mesh_get_view": (C.c_int, [C.c_void_p, P(MeshView)]),
    "odis_mesh_free": (None, [C.c_void_p]),
    "odis_grid_generate": (C.c_int, [c_i32, P(c_i32), P(P(c_f64)), P(P(c_i32)), P(P(c_f64))]),
    "odis_grid_write_file": (C.c_int, [C.c_char_p, c_i32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_free": (None, [C.c_void_p]),
    "odis_create": (C.c_int, [P(MeshView), P(Params), c_i32, P(C.c_void_p)]),
    "odis_create_partitioned": (C.c_int, [P(MeshView), P(Params), c_i32, c_i32, c_i32, P(C.c_void_p)]),
    "odis_halo_blob_size": (C.c_int, []),
    "odis_halo_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "odis_halo_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "odis_get_partition": (C.c_int, [C.c_void_p] + [P(c_i32)] * 7),
    "odis_get_partition_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_partition_plan": (C.c_int, [P(MeshView), c_i32, c_i32, c_i32, P(PartitionPlan)]),
    "odis_partition_plan_free": (None, [P(PartitionPlan)]),
    "odis_analytical_state": (C.c_int, [P(MeshView), P(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64]),
    "odis_step": (C.c_int, [C.c_void_p, c_i32]),
    "odis_step_timed": (C.c_int, [C.c_void_p, c_i32, P(C.c_float)]),
    "odis_step_profiled": (C.c_int, [C.c_void_p, c_i32, P(C.c_float), P(C.c_float)]),
    "odis_nonlinear_create": (C.c_int, [C.c_void_p, c_f64, P(C.c_void_p)]),
    "odis_nonlinear_get_view": (C.c_int, [C.c_void_p, P(NonlinearView)]),
    "odis_nonlinear_free": (None, [C.c_void_p]),
    "odis_enable_advection": (C.c_int, [C.c_void_p, P(MeshView), P(NonlinearView)]),
    "odis_enable_self_gravity": (C.c_int, [C.c_void_p, P(MeshView), c_i32, C.c_void_p, c_i32]),
    "odis_get_sh_coefficients": (C.c_int, [C.c_void_p, C.c_void_p]),
    "odis_step_profiled_sh": (C.c_int, [C.c_void_p, c_i32, P(C.c_float), P(C.c_float), P(C.c_float)]),
    "odis_sh_basis": (C.c_int, [c_i32, C.c_void_p, c_i32, C.c_void_p]),
    "odis_sh_normal_inverse": (C.c_int, [c_i32, C.c_void_p, c_i32, C.c_void_p]),
    "odis_get_field": (C.c_int, [C.c_void_p, c_i32, C.c_void_p]),
    "odis_get_dissipation_avg": (C.c_int, [C.c_void_p, P(c_f64)]),
    "odis_get_dissipation_series": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p]),
    "odis_trim_dissipation_series": (C.c_int, [C.c_void_p]),
    "odis_op_update_momentum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_op_update_eta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_op_forcing": (C.c_int, [C.c_void_p, c_f64, C.c_void_p]),
    "odis_op_integrate_ab3_scalar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i32]),
    "odis_op_interpolate_velocity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_op_update_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(c_f64)]),
    "odis_stage_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "odis_commit_state": (C.c_int, [C.c_void_p, c_i64]),
    "odis_snapshot_begin": (C.c_int, [C.c_void_p, c_i32, C.c_uint32]),
    "odis_snapshot_wait": (C.c_int, [C.c_void_p, c_i32, P(SnapshotView)]),
    "odis_get_iter": (C.c_int, [C.c_void_p, P(c_i64)]),
    "odis_get_footprint": (C.c_int, [C.c_void_p, P(c_i64), P(c_i64)]),
    "odis_get_launch_count": (C.c_int, [C.c_void_p, P(c_i64)]),
    "odis_synchronize": (C.c_int, [C.c_void_p]),
    "odis_destroy": (None, [C.c_void_p]),
    "odis_ensemble_create": (C.c_int, [P(MeshView), C.c_void_p, c_i32, c_i32, P(C.c_void_p)]),
    "odis_ensemble_set_state": (C.c_int, [C.c_void_p, c_i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64]),
    "odis_ensemble_enable_self_gravity": (C.c_int, [C.c_void_p, P(MeshView), c_i32, C.c_void_p]),
    "odis_ensemble_get_sh_coefficients": (C.c_int, [C.c_void_p, c_i32, C.c_void_p]),
    "odis_ensemble_step": (C.c_int, [C.c_void_p, c_i32]),
    "odis_ensemble_step_timed": (C.c_int, [C.c_void_p, c_i32, P(C.c_float)]),
    "odis_ensemble_get_field": (C.c_int, [C.c_void_p, c_i32, c_i32, C.c_void_p]),
    "odis_ensemble_get_dissipation_series": (C.c_int, [C.c_void_p, c_i32, c_i64, c_i64, C.c_void_p]),
    "odis_ensemble_get_info": (C.c_int, [C.c_void_p, P(c_i32), P(c_i64), P(c_i64), P(c_i64), P(c_i64)]),
    "odis_ensemble_destroy": (None, [C.c_void_p]),
    "odis_h5_create": (C.c_int, [C.c_char_p, P(C.c_void_p)]),
    "odis_h5_add_dataset": (C.c_int, [C.c_void_p, C.c_char_p, c_i32, P(C.c_uint64), P(c_i32)]),
    "odis_h5_write_rows": (C.c_int, [C.c_void_p, c_i32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "odis_h5_close": (C.c_int, [C.c_void_p]),
    "odis_run": (C.c_int, [C.c_char_p, P(RunOptions), P(RunResult)]),
}

_lib = None


class OdisError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"odis_b200 error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -m geodesicodis_b200.build "
                "or __graft_entry__.build()). geodesicodis_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if HOST_ONLY and not hasattr(lib, name):
                continue                 # solver / run entry points live in the CUDA library only
            fn = getattr(lib, name)      # AttributeError here means the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise OdisError(rc, load().odis_last_error().decode(errors="replace"))
