"""B200-native LTE time-step engine, drop-in for the GeodesicODIS hot path.

Python here is a thin host-side mirror of the reference's call surface (Globals / Mesh / solveODIS /
ab3Explicit) over the C ABI in include/odis_b200.h; all computation is in libodis_b200.so (C++ host
code + sm_100a CUDA kernels). Nothing in this package imports oracle/.
"""
from .api import (Globals, Mesh, Solver, Ensemble, H5Writer, run, partition_plan, generate_grid, write_grid_file, params_from_globals, quantise_time_step, sh_basis, sh_normal_inverse, nonlinear_tables, analytical_state,
                  FIELD_VELOCITY, FIELD_ETA, FIELD_DVDT, FIELD_DETADT, FIELD_VELOCITY_EN, FIELD_DISSIPATION, FIELD_POTENTIAL)
from ._lib import OdisError

__all__ = ["Globals", "Mesh", "Solver", "Ensemble", "H5Writer", "run", "partition_plan", "generate_grid", "write_grid_file", "params_from_globals", "quantise_time_step", "sh_basis", "sh_normal_inverse", "nonlinear_tables", "analytical_state",
           "OdisError", "FIELD_VELOCITY", "FIELD_ETA", "FIELD_DVDT", "FIELD_DETADT", "FIELD_VELOCITY_EN",
           "FIELD_DISSIPATION", "FIELD_POTENTIAL"]
