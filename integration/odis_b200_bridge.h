// Reference-side glue between GeodesicODIS's own types (Globals, Mesh, Array1D/2D) and the C ABI of
// libodis_b200.so (include/odis_b200.h). These files are compiled INSIDE a GeodesicODIS build, against
// the reference's headers — they are what a reference maintainer adds to the tree; nothing here is part
// of the library. Two levels:
//
//   integration/timeIntegrator_b200.cpp   replaces src/timeIntegrator.cpp: ab3Explicit drives the fused
//                                         device loop (odis_step) and keeps the reference's dumps / log /
//                                         restart files;
//   integration/operators_b200.cpp        keeps the reference's ab3Explicit and replaces the free
//                                         functions its loop calls (updateMomentum, updateEta, forcing,
//                                         integrateAB3scalar, interpolateVelocity, updateEnergy) one by one.
//
// Both share one device solver per process, created from the reference's Mesh tables at first use.
#pragma once

#include "globals.h"
#include "mesh.h"

#include "odis_b200.h"

#include <cstddef>
#include <cstdint>

namespace odis_bridge {

// Borrowed view of the reference's Mesh tables (Array2D is row-major, &a(0,0) is the flat pointer).
odis_mesh_view mesh_view(Globals* globals, Mesh* grid);
// Scalars as the reference holds them after Globals / applySurfaceBCs / CalcMaxTimeStep have run.
odis_params params(Globals* globals);
// The process-wide device solver for (globals, grid); created on first call (nonlinear operators handed over when
// `advection; true`); grid == nullptr returns the solver already created for `globals` (updateEnergy has no Mesh argument).
// Any failure goes through Output->Write(ERR_MESSAGE) + TerminateODIS like the reference's own
// fatal paths (src/outFiles.cpp:123-130) — there is no CPU fallback.
odis_solver* solver(Globals* globals, Mesh* grid);
// rc != ODIS_OK: report `what` + odis_last_error() through the reference's error channel and terminate.
void check(Globals* globals, int rc, const char* what);
void release();

// Several GPUs from the reference's one process (the reference has no such notion; the environment variable ODIS_B200_GPUS=N is the
// switch, input.in stays the reference's): the grid is cut into N parts, one partitioned solver per GPU (devices 0..N-1), halos
// exchanged by the step kernels through peer memory. `group()` creates them on first use (N = 1: the one solver() above); the
// helpers below issue a call on every rank and assemble whole-grid results — the ranks fill their own entries of a field and leave
// zeros elsewhere, the dissipation is the sum of the ranks' shares. Linear branch only (`advection; false`) when N > 1.
struct Group {
    int world = 1;
    odis_solver* rank[8] = {nullptr};
};
const Group& group(Globals* globals, Mesh* grid);
void set_state_all(Globals* globals, const Group& g, const double* v, const double* eta, const double* dvdt, const double* detadt, int64_t iter);
void step_all(Globals* globals, const Group& g, int32_t nsteps);              // enqueued on every rank, then every rank synchronised
void get_field_all(Globals* globals, const Group& g, int32_t field, double* out, size_t count);
double dissipation_avg_all(Globals* globals, const Group& g);

}  // namespace odis_bridge
