// Analytical solution of the linear LTE used as an initial state (`initial conditions; ANALYTICAL`): host only.
// Replaces analyticalInitialConditions (/root/reference/src/initialConditions.cpp:146-208) and analyticalLTE
// (/root/reference/src/analyticalLTE.cpp:47-181), which the reference implements for the westward obliquity tide only.
#pragma once

namespace odis {

struct AnalyticParams {
    double radius, omega, g, h, alpha, obl, dt;
};

// State and AB3 history of the analytical OBLIQ_WEST response at t = 0 (history levels at 0, -dt, -2dt), sampled at the edge
// midpoints (projected on the edge normals) and at the cell centres. Arrays in reference numbering and layouts:
// v[F], dvdt[F][3], eta[N], detadt[N][3].
void analytical_state_obliq_west(const AnalyticParams& p, int n_cells, int n_edges, const double* node_pos_sph /*[N][2]*/,
                                 const double* face_centre_pos_sph /*[F][2]*/, const double* face_normal_vec_map /*[F][2]*/, double* v, double* dvdt,
                                 double* eta, double* detadt);

}  // namespace odis
