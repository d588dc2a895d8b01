// Device kernels of the nonlinear branch of the time step (`advection; true`), sm_100a, FP64, -fmad=false.
//
// What the reference does per step on this branch (all /root/reference/src):
//   updateMomentum.cpp:37-38      dv/dt = -g G eta, then calculateMomentumAdvection (momAdvection.cpp:11-268):
//       :31       relative vorticity at the vertices     zeta = operatorCurl v                       (3 edges per vertex)
//       :39-67    potential vorticity                    q_v = (zeta + f_v) / h_v,  h_v = (sum_j h_j A_j R_vj) / A_v
//       :77-91    q_e = (q_v1 + q_v2)/2,  h_e = (h_n1 + h_n2)/2
//       :100-142  dv/dt += sum_j (w_ej l_e' / d_e) (h_e' v_e') (q_e + q_e')/2   over the 10-point TRiSK stencil
//       :253-267  kinetic energy at the cells from operatorRBFinterp v (interpolation.cpp:116), dv/dt += -G Ekin
//   timeIntegrator.cpp:218-242    forcing, drag + forcing gradient, AB3, forward-Euler add   (as on the linear branch)
//   updateEta.cpp:32-33           d eta/dt = Div flux, flux from interpolateLSQFlux (interpolation.cpp:311-364):
//                                 d2 = operatorDirectionalSecondDeriv (h + eta); third-order upwind-biased edge value of h times v
//   timeIntegrator.cpp:266-269    h_total = h + eta for the next step
// Here: five gather kernels per step in front of / around the same update arithmetic as edge_step / cell_step. Every sum
// runs in the reference's order (CSR columns ascending for the operator products, stencil slot order for the vorticity flux),
// so the fields are bit-identical to the reference's CPU solver given the same operator coefficients.
#pragma once
#include "odis_kernels.cuh"

namespace odis {

// ELL form of a CSR operator in device numbering: entry (slot k, row r) at [k * stride + r]; slots in ascending reference column
// id (the reference's CSR order), id -1 beyond the row's length.
struct Ell {
    int width, stride;
    const int* id;
    const double* w;
};

struct NlTables {
    int n_vertices, n_edges, n_cells;
    Ell curl;                  // [V] rows -> edges                          mesh.cpp:3122-3175
    Ell rbf[3];                // [N] rows (x, y, z components) -> edges     mesh.cpp:2263-2361
    Ell d2[2];                 // [F] rows (inner, outer) -> cells           mesh.cpp:2364-2719
    int vstride;
    const int* vnode;          // [3][vstride] vertex_nodes, reference slot order
    const double* vR;          // [3][vstride] vertex_R
    const double* vsin;        // [V] vertex_sinlat
    const double* varea_r;     // [V] 1 / vertex_area                         mesh.cpp:1147
    const double* carea;       // [N] control_volume_surf_area_map
    const int2* fvert;         // [F] face_vertexes (device vertex ids)
    int estride;
    const int* nid;            // [10][estride] face_interp_friends in the reference's slot order (NOT sorted), -1 beyond friend_num
    const double* ncoef;       // [10][estride] face_interp_weights * face_len(friend) * face_node_dist_r
    const int2* cells;         // [F] inner, outer cell
    const double2* grad;       // [F] gradient coefficients (as EdgeTables::grad)
    const double* dist;        // [F] face_node_dist
    const int* eid;            // [6][cstride] cell -> edges with the outer-cell flag in bit 31 (as CellTables::eid)
    const double* area;        // == carea
    int cstride;
    double omega;              // rotation rate (vorticity of the frame: f_v = -2 omega sin(lat_v))
};

struct NlState {
    const double2* vl_in;      // [F] {v^n, l_e}
    double2* vl_out;           // [F] {v^{n+1}, l_e}
    const double2* eu_in;      // [N] {eta^n, U(t_n + dt)}
    double2* eu_out;           // [N] {eta^{n+1}, U left for the potential pass}
    double* h1; double* h2;    // edge tendency history (as EdgeState)
    const double* ch1; const double* ch2; double* chw;   // cell tendency history (as CellState)
    double* qv;                // [V] scratch: potential vorticity at the vertices
    double2* fq;               // [F] scratch: {h_e v_e, q_e}
    double* ekin;              // [N] scratch
    double* flux;              // [F] scratch
    // folded step only: the dissipated-energy sum of v^n leaves the edge update (per-block partials, last-block ticket, result)
    double* block_partial;     // [>= ceil(F / 128)]
    unsigned int* ticket;      // zero before the launch, left zero
    double* energy_out;
};

// launches of one nonlinear step (before the potential pass and the diagnostics): vertex PV, edge {F_e, q_e}, cell Ekin,
// edge update, flux, cell update
constexpr int kNlLaunches = 6;
void launch_step_nonlinear(const NlTables& t, const Physics& p, const NlState& s, int mode, cudaStream_t stream);
// Default: the same arithmetic in 4 launches — vertex PV + cell Ekin in one grid; edge {F_e, q_e}; edge update + thickness flux; cell
// update — with the two passes AROUND the step folded in: the dissipated energy of v^n is summed inside the edge update
// (instead of edge_diag_kernel before the step) and the cell update evaluates the next step's tidal potential itself (instead of a
// cell_step_kernel pass after it): 4 launches per step instead of 6. Fields bit-identical; the energy sum within ~1e-15 (bar 1e-12).
// ct: the trig rows of the potential; next: its time factors (forcing(current_time + dt) of the next step).
constexpr int kNlLaunchesFolded = 4;
void launch_step_nonlinear_folded(const NlTables& t, const Physics& p, const NlState& s, int mode, const CellTables& ct, const StepScalars& next,
                                  cudaStream_t stream);

}  // namespace odis
