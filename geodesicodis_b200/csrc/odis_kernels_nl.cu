// Kernels of the nonlinear branch: see odis_kernels_nl.cuh.
#include "odis_kernels_nl.cuh"

namespace odis {
namespace {

constexpr int kNlThreads = 128;

__device__ __forceinline__ double exact_div(double x, double d, double y) {      // as in odis_kernels.cu: IEEE x / d given y = rcp(d)
    const double q0 = __dmul_rn(x, y);
    const double r0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, x);
    return __fma_rn(r1, y, q1);
}

__device__ __forceinline__ double ab3_increment(double f0, double f1, double f2, double dt, int mode) {   // temporalOperators.cpp:41-43,55,64
    const double a = 23. / 12., b = -16. / 12., c = 5. / 12.;
    if (mode == AB3_FULL) return (a * f0 + b * f1 + c * f2) * dt;
    return f0 * dt;
}

// row r of an ELL operator times a scalar field read through `get(id)`: columns ascending, accumulator from 0 (Eigen's row-major product)
template <typename Get>
__device__ __forceinline__ double ell_row(const Ell& A, int r, Get get) {
    double tmp = 0.0;
    for (int k = 0; k < A.width; k++) {
        const int id = A.id[(size_t)k * A.stride + r];
        if (id >= 0) tmp += A.w[(size_t)k * A.stride + r] * get(id);
    }
    return tmp;
}

// momAdvection.cpp:31-67: q_v = (curl v + f_v) / h_v
__global__ void __launch_bounds__(kNlThreads) nl_vertex_kernel(NlTables t, Physics p, NlState s) {
    const int i = blockIdx.x * kNlThreads + threadIdx.x;
    if (i >= t.n_vertices) return;
    const double zeta = ell_row(t.curl, i, [&](int e) { return s.vl_in[e].x; });
    const double f = -2 * t.omega * t.vsin[i];
    double thickness = 0.0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int node = t.vnode[(size_t)j * t.vstride + i];
        const double hn = p.h + s.eu_in[node].x;                                   // h_total, timeIntegrator.cpp:203,268
        thickness += hn * t.carea[node] * t.vR[(size_t)j * t.vstride + i];
    }
    thickness *= t.varea_r[i];
    s.qv[i] = (zeta + f) / thickness;
}

// momAdvection.cpp:77-91 and the mass flux h_e v_e of :126
__global__ void __launch_bounds__(kNlThreads) nl_edge_prep_kernel(NlTables t, Physics p, NlState s) {
    const int e = blockIdx.x * kNlThreads + threadIdx.x;
    if (e >= t.n_edges) return;
    const int2 v = t.fvert[e], c = t.cells[e];
    const double q = 0.5 * (s.qv[v.x] + s.qv[v.y]);
    const double he = 0.5 * ((p.h + s.eu_in[c.x].x) + (p.h + s.eu_in[c.y].x));
    s.fq[e] = make_double2(he * s.vl_in[e].x, q);
}

// interpolation.cpp:116 + momAdvection.cpp:257-259
__global__ void __launch_bounds__(kNlThreads) nl_cell_ekin_kernel(NlTables t, NlState s) {
    const int i = blockIdx.x * kNlThreads + threadIdx.x;
    if (i >= t.n_cells) return;
    auto vel = [&](int e) { return s.vl_in[e].x; };
    const double x = ell_row(t.rbf[0], i, vel), y = ell_row(t.rbf[1], i, vel), z = ell_row(t.rbf[2], i, vel);
    s.ekin[i] = 0.5 * (x * x + y * y + z * z);
}

// updateMomentum.cpp:37-38 with momAdvection.cpp:100-142,267, then timeIntegrator.cpp:215-242 as in edge_step_kernel
__global__ void __launch_bounds__(kNlThreads) nl_edge_step_kernel(NlTables t, Physics p, NlState s, int mode) {
    const int e = blockIdx.x * kNlThreads + threadIdx.x;
    if (e >= t.n_edges) return;
    const int2 c = t.cells[e];
    const double2 G = t.grad[e];
    const double2 own = s.vl_in[e];
    const double2 in = s.eu_in[c.x], out = s.eu_in[c.y];
    double dv = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;                       // dvdt = -g G eta
    const double q_e = s.fq[e].y;
    double F_tang_q = 0.0;
#pragma unroll
    for (int j = 0; j < kStencil; j++) {
        const int f = t.nid[(size_t)j * t.estride + e];
        if (f >= 0) {
            const double2 o = s.fq[f];                                            // {F_e', q_e'}
            F_tang_q += t.ncoef[(size_t)j * t.estride + e] * o.x * (q_e + o.y) * 0.5;
        }
    }
    dv -= -F_tang_q;
    dv += (-G.x) * s.ekin[c.x] + (-G.y) * s.ekin[c.y];                            // dvdt += -G Ekin
    const double f0 = dv;
    const double drag = (-p.alpha) * own.x + (G.x * in.y + G.y * out.y);          // timeIntegrator.cpp:219
    double v = own.x + ab3_increment(f0, s.h1[e], s.h2[e], p.dt, mode);
    v += p.dt * drag;                                                             // timeIntegrator.cpp:242
    s.vl_out[e] = make_double2(v, own.y);
    if (mode == AB3_SECOND) s.h1[e] = f0;
    else s.h2[e] = f0;
}

// interpolateLSQFlux, interpolation.cpp:311-364
__global__ void __launch_bounds__(kNlThreads) nl_flux_kernel(NlTables t, Physics p, NlState s) {
    const int e = blockIdx.x * kNlThreads + threadIdx.x;
    if (e >= t.n_edges) return;
    auto htot = [&](int i) { return p.h + s.eu_in[i].x; };
    const double d2_inner = ell_row(t.d2[0], e, htot), d2_outer = ell_row(t.d2[1], e, htot);
    const int2 c = t.cells[e];
    const double fact = 1. / 12.0, beta = 1.0;
    const double dx = t.dist[e];
    const double dx2 = dx * dx * fact;
    const double vel = s.vl_out[e].x;
    s.flux[e] = vel * 0.5 * (htot(c.x) + htot(c.y)) - dx2 * (d2_outer + d2_inner) * vel + dx2 * beta * fabs(vel) * (d2_outer - d2_inner);
}

// updateEta.cpp:33 (d eta/dt = Div flux) + integrateAB3scalar; the potential of the next step is left to the potential pass
__global__ void __launch_bounds__(kNlThreads) nl_cell_step_kernel(NlTables t, Physics p, NlState s, int mode) {
    const int i = blockIdx.x * kNlThreads + threadIdx.x;
    if (i >= t.n_cells) return;
    double2 st = s.eu_in[i];
    const double area = t.area[i], ra = __drcp_rn(area);
    double div = 0.0;
#pragma unroll
    for (int j = 0; j < kCellEdges; j++) {
        const int packed = t.eid[(size_t)j * t.cstride + i];
        if (packed != -1) {
            const int e = packed & 0x7fffffff;
            const double ndir = (packed < 0) ? 1.0 : -1.0;                        // -dir, mesh.cpp:3246
            const double coeff = exact_div(ndir * s.vl_out[e].y, area, ra);
            div += coeff * s.flux[e];
        }
    }
    st.x += ab3_increment(div, s.ch1[i], s.ch2[i], p.dt, mode);
    s.chw[i] = div;
    s.eu_out[i] = st;
}

// ---- 4-launch variant (opt-in, odis_params.reserved[0] bit 5): the same arithmetic in fewer passes ----
// (a) vertex potential vorticity and cell kinetic energy have no dependency on each other: one grid, the first blocks take the
//     vertices, the rest the cells;
// (b) the third-order thickness flux of an edge needs only that edge's own new velocity (in a register at the end of the edge
//     update) and second derivatives of h + eta^n: computed there instead of in a pass of its own, which also saves re-reading
//     {v^{n+1}, l} and the cell ids.
// The bodies repeat nl_vertex / nl_cell_ekin / nl_edge_step / nl_flux statement for statement (same order of operations).
__global__ void __launch_bounds__(kNlThreads) nl_vertex_ekin_kernel(NlTables t, Physics p, NlState s, int vertex_blocks) {
    if ((int)blockIdx.x < vertex_blocks) {
        const int i = blockIdx.x * kNlThreads + threadIdx.x;
        if (i >= t.n_vertices) return;
        const double zeta = ell_row(t.curl, i, [&](int e) { return s.vl_in[e].x; });
        const double f = -2 * t.omega * t.vsin[i];
        double thickness = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int node = t.vnode[(size_t)j * t.vstride + i];
            const double hn = p.h + s.eu_in[node].x;
            thickness += hn * t.carea[node] * t.vR[(size_t)j * t.vstride + i];
        }
        thickness *= t.varea_r[i];
        s.qv[i] = (zeta + f) / thickness;
    } else {
        const int i = ((int)blockIdx.x - vertex_blocks) * kNlThreads + threadIdx.x;
        if (i >= t.n_cells) return;
        auto vel = [&](int e) { return s.vl_in[e].x; };
        const double x = ell_row(t.rbf[0], i, vel), y = ell_row(t.rbf[1], i, vel), z = ell_row(t.rbf[2], i, vel);
        s.ekin[i] = 0.5 * (x * x + y * y + z * z);
    }
}

__global__ void __launch_bounds__(kNlThreads) nl_edge_step_flux_kernel(NlTables t, Physics p, NlState s, int mode) {
    const int e = blockIdx.x * kNlThreads + threadIdx.x;
    if (e >= t.n_edges) return;
    const int2 c = t.cells[e];
    const double2 G = t.grad[e];
    const double2 own = s.vl_in[e];
    const double2 in = s.eu_in[c.x], out = s.eu_in[c.y];
    double dv = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;
    const double q_e = s.fq[e].y;
    double F_tang_q = 0.0;
#pragma unroll
    for (int j = 0; j < kStencil; j++) {
        const int f = t.nid[(size_t)j * t.estride + e];
        if (f >= 0) {
            const double2 o = s.fq[f];
            F_tang_q += t.ncoef[(size_t)j * t.estride + e] * o.x * (q_e + o.y) * 0.5;
        }
    }
    dv -= -F_tang_q;
    dv += (-G.x) * s.ekin[c.x] + (-G.y) * s.ekin[c.y];
    const double f0 = dv;
    const double drag = (-p.alpha) * own.x + (G.x * in.y + G.y * out.y);
    double v = own.x + ab3_increment(f0, s.h1[e], s.h2[e], p.dt, mode);
    v += p.dt * drag;
    s.vl_out[e] = make_double2(v, own.y);
    if (mode == AB3_SECOND) s.h1[e] = f0;
    else s.h2[e] = f0;
    // interpolateLSQFlux (interpolation.cpp:311-364) for this edge, with the velocity just computed
    auto htot = [&](int i) { return p.h + s.eu_in[i].x; };
    const double d2_inner = ell_row(t.d2[0], e, htot), d2_outer = ell_row(t.d2[1], e, htot);
    const double fact = 1. / 12.0, beta = 1.0;
    const double dx = t.dist[e];
    const double dx2 = dx * dx * fact;
    const double vel = v;
    s.flux[e] = vel * 0.5 * (htot(c.x) + htot(c.y)) - dx2 * (d2_outer + d2_inner) * vel + dx2 * beta * fabs(vel) * (d2_outer - d2_inner);
}

}  // namespace

void launch_step_nonlinear(const NlTables& t, const Physics& p, const NlState& s, int mode, cudaStream_t stream) {
    auto grid = [](int n) { return (unsigned)((n + kNlThreads - 1) / kNlThreads); };
    nl_vertex_kernel<<<grid(t.n_vertices), kNlThreads, 0, stream>>>(t, p, s);
    nl_edge_prep_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_cell_ekin_kernel<<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, s);
    nl_edge_step_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s, mode);
    nl_flux_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_cell_step_kernel<<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, mode);
}

void launch_step_nonlinear_fused(const NlTables& t, const Physics& p, const NlState& s, int mode, cudaStream_t stream) {
    auto grid = [](int n) { return (unsigned)((n + kNlThreads - 1) / kNlThreads); };
    const unsigned vb = grid(t.n_vertices);
    nl_vertex_ekin_kernel<<<vb + grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, (int)vb);
    nl_edge_prep_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_edge_step_flux_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s, mode);
    nl_cell_step_kernel<<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, mode);
}

}  // namespace odis
