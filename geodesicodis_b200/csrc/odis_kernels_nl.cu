// Kernels of the nonlinear branch: see odis_kernels_nl.cuh.
#include "odis_kernels_nl.cuh"
#include "odis_potential.cuh"

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
// The loads are batched: all ids and weights of the row first, then all gathers, then the sum in slot order -- one DRAM round trip and
// one gather round trip per row instead of one of each per slot (the per-slot loop kept every load behind the previous slot's branch).
// Same products, same order of additions.
// Table values (ids, weights: streamed once) are read with ld.volatile: neither nvcc nor ptxas may reorder volatile accesses among
// themselves, so all of a row's table loads are issued back to back before the first gather has to wait for its id. (With plain loads
// ptxas sinks each slot's loads next to their use to save registers -- 32 registers, and ten serial DRAM + gather round trips per edge.)
// Gathers of field values stay ordinary cached loads.
__device__ __forceinline__ int ldt(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldt(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldo(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldo(const double2* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldo_x(const double2* p) { return ldo(reinterpret_cast<const double*>(p)); }
__device__ __forceinline__ double ldo_y(const double2* p) { return ldo(reinterpret_cast<const double*>(p) + 1); }

template <int W, typename Get>
__device__ __forceinline__ double ell_row_batched(const Ell& A, int r, Get get) {
    int id[W];
    double w[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        const bool in = k < A.width;
        id[k] = in ? ldt(A.id + (size_t)k * A.stride + r) : -1;
        w[k] = in ? ldt(A.w + (size_t)k * A.stride + r) : 0.0;
    }
    // z is 0 (ids are >= -1, so their AND is never 0x80000001), but only at run time: every gather address depends on every table
    // value, so all table loads are issued before the first gather -- ptxas otherwise interleaves load, gather and arithmetic slot by slot
    int all_ids = id[0], all_hi = __double2hiint(w[0]);
#pragma unroll
    for (int k = 1; k < W; k++) { all_ids &= id[k]; all_hi &= __double2hiint(w[k]); }
    const int z = (int)(all_ids == (int)0x80000001) & (int)(all_hi == (int)0x80000001);
    double x[W];
#pragma unroll
    for (int k = 0; k < W; k++) x[k] = get((id[k] >= 0 ? id[k] : 0) + z);
    double tmp = 0.0;
#pragma unroll
    for (int k = 0; k < W; k++)
        if (id[k] >= 0) tmp += w[k] * x[k];
    return tmp;
}
template <typename Get>
__device__ __forceinline__ double ell_row(const Ell& A, int r, Get get) {
    if (A.width <= 3) return ell_row_batched<3>(A, r, get);       // curl
    if (A.width <= 6) return ell_row_batched<6>(A, r, get);       // RBF reconstruction
    if (A.width <= 8) return ell_row_batched<8>(A, r, get);       // directional second derivative (<= 7)
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
    const double zeta = ell_row(t.curl, i, [&](int e) { return ldo_x(s.vl_in + e); });
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
    auto vel = [&](int e) { return ldo_x(s.vl_in + e); };
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
    int nf[kStencil];
    double nc[kStencil];
    double2 no[kStencil];
#pragma unroll
    for (int j = 0; j < kStencil; j++) {                                           // table values, then gathers, then the sum in slot order
        nf[j] = ldt(t.nid + (size_t)j * t.estride + e);
        nc[j] = ldt(t.ncoef + (size_t)j * t.estride + e);
    }
    // every gather address depends on every table value (z is 0: ids are >= -1 and never all 0x80000001 ... but only at run time), so
    // the table loads have to be issued, all of them, before the first gather -- ptxas otherwise interleaves slot by slot
    int all_ids = nf[0], all_hi = __double2hiint(nc[0]);
#pragma unroll
    for (int j = 1; j < kStencil; j++) { all_ids &= nf[j]; all_hi &= __double2hiint(nc[j]); }
    const int z = (int)(all_ids == (int)0x80000001) & (int)(all_hi == (int)0x80000001);
#pragma unroll
    for (int j = 0; j < kStencil; j++) no[j] = ldo(s.fq + ((nf[j] >= 0 ? nf[j] : e) + z));   // {F_e', q_e'}
#pragma unroll
    for (int j = 0; j < kStencil; j++) {
        if (nf[j] >= 0) {
            const double2 o = no[j];
            F_tang_q += nc[j] * o.x * (q_e + o.y) * 0.5;
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
    auto htot = [&](int i) { return p.h + ldo_x(s.eu_in + i); };
    const double d2_inner = ell_row(t.d2[0], e, htot), d2_outer = ell_row(t.d2[1], e, htot);
    const int2 c = t.cells[e];
    const double fact = 1. / 12.0, beta = 1.0;
    const double dx = t.dist[e];
    const double dx2 = dx * dx * fact;
    const double vel = s.vl_out[e].x;
    s.flux[e] = vel * 0.5 * (htot(c.x) + htot(c.y)) - dx2 * (d2_outer + d2_inner) * vel + dx2 * beta * fabs(vel) * (d2_outer - d2_inner);
}

// updateEta.cpp:33 (d eta/dt = Div flux) + integrateAB3scalar. kPot = false: the potential of the next step is left to the potential
// pass (a launch of cell_step_kernel); kPot = true: evaluated here from the same table rows with the same expression (tidal_potential,
// odis_potential.cuh), which saves that launch and a second pass over {eta,U}
template <bool kPot>
__global__ void __launch_bounds__(kNlThreads) nl_cell_step_kernel(NlTables t, Physics p, NlState s, int mode, CellTables ct, StepScalars next) {
    const int i = blockIdx.x * kNlThreads + threadIdx.x;
    if (i >= t.n_cells) return;
    double2 st = s.eu_in[i];
    const double area = t.area[i], ra = __drcp_rn(area);
    double div = 0.0;
    int packed[kCellEdges];
    double le[kCellEdges], fl[kCellEdges];
#pragma unroll
    for (int j = 0; j < kCellEdges; j++) packed[j] = ldt(t.eid + (size_t)j * t.cstride + i);
    TrigValues tv = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (kPot) {                                                                   // the rows the potential reads (as load_trig, odis_kernels.cu)
        const size_t N = (size_t)ct.n_cells;
        const double* T = ct.trig;
        switch (p.potential) {
            case P_ECC:
                tv.cosSq = ldt(ct.trig_sq + i); tv.sinSq = ldt(ct.trig_sq + N + i); tv.cos2Lon = ldt(T + 6 * N + i); tv.sin2Lon = ldt(T + 7 * N + i);
                break;
            case P_OBLIQ:
                tv.sin2Lat = ldt(T + 5 * N + i); tv.cosLon = ldt(T + 2 * N + i);
                break;
            case P_OBLIQ_WEST:
                tv.cosLat = ldt(T + i); tv.sinLat = ldt(T + N + i); tv.cosLon = ldt(T + 2 * N + i); tv.sinLon = ldt(T + 3 * N + i);
                break;
            case P_FULL:
                tv.cosSq = ldt(ct.trig_sq + i); tv.sinSq = ldt(ct.trig_sq + N + i); tv.cos2Lon = ldt(T + 6 * N + i); tv.sin2Lon = ldt(T + 7 * N + i);
                tv.sin2Lat = ldt(T + 5 * N + i); tv.cosLon = ldt(T + 2 * N + i);
                break;
            case P_FULL2:
                tv.cosLat = ldt(T + i); tv.sinLat = ldt(T + N + i); tv.cosLon = ldt(T + 2 * N + i); tv.sinLon = ldt(T + 3 * N + i);
                tv.cos2Lat = ldt(T + 4 * N + i); tv.cos2Lon = ldt(T + 6 * N + i); tv.sin2Lon = ldt(T + 7 * N + i); tv.cosSq = ldt(ct.trig_sq + i);
                break;
            default: break;
        }
    }
    int all_and = packed[0], all_or = packed[0];
#pragma unroll
    for (int j = 1; j < kCellEdges; j++) { all_and &= packed[j]; all_or |= packed[j]; }
    const int z = (int)(all_and < 0) & (int)(all_or >= 0);                         // never both (see ell_row_batched): 0, known at run time only
#pragma unroll
    for (int j = 0; j < kCellEdges; j++) {                                        // gathers together, arithmetic after
        const int e = (packed[j] == -1 ? 0 : (packed[j] & 0x7fffffff)) + z;
        le[j] = ldo_y(s.vl_out + e);
        fl[j] = ldo(s.flux + e);
    }
#pragma unroll
    for (int j = 0; j < kCellEdges; j++) {
        if (packed[j] != -1) {
            const double ndir = (packed[j] < 0) ? 1.0 : -1.0;                     // -dir, mesh.cpp:3246
            const double coeff = exact_div(ndir * le[j], area, ra);
            div += coeff * fl[j];
        }
    }
    st.x += ab3_increment(div, s.ch1[i], s.ch2[i], p.dt, mode);
    s.chw[i] = div;
    if (kPot && p.potential != P_NONE) st.y = tidal_potential(p, next, tv);        // forcing(current_time + dt) of the NEXT step, tidalPotentials.cpp:80-172
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
        const double zeta = ell_row(t.curl, i, [&](int e) { return ldo_x(s.vl_in + e); });
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
        auto vel = [&](int e) { return ldo_x(s.vl_in + e); };
        const double x = ell_row(t.rbf[0], i, vel), y = ell_row(t.rbf[1], i, vel), z = ell_row(t.rbf[2], i, vel);
        s.ekin[i] = 0.5 * (x * x + y * y + z * z);
    }
}

// kEnergy: the dissipated energy of v^n (interpolation.cpp:31-59 + energy.cpp:32-60) is summed on the way — the tangential velocity of
// an edge is the sum over the SAME ten neighbours this kernel gathers {F, q} from, with the weights w l_e' / d_e it already holds
// (ncoef), so only v_e' itself is gathered on top — instead of in a pass of its own over the TRiSK tables (edge_diag_kernel: 150 B per
// edge). Order and association of that sum differ from the diagnostic kernel's (reference slot order instead of ascending ids, the
// division by d_e folded into the weight): ~1e-16 relative on a quantity whose bar is 1e-12 (a parallel sum either way).
template <bool kEnergy>
__global__ void __launch_bounds__(kNlThreads) nl_edge_step_flux_kernel(NlTables t, Physics p, NlState s, int mode) {
    const int e = blockIdx.x * kNlThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < t.n_edges) {
        const int2 c = t.cells[e];
        const double2 G = t.grad[e];
        const double2 own = s.vl_in[e];
        const double2 in = s.eu_in[c.x], out = s.eu_in[c.y];
        double dv = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;
        const double q_e = s.fq[e].y;
        double F_tang_q = 0.0;
        int nf[kStencil];
        double nc[kStencil];
        double2 no[kStencil];
        double nv[kEnergy ? kStencil : 1];
#pragma unroll
        for (int j = 0; j < kStencil; j++) {                                           // table values, then gathers, then the sum in slot order
            nf[j] = ldt(t.nid + (size_t)j * t.estride + e);
            nc[j] = ldt(t.ncoef + (size_t)j * t.estride + e);
        }
        // every gather address depends on every table value (z is 0: ids are >= -1 and never all 0x80000001 ... but only at run time), so
        // the table loads have to be issued, all of them, before the first gather -- ptxas otherwise interleaves slot by slot
        int all_ids = nf[0], all_hi = __double2hiint(nc[0]);
#pragma unroll
        for (int j = 1; j < kStencil; j++) { all_ids &= nf[j]; all_hi &= __double2hiint(nc[j]); }
        const int z = (int)(all_ids == (int)0x80000001) & (int)(all_hi == (int)0x80000001);
#pragma unroll
        for (int j = 0; j < kStencil; j++) no[j] = ldo(s.fq + ((nf[j] >= 0 ? nf[j] : e) + z));   // {F_e', q_e'}
        if (kEnergy) {
#pragma unroll
            for (int j = 0; j < kStencil; j++) nv[j] = ldo_x(s.vl_in + ((nf[j] >= 0 ? nf[j] : e) + z));
        }
        double vt = 0.0;
#pragma unroll
        for (int j = 0; j < kStencil; j++) {
            if (nf[j] >= 0) {
                const double2 o = no[j];
                F_tang_q += nc[j] * o.x * (q_e + o.y) * 0.5;
                if (kEnergy) vt += nc[j] * nv[j];
            }
        }
        const double dx = t.dist[e];
        if (kEnergy) {                                                                 // the stencil values are dead from here on
            const double sq = own.x * own.x + vt * vt;
            const double eps = p.friction == 0 ? p.alpha * 1000.0 * p.h * sq : p.alpha / p.h * sqrt(sq) * sq;      // energy.cpp:34 / :48-49
            e_area = eps * (dx * own.y);
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
        auto htot = [&](int i) { return p.h + ldo_x(s.eu_in + i); };
        const double d2_inner = ell_row(t.d2[0], e, htot), d2_outer = ell_row(t.d2[1], e, htot);
        const double fact = 1. / 12.0, beta = 1.0;
        const double dx2 = dx * dx * fact;
        const double vel = v;
        s.flux[e] = vel * 0.5 * (htot(c.x) + htot(c.y)) - dx2 * (d2_outer + d2_inner) * vel + dx2 * beta * fabs(vel) * (d2_outer - d2_inner);
    }
    if (!kEnergy) return;
    // deterministic grid-wide sum (as block_sum_and_publish, odis_kernels.cu): warp -> block -> the last block adds the partials in order
    __shared__ double warp_sums[kNlThreads / 32];
    __shared__ bool is_last;
    for (int o = 16; o > 0; o >>= 1) e_area += __shfl_down_sync(0xffffffffu, e_area, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = e_area;
    __syncthreads();
    if (warp == 0) {
        double y = (lane < kNlThreads / 32) ? warp_sums[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) y += __shfl_down_sync(0xffffffffu, y, o);
        if (lane == 0) {
            s.block_partial[blockIdx.x] = y;
            __threadfence();
            is_last = (atomicAdd(s.ticket, 1u) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kNlThreads) acc += ((volatile double*)s.block_partial)[i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) warp_sums[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < kNlThreads / 32; w++) tot += warp_sums[w];
            *s.energy_out = tot;
            *s.ticket = 0u;
        }
    }
}

}  // namespace

void launch_step_nonlinear(const NlTables& t, const Physics& p, const NlState& s, int mode, cudaStream_t stream) {
    auto grid = [](int n) { return (unsigned)((n + kNlThreads - 1) / kNlThreads); };
    nl_vertex_kernel<<<grid(t.n_vertices), kNlThreads, 0, stream>>>(t, p, s);
    nl_edge_prep_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_cell_ekin_kernel<<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, s);
    nl_edge_step_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s, mode);
    nl_flux_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_cell_step_kernel<false><<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, mode, CellTables{}, StepScalars{});
}

void launch_step_nonlinear_folded(const NlTables& t, const Physics& p, const NlState& s, int mode, const CellTables& ct, const StepScalars& next,
                                  cudaStream_t stream) {
    auto grid = [](int n) { return (unsigned)((n + kNlThreads - 1) / kNlThreads); };
    const unsigned vb = grid(t.n_vertices);
    nl_vertex_ekin_kernel<<<vb + grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, (int)vb);
    nl_edge_prep_kernel<<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s);
    nl_edge_step_flux_kernel<true><<<grid(t.n_edges), kNlThreads, 0, stream>>>(t, p, s, mode);
    nl_cell_step_kernel<true><<<grid(t.n_cells), kNlThreads, 0, stream>>>(t, p, s, mode, ct, next);
}

}  // namespace odis
