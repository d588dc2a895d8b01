// LTE time-step kernels (see odis_kernels.cuh for the mapping onto the reference functions).
#include "odis_kernels.cuh"

namespace odis {

namespace {

// Streaming (read-once) table loads: keep them out of L1 so that L1 holds the gathered
// velocity / displacement neighbourhoods instead.
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_stream(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// Gathered loads of field values that other kernels write (plain coherent loads, cached).
__device__ __forceinline__ double2 ld_gather(const double2* p) { return *p; }

// Adams-Bashforth-3 increment, operation order of temporalOperators.cpp:41-43 / :55 / :64.
__device__ __forceinline__ double ab3_increment(double f0, double f1, double f2, double dt, int mode) {
    const double a = 23. / 12., b = -16. / 12., c = 5. / 12.;
    if (mode == AB3_FULL) return (a * f0 + b * f1 + c * f2) * dt;
    return f0 * dt;
}

// Deterministic grid-wide sum: block tree in shared memory, per-block partials, and the last
// block to finish adds the partials in index order (energy.cpp:36-40 sums serially on the CPU;
// only the association differs).
template <int kThreads>
__device__ __forceinline__ void block_sum_and_publish(double x, double* block_partial, unsigned int* ticket,
                                                      double* out) {
    __shared__ double warp_sums[kThreads / 32];
    __shared__ bool is_last;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        double y = (lane < kThreads / 32) ? warp_sums[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) y += __shfl_down_sync(0xffffffffu, y, o);
        if (lane == 0) {
            block_partial[blockIdx.x] = y;
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) acc += ((volatile double*)block_partial)[i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) warp_sums[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < kThreads / 32; w++) tot += warp_sums[w];
            *out = tot;
            *ticket = 0u;
        }
    }
}

__device__ __forceinline__ double dissipation_flux(const Physics& p, double vn, double vt) {
    const double sq = vn * vn + vt * vt;     // = u^2 + v^2 of interpolation.cpp:57-58 (n, t orthonormal)
    if (p.friction == 0) return p.alpha * 1000.0 * p.h * sq;     // energy.cpp:34
    return p.alpha / p.h * sqrt(sq) * sq;                         // energy.cpp:48-49
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) edge_step_kernel(EdgeTables t, Physics p, EdgeState s, int mode) {
    const int e = blockIdx.x * kThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < t.n_edges) {
        const int F = t.n_edges;
        const double2 own = ld_gather(s.vl_in + e);        // {v_e, l_e}
        const double d = ld_stream(t.dist + e);
        const double fc = ld_stream(t.fcor + e);
        // Coriolis / tangential reconstruction over the stencil (mesh.cpp:2874-2883 coefficients;
        // interpolation.cpp:41-45 for v_tang)
        double cor = 0.0, vt = 0.0;
#pragma unroll
        for (int j = 0; j < kStencil; j++) {
            const int id = ld_stream(t.sid + (size_t)j * F + e);
            const double w = ld_stream(t.sw + (size_t)j * F + e);
            if (id >= 0) {
                const double2 nb = ld_gather(s.vl_in + id);                 // {v_e', l_e'}
                const double coeff = fc * w * nb.y / d;                      // -2 Omega sin(lat) w l_e' / d_e
                cor += coeff * nb.x;
                vt += nb.x * w * nb.y;
            }
        }
        vt /= d;
        e_area = dissipation_flux(p, own.x, vt) * (d * own.y);               // eps_e * A_e, A_e = d_e l_e (mesh.cpp:1093)

        const int2 c = ld_stream(t.cells + e);
        const double2 G = ld_stream(t.grad + e);
        const double2 in = ld_gather(s.eu + c.x), out = ld_gather(s.eu + c.y);
        // dv/dt = -g G eta + C v      (updateMomentum.cpp:42)
        const double grad = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;
        const double f0 = grad + cor;
        // drag + tidal forcing         (timeIntegrator.cpp:219)
        const double drag = (-p.alpha) * own.x + (G.x * in.y + G.y * out.y);
        const double f1 = s.h1[e], f2 = s.h2[e];
        double v = own.x + ab3_increment(f0, f1, f2, p.dt, mode);            // temporalOperators.cpp:41,55,64
        v += p.dt * drag;                                                    // timeIntegrator.cpp:242
        s.vl_out[e] = make_double2(v, own.y);
        // history: FIRST keeps f0 as level 2, SECOND as level 1, FULL shifts (the host swaps h1/h2)
        if (mode == AB3_SECOND) s.h1[e] = f0;
        else s.h2[e] = f0;
    }
    block_sum_and_publish<kThreads>(e_area, s.block_partial, s.ticket, s.energy_out);
}

// Tidal potential at one cell (tidalPotentials.cpp:80-172), same expression shapes.
__device__ __forceinline__ double tidal_potential(const CellTables& t, const Physics& p, const StepScalars& m, int i) {
    const int N = t.n_cells;
    const double* T = t.trig;
    switch (p.potential) {
        case P_ECC: {
            const double cosSq = ld_stream(t.trig_sq + i), sinSq = ld_stream(t.trig_sq + N + i);
            const double cos2Lon = ld_stream(T + 6 * (size_t)N + i), sin2Lon = ld_stream(T + 7 * (size_t)N + i);
            return p.factor * ((1. - 3. * sinSq) * m.cosM + cosSq * (3. * m.cosM * cos2Lon + 4. * m.sinM * sin2Lon));
        }
        case P_OBLIQ: {
            const double sin2Lat = ld_stream(T + 5 * (size_t)N + i), cosLon = ld_stream(T + 2 * (size_t)N + i);
            return p.factor * m.cosM * sin2Lat * cosLon;
        }
        case P_OBLIQ_WEST: {
            const double cosLat = ld_stream(T + i), sinLat = ld_stream(T + (size_t)N + i);
            const double cosLon = ld_stream(T + 2 * (size_t)N + i), sinLon = ld_stream(T + 3 * (size_t)N + i);
            return 3 * p.factor * sinLat * cosLat * (cosLon * m.cosM - sinLon * m.sinM);
        }
        case P_FULL: {
            const double cosSq = ld_stream(t.trig_sq + i), sinSq = ld_stream(t.trig_sq + N + i);
            const double cos2Lon = ld_stream(T + 6 * (size_t)N + i), sin2Lon = ld_stream(T + 7 * (size_t)N + i);
            const double sin2Lat = ld_stream(T + 5 * (size_t)N + i), cosLon = ld_stream(T + 2 * (size_t)N + i);
            return p.factor * ((1 - 3 * sinSq) * m.cosM + cosSq * (3 * m.cosM * cos2Lon + 4 * m.sinM * sin2Lon)) +
                   p.factor2 * m.cosM * sin2Lat * cosLon;
        }
        case P_FULL2: {
            const double cosLat = ld_stream(T + i), sinLat = ld_stream(T + (size_t)N + i);
            const double cosLon = ld_stream(T + 2 * (size_t)N + i), sinLon = ld_stream(T + 3 * (size_t)N + i);
            const double cos2Lat = ld_stream(T + 4 * (size_t)N + i);
            const double cos2Lon = ld_stream(T + 6 * (size_t)N + i), sin2Lon = ld_stream(T + 7 * (size_t)N + i);
            const double cosSq = ld_stream(t.trig_sq + i);
            const double ecc = p.ecc, obl = p.obl;
            double T1, T2, T3;
            T1 = 3. * ecc * (4. - 7. * obl * obl) * m.cosM + 6 * (obl * obl + ecc * ecc * (3 - 7 * obl * obl)) * m.cos2M;
            T1 += 3 * ecc * obl * obl * (7 * m.cos3M + 17 * ecc * m.cos4M);
            T1 *= -(1 - 3 * cos2Lat);
            T2 = (4 + 15 * ecc * ecc + 20 * ecc * m.cosM + 43 * ecc * ecc * m.cos2M) * cosLon;
            T2 += 2 * ecc * (4 + 25 * ecc * m.cosM) * m.sinM * sinLon;
            T2 *= 24 * obl * cosLat * sinLat * m.sinM;
            T3 = obl * obl * (2 + 3 * ecc * ecc + 6 * ecc * m.cosM + 9 * ecc * ecc * m.cos2M) * (m.cosM * cosLon + m.sinM * sinLon);
            T3 += -(obl * obl - 2) * ((6 * ecc * m.cosM + 17 * ecc * ecc * m.cos2M) * cos2Lon + 2 * ecc * (4 + 17 * ecc * m.cosM) * m.sinM * sin2Lon);
            T3 *= 6 * cosSq;
            return p.factor * (T1 + T2 + T3);
        }
        default:
            return 0.0;   // NONE leaves the (zero-initialised) potential untouched, tidalPotentials.cpp:283
    }
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) cell_step_kernel(CellTables t, Physics p, CellState s, int mode, StepScalars next,
                                                             int update_eta) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= t.n_cells) return;
    const int N = t.n_cells;
    double2 st = s.eu[i];
    if (update_eta) {
        const double area = ld_stream(t.area + i);
        // d eta/dt = h Div v   (updateEta.cpp:39; D_ie = -dir l_e / A_i, mesh.cpp:3246)
        double div = 0.0;
#pragma unroll
        for (int j = 0; j < kCellEdges; j++) {
            const int packed = ld_stream(t.eid + (size_t)j * N + i);
            if (packed != -1) {
                const int id = packed & 0x7fffffff;
                const double ndir = (packed < 0) ? 1.0 : -1.0;       // -dir: dir = -1 for the outer cell
                const double2 ed = ld_gather(s.vl + id);
                const double coeff = ndir * ed.y / area;
                div += (p.h * coeff) * ed.x;
            }
        }
        const double f0 = div;
        const double f1 = s.h1[i], f2 = s.h2[i];
        st.x += ab3_increment(f0, f1, f2, p.dt, mode);
        if (mode == AB3_SECOND) s.h1[i] = f0;
        else s.h2[i] = f0;
    }
    if (p.potential != P_NONE) st.y = tidal_potential(t, p, next, i);
    s.eu[i] = st;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) edge_diag_kernel(EdgeTables t, Physics p, const double2* vl, const double2* normal,
                                                             double2* v_avg, double* energy_diss, double* block_partial,
                                                             unsigned int* ticket, double* energy_out) {
    const int e = blockIdx.x * kThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < t.n_edges) {
        const int F = t.n_edges;
        const double2 own = vl[e];
        const double d = t.dist[e];
        double vt = 0.0;
#pragma unroll
        for (int j = 0; j < kStencil; j++) {
            const int id = t.sid[(size_t)j * F + e];
            const double w = t.sw[(size_t)j * F + e];
            if (id >= 0) {
                const double2 nb = vl[id];
                vt += nb.x * w * nb.y;                                 // interpolation.cpp:43
            }
        }
        vt /= d;
        const double2 n = normal[e];
        const double tx = n.y, ty = -n.x;                              // interpolation.cpp:52-55
        const double u = n.x * own.x + tx * vt, v = n.y * own.x + ty * vt;
        if (v_avg) v_avg[e] = make_double2(u, v);
        double eps;
        if (p.friction == 0) eps = p.alpha * 1000.0 * p.h * (u * u + v * v);            // energy.cpp:34
        else eps = p.alpha / p.h * sqrt(u * u + v * v) * (u * u + v * v);                // energy.cpp:48-49
        if (energy_diss) energy_diss[e] = eps;
        e_area = eps * (d * own.y);
    }
    block_sum_and_publish<kThreads>(e_area, block_partial, ticket, energy_out);
}

template <typename F>
void dispatch_threads(int block_threads, F&& f) {
    switch (block_threads) {
        case 128: f(std::integral_constant<int, 128>()); break;
        case 512: f(std::integral_constant<int, 512>()); break;
        default: f(std::integral_constant<int, 256>()); break;
    }
}

}  // namespace

int edge_grid_blocks(int n_edges, int block_threads) {
    const int bt = (block_threads == 128 || block_threads == 512) ? block_threads : 256;
    return (n_edges + bt - 1) / bt;
}

void launch_edge_step(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, int block_threads,
                      cudaStream_t stream) {
    dispatch_threads(block_threads, [&](auto bt) {
        constexpr int kT = decltype(bt)::value;
        edge_step_kernel<kT><<<(t.n_edges + kT - 1) / kT, kT, 0, stream>>>(t, p, s, mode);
    });
}

void launch_cell_step(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                      int update_eta, int block_threads, cudaStream_t stream) {
    dispatch_threads(block_threads, [&](auto bt) {
        constexpr int kT = decltype(bt)::value;
        cell_step_kernel<kT><<<(t.n_cells + kT - 1) / kT, kT, 0, stream>>>(t, p, s, mode, next, update_eta);
    });
}

void launch_edge_diagnostics(const EdgeTables& t, const Physics& p, const double2* vl, const double2* normal, double2* v_avg,
                             double* energy_diss, double* block_partial, unsigned int* ticket, double* energy_out,
                             int block_threads, cudaStream_t stream) {
    dispatch_threads(block_threads, [&](auto bt) {
        constexpr int kT = decltype(bt)::value;
        edge_diag_kernel<kT><<<(t.n_edges + kT - 1) / kT, kT, 0, stream>>>(t, p, vl, normal, v_avg, energy_diss, block_partial,
                                                                           ticket, energy_out);
    });
}

}  // namespace odis
