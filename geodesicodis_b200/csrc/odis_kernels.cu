// LTE time-step kernels (see odis_kernels.cuh for the mapping onto the reference functions).
#include "odis_kernels.cuh"
#include "odis_potential.cuh"

namespace odis {

namespace {

// Streaming (read-once) table loads: keep them out of L1 so that L1 holds the gathered velocity / displacement
// neighbourhoods instead, and tag them evict-first in L2 so that they do not push the state arrays the next kernel
// gathers from ({v,l}, {eta,U}) out of the 126 MB L2.
__device__ __forceinline__ unsigned long long stream_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(stream_policy()));
    return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(stream_policy()));
    return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(stream_policy()));
    return v;
}
__device__ __forceinline__ int2 ld_stream(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(stream_policy()));
    return v;
}
// Gathered loads of field values that other kernels write (plain coherent loads, cached). volatile asm:
// the kernels group every load of a phase at its start, and program order must stay issue order.
__device__ __forceinline__ double2 ld_gather(const double2* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_plain(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// boundary CTAs of a partitioned run: ghost slots are written by other GPUs while the kernel runs, so bypass L1
__device__ __forceinline__ double2 ld_gather_cg(const double2* p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// spin (bounded) until all neighbours have published the exchange ctl->epoch[0] (see HaloInline)
__device__ __forceinline__ void halo_wait_all(const HaloWait& w, StepCtl* ctl) {
    const unsigned long long ev = ((volatile unsigned long long*)ctl->epoch)[0];
    const long long limit = ctl->spin_cycles > 0 ? ctl->spin_cycles : kHaloSpinCycles;
    const long long t0 = clock64();
    bool late = false;
    for (int k = 0; k < w.n_peers; k++) {
        unsigned long long seen;
        const unsigned long long* fv = w.flag[k];
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fv) : "memory");
        } while (seen < ev && clock64() - t0 < limit);
        late |= seen < ev;
    }
    if (late) ctl->pad = 1ull;      // a neighbour never arrived: reported by the host (ODIS_ERR_STATE), no hang
}

// x / d with IEEE round-to-nearest result, given y = RN(1/d) (Markstein: one reciprocal shared by all the
// quotients of an edge / cell instead of a ~35-instruction division each). q0 = RN(x*y) is refined twice
// through exactly computed residuals; the final fused multiply-add rounds to the correctly rounded quotient
// (checked against hardware division on 5e8 random and adversarial operand pairs).
__device__ __forceinline__ double exact_div(double x, double d, double y) {
    const double q0 = __dmul_rn(x, y);
    const double r0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, x);
    return __fma_rn(r1, y, q1);
}

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

#ifndef ODIS_EDGE_MIN_BLOCKS
#define ODIS_EDGE_MIN_BLOCKS 1
#endif
template <int kThreads>
__global__ void __launch_bounds__(kThreads, ODIS_EDGE_MIN_BLOCKS * 256 / kThreads) edge_step_kernel(EdgeTables t, Physics p, EdgeState s, int mode) {
    const int e = blockIdx.x * kThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < t.n_edges) {
        const int F = t.stride;
        // ---- phase A: every load that does not depend on another load, issued back to back ----
        int id[kStencil];
        double w[kStencil];
#pragma unroll
        for (int j = 0; j < kStencil; j++) id[j] = ld_stream(t.sid + (size_t)j * F + e);
        const int2 c = ld_stream(t.cells + e);
#pragma unroll
        for (int j = 0; j < kStencil; j++) w[j] = ld_stream(t.sw + (size_t)j * F + e);
        const double2 G = ld_stream(t.grad + e);
        const double d = ld_stream(t.dist + e);
        const double fc = ld_stream(t.fcor + e);
        const double2 own = ld_gather(s.vl_in + e);        // {v_e, l_e}
        const double f1 = ld_plain(s.h1 + e), f2 = ld_plain(s.h2 + e);
        // ---- phase B: the gathers (pad slots, id -1 / weight 0, gather the edge itself and add an exact zero) ----
        double2 nb[kStencil];
#pragma unroll
        for (int j = 0; j < kStencil; j++) nb[j] = ld_gather(s.vl_in + (id[j] < 0 ? e : id[j]));   // {v_e', l_e'}
        const double2 in = ld_gather(s.eu + c.x), out = ld_gather(s.eu + c.y);
        // ---- phase C: arithmetic in the reference's order ----
        // Coriolis / tangential reconstruction over the stencil (mesh.cpp:2874-2883 coefficients;
        // interpolation.cpp:41-45 for v_tang)
        double cor = 0.0, vt = 0.0;
        const double rd = __drcp_rn(d);
#pragma unroll
        for (int j = 0; j < kStencil; j++) {
            const double coeff = exact_div(fc * w[j] * nb[j].y, d, rd);      // -2 Omega sin(lat) w l_e' / d_e
            cor += coeff * nb[j].x;
            vt += nb[j].x * w[j] * nb[j].y;
        }
        vt = exact_div(vt, d, rd);
        e_area = dissipation_flux(p, own.x, vt) * (d * own.y);               // eps_e * A_e, A_e = d_e l_e (mesh.cpp:1093)
        // dv/dt = -g G eta + C v      (updateMomentum.cpp:42)
        const double grad = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;
        const double f0 = grad + cor;
        // drag + tidal forcing         (timeIntegrator.cpp:219)
        const double drag = (-p.alpha) * own.x + (G.x * in.y + G.y * out.y);
        double v = own.x + ab3_increment(f0, f1, f2, p.dt, mode);            // temporalOperators.cpp:41,55,64
        v += p.dt * drag;                                                    // timeIntegrator.cpp:242
        s.vl_out[e] = make_double2(v, own.y);
        // history: FIRST keeps f0 as level 2, SECOND as level 1, FULL shifts (the host swaps h1/h2)
        if (mode == AB3_SECOND) s.h1[e] = f0;
        else s.h2[e] = f0;
    }
    // energy diagnostic: one partial per warp, no block barrier (warps retire independently); the
    // partials are summed in index order by block 0 of the cell kernel that follows.
    for (int o = 16; o > 0; o >>= 1) e_area += __shfl_down_sync(0xffffffffu, e_area, o);
    if ((threadIdx.x & 31) == 0) s.block_partial[(blockIdx.x * kThreads + threadIdx.x) >> 5] = e_area;
}

// Sum of n partials in a fixed order by one block (strided per-thread sums, then a shuffle/shared tree).
template <int kThreads>
__device__ __forceinline__ void block_reduce_partials(const double* partial, int n, double* out) {
    __shared__ double warp_sums[kThreads / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kThreads / 32; w++) tot += warp_sums[w];
        *out = tot;
    }
}

__device__ __forceinline__ TrigValues load_trig(const CellTables& t, int potential, int i) {
    const size_t N = (size_t)t.n_cells;
    const double* T = t.trig;
    TrigValues v = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    switch (potential) {
        case P_ECC:
            v.cosSq = ld_stream(t.trig_sq + i); v.sinSq = ld_stream(t.trig_sq + N + i);
            v.cos2Lon = ld_stream(T + 6 * N + i); v.sin2Lon = ld_stream(T + 7 * N + i);
            break;
        case P_OBLIQ:
            v.sin2Lat = ld_stream(T + 5 * N + i); v.cosLon = ld_stream(T + 2 * N + i);
            break;
        case P_OBLIQ_WEST:
            v.cosLat = ld_stream(T + i); v.sinLat = ld_stream(T + N + i);
            v.cosLon = ld_stream(T + 2 * N + i); v.sinLon = ld_stream(T + 3 * N + i);
            break;
        case P_FULL:
            v.cosSq = ld_stream(t.trig_sq + i); v.sinSq = ld_stream(t.trig_sq + N + i);
            v.cos2Lon = ld_stream(T + 6 * N + i); v.sin2Lon = ld_stream(T + 7 * N + i);
            v.sin2Lat = ld_stream(T + 5 * N + i); v.cosLon = ld_stream(T + 2 * N + i);
            break;
        case P_FULL2:
            v.cosLat = ld_stream(T + i); v.sinLat = ld_stream(T + N + i);
            v.cosLon = ld_stream(T + 2 * N + i); v.sinLon = ld_stream(T + 3 * N + i);
            v.cos2Lat = ld_stream(T + 4 * N + i);
            v.cos2Lon = ld_stream(T + 6 * N + i); v.sin2Lon = ld_stream(T + 7 * N + i);
            v.cosSq = ld_stream(t.trig_sq + i);
            break;
        default: break;
    }
    return v;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) cell_step_kernel(CellTables t, Physics p, CellState s, int mode, StepScalars next,
                                                             int flags, HaloInline halo) {
    if (blockIdx.x == 0 && s.energy_out != nullptr)      // finish the edge kernel's energy sum (see edge_step_kernel)
        block_reduce_partials<kThreads>(s.energy_partial, s.n_energy_partials, s.energy_out);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    // partitioned runs: the last CTAs hold the cells that read ghost edges (own boundary cells, then the ghost cells)
    const bool bnd_cta = (int)((blockIdx.x + 1) * kThreads) > halo.wait_from;
    if (bnd_cta) {
        if (threadIdx.x == 0) halo_wait_all(halo.wait_v, halo.ctl);
        __syncthreads();
    }
    if (i < t.n_active) {
        const int N = t.n_cells;
        // ---- phase A: independent loads ----
        const int update_eta = flags & CELL_UPDATE_ETA;
        int packed[kCellEdges];
#pragma unroll
        for (int j = 0; j < kCellEdges; j++) packed[j] = update_eta ? ld_stream(t.eid + (size_t)j * N + i) : -1;
        double2 st = ld_gather(s.eu_in + i);
        double area = 1.0, f1 = 0.0, f2 = 0.0;
        if (update_eta) {
            area = ld_stream(t.area + i);
            f1 = ld_plain(s.h1 + i);
            f2 = ld_plain(s.h2 + i);
        }
        TrigValues tv = load_trig(t, (flags & CELL_UPDATE_U) ? p.potential : (int)P_NONE, i);
        if (s.next_dev != nullptr) {                         // graph replay: this step's time factors were left by the edge kernel
            next.cosM = ld_plain(&s.next_dev->cosM);
            next.sinM = ld_plain(&s.next_dev->sinM);
            if (p.potential == P_FULL2) {
                next.cos2M = ld_plain(&s.next_dev->cos2M); next.sin2M = ld_plain(&s.next_dev->sin2M);
                next.cos3M = ld_plain(&s.next_dev->cos3M); next.cos4M = ld_plain(&s.next_dev->cos4M);
            }
        }
        // ---- phase B: gathers ----
        double2 ed[kCellEdges];
        if (!bnd_cta) {
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) ed[j] = ld_gather(s.vl + (packed[j] == -1 ? 0 : (packed[j] & 0x7fffffff)));
        } else {
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) ed[j] = ld_gather_cg(s.vl + (packed[j] == -1 ? 0 : (packed[j] & 0x7fffffff)));
        }
        // ---- phase C ----
        if (update_eta) {
            // d eta/dt = h Div v   (updateEta.cpp:39; D_ie = -dir l_e / A_i, mesh.cpp:3246)
            double div = 0.0;
            const double ra = __drcp_rn(area);
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) {
                if (packed[j] != -1) {                                    // the 12 pentagons have 5 edges
                    const double ndir = (packed[j] < 0) ? 1.0 : -1.0;     // -dir: dir = -1 for the outer cell
                    const double coeff = exact_div(ndir * ed[j].y, area, ra);
                    div += (p.h * coeff) * ed[j].x;
                }
            }
            const double f0 = div;
            st.x += ab3_increment(f0, f1, f2, p.dt, mode);
            s.hw[i] = f0;
        }
        if ((flags & CELL_UPDATE_U) && p.potential != P_NONE) st.y = tidal_potential(p, next, tv);
        s.eu_out[i] = st;
    }
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) edge_diag_kernel(EdgeTables t, Physics p, const double2* vl, const double2* normal,
                                                             double2* v_avg, double* energy_diss, double* block_partial,
                                                             unsigned int* ticket, double* energy_out) {
    const int e = blockIdx.x * kThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < t.n_edges) {
        const int F = t.stride;
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

// ---- halo exchange ----
__global__ void halo_exchange_kernel(int n, const int* __restrict__ local_idx, const int* __restrict__ remote_idx, const int* __restrict__ peer,
                                     const double2* __restrict__ src, HaloRemote remote, HaloWait w, int flag_slot, int kind, StepCtl* ctl,
                                     unsigned int* ticket) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) remote.data[peer[k]][remote_idx[k]] = src[local_idx[k]];      // direct store into the neighbour GPU over NVLink
    __threadfence_system();                                                  // my stores are performed before the ticket
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    // every block's data is out: publish the epoch to the peers, then wait for theirs
    __threadfence_system();
    const unsigned long long epoch = ((volatile unsigned long long*)ctl->epoch)[kind] + 1ull;
    if ((int)threadIdx.x < w.n_peers) {
        unsigned long long* f = remote.flags[threadIdx.x] + flag_slot;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        unsigned long long seen;
        const long long t0 = clock64(), limit = ctl->spin_cycles > 0 ? ctl->spin_cycles : kHaloSpinCycles;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(w.flag[threadIdx.x]) : "memory");
        } while (seen < epoch && clock64() - t0 < limit);
        if (seen < epoch) ctl->pad = 1ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ctl->epoch[kind] = epoch;
        *ticket = 0u;
    }
}

// bounded wait for the neighbours' last pushes (before state is overwritten / read back, and before a cell kernel
// variant that does not wait itself)
__global__ void halo_drain_kernel(HaloWait wv, StepCtl* ctl) {
    if (threadIdx.x == 0) halo_wait_all(wv, ctl);
}

// ---- renumbering kernels (set_state / get_field; not on the per-step path) ----
__global__ void scatter_x_kernel(int n, const int* __restrict__ perm, const double* __restrict__ src, double2* dst, int zero_y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 v = dst[i];
    v.x = src ? src[perm ? perm[i] : i] : 0.0;          // perm == nullptr: the source is already in device order
    if (zero_y) v.y = 0.0;
    dst[i] = v;
}
__global__ void scatter_history_kernel(int n, const int* __restrict__ perm, const double* __restrict__ src3, double* lvl0, double* h1,
                                       double* h2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)(perm ? perm[i] : i) * 3;
    lvl0[i] = src3 ? src3[o] : 0.0;
    h1[i] = src3 ? src3[o + 1] : 0.0;
    h2[i] = src3 ? src3[o + 2] : 0.0;
}
__global__ void gather_component_kernel(int n, const int* __restrict__ perm, const double2* __restrict__ src, int component, double* dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 v = src[i];
    dst[perm ? perm[i] : i] = component ? v.y : v.x;    // perm == nullptr: compact, device order
}
__global__ void gather_pair_kernel(int n, const int* __restrict__ perm, const double2* __restrict__ src, double* dst2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 v = src[i];
    const size_t o = (size_t)(perm ? perm[i] : i) * 2;
    dst2[o] = v.x;
    dst2[o + 1] = v.y;
}
__global__ void gather_scalar_kernel(int n, const int* __restrict__ perm, const double* __restrict__ src, double* dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm ? perm[i] : i] = src[i];
}
__global__ void gather_history_kernel(int n, const int* __restrict__ perm, const double* __restrict__ lvl0, const double* __restrict__ h1,
                                      const double* __restrict__ h2, int which0, double* dst3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a1 = h1[i], a2 = h2[i];
    const size_t o = (size_t)(perm ? perm[i] : i) * 3;
    dst3[o] = which0 == 0 ? lvl0[i] : (which0 == 1 ? a1 : a2);
    dst3[o + 1] = a1;
    dst3[o + 2] = a2;
}

// PLANET forcing, tidalPotentials.cpp:215-221 (see launch_planet_potential)
__global__ void __launch_bounds__(256) planet_potential_kernel(CellTables t, StepScalars m, const StepScalars* dev, double2* eu, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (dev != nullptr) {
        m.cosM = ld_plain(&dev->cosM); m.sinM = ld_plain(&dev->sinM); m.cos2M = ld_plain(&dev->cos2M); m.sin2M = ld_plain(&dev->sin2M);
    }
    const size_t N = (size_t)t.n_cells;
    const double cosLat = t.trig[i], cosLon = t.trig[2 * N + i], sinLon = t.trig[3 * N + i];
    const double cosgam = cosLat * (cosLon * m.cosM + sinLon * m.sinM);
    eu[i].y = m.cos2M * (3. * (cosgam * cosgam) - m.sin2M);
}

// ---- operator surface (odis_op_*, odis_engine.cu): the two loop-level functions of the reference that no step kernel
// covers on their own. Both work on reference-ordered arrays staged on the device. ----
// integrateAB3scalar (temporalOperators.cpp:17-68): the solution and its [n][3] tendency history, updated in place
__global__ void ab3_scalar_kernel(int n, double* __restrict__ sol, double* __restrict__ hist, double dt, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)i * 3;
    const double f0 = hist[o], f1 = hist[o + 1], f2 = hist[o + 2];
    sol[i] += ab3_increment(f0, f1, f2, dt, mode);
    if (mode == AB3_FULL) { hist[o + 2] = f1; hist[o + 1] = f0; }      // :44-45
    else if (mode == AB3_FIRST) hist[o + 2] = f0;                       // :55
    else hist[o + 1] = f0;                                              // :64
}
// updateEnergy (energy.cpp:13-62) from the east/north components [F][2] and the edge areas [F]
template <int kThreads>
__global__ void __launch_bounds__(kThreads) energy_from_components_kernel(int n, Physics p, const double* __restrict__ vel2,
                                                                          const double* __restrict__ areas, double* __restrict__ e_flux,
                                                                          double* block_partial, unsigned int* ticket, double* energy_out) {
    const int e = blockIdx.x * kThreads + threadIdx.x;
    double e_area = 0.0;
    if (e < n) {
        const double u = vel2[(size_t)e * 2], v = vel2[(size_t)e * 2 + 1];
        double eps;
        if (p.friction == 0) eps = p.alpha * 1000.0 * p.h * (u * u + v * v);            // energy.cpp:34
        else eps = p.alpha / p.h * sqrt(u * u + v * v) * (u * u + v * v);                // energy.cpp:48-49
        e_flux[e] = eps;
        e_area = eps * areas[e];                                                          // energy.cpp:36,51
    }
    block_sum_and_publish<kThreads>(e_area, block_partial, ticket, energy_out);
}

template <typename F>
void dispatch_threads(int block_threads, F&& f) {
    switch (block_threads) {
        case 256: f(std::integral_constant<int, 256>()); break;
        case 512: f(std::integral_constant<int, 512>()); break;
        default: f(std::integral_constant<int, 128>()); break;      // measured best on B200 (profiles/)
    }
}

}  // namespace

int edge_grid_blocks(int n_edges, int block_threads) {
    const int bt = (block_threads == 256 || block_threads == 512) ? block_threads : 128;
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
                      int flags, int block_threads, const HaloInline* halo, cudaStream_t stream) {
    HaloInline none;
    none.n_bnd = 0;
    none.wait_from = 0x7fffffff;
    dispatch_threads(block_threads, [&](auto bt) {
        constexpr int kT = decltype(bt)::value;
        cell_step_kernel<kT><<<(t.n_active + kT - 1) / kT, kT, 0, stream>>>(t, p, s, mode, next, flags, halo ? *halo : none);
    });
}
void launch_halo_drain(const HaloWait& wait_v, StepCtl* ctl, cudaStream_t stream) {
    halo_drain_kernel<<<1, 32, 0, stream>>>(wait_v, ctl);
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

void launch_halo_exchange(int n, const int* local_idx, const int* remote_idx, const int* peer, const double2* src, const HaloRemote& remote,
                          const HaloWait& w, int flag_slot, int kind, StepCtl* ctl, unsigned int* ticket, cudaStream_t stream) {
    const int blocks = n > 0 ? (n + 255) / 256 : 1;
    halo_exchange_kernel<<<blocks, 256, 0, stream>>>(n, local_idx, remote_idx, peer, src, remote, w, flag_slot, kind, ctl, ticket);
}

void launch_scatter_x(int n, const int* perm, const double* src_ref, double2* dst_new, int zero_y, cudaStream_t stream) {
    scatter_x_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, src_ref, dst_new, zero_y);
}
void launch_scatter_history(int n, const int* perm, const double* src_ref3, double* lvl0_new, double* h1_new, double* h2_new,
                            cudaStream_t stream) {
    scatter_history_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, src_ref3, lvl0_new, h1_new, h2_new);
}
void launch_gather_component(int n, const int* perm, const double2* src_new, int component, double* dst_ref, cudaStream_t stream) {
    gather_component_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, src_new, component, dst_ref);
}
void launch_gather_pair(int n, const int* perm, const double2* src_new, double* dst_ref2, cudaStream_t stream) {
    gather_pair_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, src_new, dst_ref2);
}
void launch_gather_scalar(int n, const int* perm, const double* src_new, double* dst_ref, cudaStream_t stream) {
    gather_scalar_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, src_new, dst_ref);
}
void launch_gather_history(int n, const int* perm, const double* lvl0_new, const double* h1_new, const double* h2_new, int which0,
                           double* dst_ref3, cudaStream_t stream) {
    gather_history_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, perm, lvl0_new, h1_new, h2_new, which0, dst_ref3);
}

void launch_planet_potential(const CellTables& t, const StepScalars& host, const StepScalars* dev, double2* eu, int n, cudaStream_t stream) {
    if (n > 0) planet_potential_kernel<<<(n + 255) / 256, 256, 0, stream>>>(t, host, dev, eu, n);
}
void launch_ab3_scalar(int n, double* sol, double* hist3, double dt, int mode, cudaStream_t stream) {
    if (n > 0) ab3_scalar_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, sol, hist3, dt, mode);
}
void launch_energy_from_components(int n, const Physics& p, const double* vel2, const double* areas, double* e_flux, double* block_partial,
                                   unsigned int* ticket, double* energy_out, cudaStream_t stream) {
    energy_from_components_kernel<128><<<(n + 127) / 128, 128, 0, stream>>>(n, p, vel2, areas, e_flux, block_partial, ticket, energy_out);
}


}  // namespace odis
