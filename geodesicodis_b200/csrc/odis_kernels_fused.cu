// One kernel per LTE time step (sm_100a): the cell update of step n-1 and the edge update of step n fused.
//
// In the reference loop (/root/reference/src/timeIntegrator.cpp:205-313) the surface displacement of a
// step is computed from the NEW velocities (updateEta after the velocity update), so a step needs two
// grid-wide dependencies: v^{n+1} <- (v^n, eta^n) and eta^{n+1} <- v^{n+1}. The two-launch kernels
// (odis_kernels.cu / odis_kernels_pipe.cu) follow that literally. Here the second half is deferred into
// the next step's kernel: everything eta^{n} = eta^{n-1} + dt*AB3(h Div v^{n}) needs — the velocities of
// the cell's 5/6 edges — is already gathered by every edge of that cell for its TRiSK stencil (the stencil
// of edge e IS the other edges of its two cells). So each edge thread
//   1. gathers {v,l} of its 10 stencil edges (as before),
//   2. recomputes eta^n of its two cells from them, in the reference's summation order (ascending
//      reference edge id, prepared per edge on the host as a 60-bit slot map), reading the cells'
//      {eta^{n-1}, U}, tendency history and area through L1/L2,
//   3. does the edge update with those eta^n,
// and one designated edge per cell (the lowest-numbered one this rank updates) stores {eta^n, U(next)}
// and the new tendency into the out-of-place buffers. Redundant arithmetic (each cell is recomputed by
// all of its edges) buys: one launch and one grid-wide dependency per step instead of two, no separate
// pass over the cell arrays, and — across GPUs — one halo exchange per step (velocities only; halo
// cells are recomputed locally).
// Streaming tables are staged through shared memory by cp.async.bulk + mbarrier exactly as in
// odis_kernels_pipe.cu. Results are bit-identical to the two-launch kernels and to the CPU reference.
#include "odis_kernels.cuh"

namespace odis {
namespace {

constexpr int kTile = 128;
constexpr int kGroups = 3;            // consumer groups (4 warps each) per CTA
constexpr int kStages = 5;
constexpr int kConsumerWarps = kGroups * kTile / 32;
constexpr int kFusedThreads = 32 + kGroups * kTile;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// read-once table rows are tagged evict-first in L2 so that they do not push the gathered state ({v,l}, {eta,U}),
// which the next kernel re-reads, out of the 126 MB L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ double2 ld_gather(const double2* p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_gather(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
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
__device__ __forceinline__ double ab3_increment(double f0, double f1, double f2, double dt, int mode) {
    const double a = 23. / 12., b = -16. / 12., c = 5. / 12.;
    if (mode == AB3_FULL) return (a * f0 + b * f1 + c * f2) * dt;
    return f0 * dt;
}
__device__ __forceinline__ double dissipation_flux(const Physics& p, double vn, double vt) {
    const double sq = vn * vn + vt * vt;
    if (p.friction == 0) return p.alpha * 1000.0 * p.h * sq;     // energy.cpp:34
    return p.alpha / p.h * sqrt(sq) * sq;                         // energy.cpp:48-49
}

// Tidal potential of one cell (tidalPotentials.cpp:80-172); only the edge that stores a cell evaluates it. The
// table values are fetched early, together with the other gathers, and combined after the cell update.
struct TrigVals { double a[8]; };
__device__ __forceinline__ TrigVals load_trig_at(const FusedTables& t, int potential, int c) {
    const size_t N = (size_t)t.cell_stride;
    const double* T = t.trig;
    TrigVals v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.a[i] = 0.0;
    switch (potential) {
        case P_ECC:
            v.a[0] = ld_gather(t.trig_sq + c); v.a[1] = ld_gather(t.trig_sq + N + c); v.a[2] = ld_gather(T + 6 * N + c); v.a[3] = ld_gather(T + 7 * N + c);
            break;
        case P_OBLIQ:
            v.a[0] = ld_gather(T + 5 * N + c); v.a[1] = ld_gather(T + 2 * N + c);
            break;
        case P_OBLIQ_WEST:
            v.a[0] = ld_gather(T + c); v.a[1] = ld_gather(T + N + c); v.a[2] = ld_gather(T + 2 * N + c); v.a[3] = ld_gather(T + 3 * N + c);
            break;
        case P_FULL:
            v.a[0] = ld_gather(t.trig_sq + c); v.a[1] = ld_gather(t.trig_sq + N + c); v.a[2] = ld_gather(T + 6 * N + c); v.a[3] = ld_gather(T + 7 * N + c);
            v.a[4] = ld_gather(T + 5 * N + c); v.a[5] = ld_gather(T + 2 * N + c);
            break;
        case P_FULL2:
            v.a[0] = ld_gather(T + c); v.a[1] = ld_gather(T + N + c); v.a[2] = ld_gather(T + 2 * N + c); v.a[3] = ld_gather(T + 3 * N + c);
            v.a[4] = ld_gather(T + 4 * N + c); v.a[5] = ld_gather(T + 6 * N + c); v.a[6] = ld_gather(T + 7 * N + c); v.a[7] = ld_gather(t.trig_sq + c);
            break;
        default: break;
    }
    return v;
}
__device__ __forceinline__ double tidal_potential_of(const Physics& p, const StepScalars& m, const TrigVals& v) {
    switch (p.potential) {
        case P_ECC: {
            const double cosSq = v.a[0], sinSq = v.a[1], cos2Lon = v.a[2], sin2Lon = v.a[3];
            return p.factor * ((1. - 3. * sinSq) * m.cosM + cosSq * (3. * m.cosM * cos2Lon + 4. * m.sinM * sin2Lon));
        }
        case P_OBLIQ:
            return p.factor * m.cosM * v.a[0] * v.a[1];
        case P_OBLIQ_WEST: {
            const double cosLat = v.a[0], sinLat = v.a[1], cosLon = v.a[2], sinLon = v.a[3];
            return 3 * p.factor * sinLat * cosLat * (cosLon * m.cosM - sinLon * m.sinM);
        }
        case P_FULL: {
            const double cosSq = v.a[0], sinSq = v.a[1], cos2Lon = v.a[2], sin2Lon = v.a[3], sin2Lat = v.a[4], cosLon = v.a[5];
            return p.factor * ((1 - 3 * sinSq) * m.cosM + cosSq * (3 * m.cosM * cos2Lon + 4 * m.sinM * sin2Lon)) +
                   p.factor2 * m.cosM * sin2Lat * cosLon;
        }
        case P_FULL2: {
            const double cosLat = v.a[0], sinLat = v.a[1], cosLon = v.a[2], sinLon = v.a[3], cos2Lat = v.a[4], cos2Lon = v.a[5], sin2Lon = v.a[6],
                         cosSq = v.a[7];
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
            return 0.0;
    }
}

struct __align__(16) FusedStage {
    int sid[kStencil][kTile];        //  5120 B
    double sw[kStencil][kTile];      // 10240 B
    int2 cells[kTile];               //  1024 B
    double2 grad[kTile];             //  2048 B
    double dist[kTile];              //  1024 B
    double fcor[kTile];              //  1024 B
    double2 own[kTile];              //  2048 B
    double h1[kTile];                //  1024 B
    double h2[kTile];                //  1024 B
    unsigned long long cmap[kTile];  //  1024 B
};
constexpr uint32_t kFusedStageBytes = sizeof(FusedStage);
static_assert(kFusedStageBytes == 25600, "fused stage layout");

__global__ void __launch_bounds__(kFusedThreads, 1)
    step_fused_kernel(FusedTables t, Physics p, FusedState s, int mode_edge, int mode_cell, int update_eta, StepScalars next, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedStage* stages = reinterpret_cast<FusedStage*>(smem_raw);
    double2* scratch_all = reinterpret_cast<double2*>(smem_raw + kStages * sizeof(FusedStage));   // [warp][11][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(FusedStage) + (size_t)kConsumerWarps * 11 * 32 * sizeof(double2));
    uint64_t* empty = full + kStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, kTile / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t S = (size_t)t.e.stride;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            for (int i = 0; i < my_tiles; i++) {
                const int st = i % kStages;
                if (i >= kStages) mbar_wait(empty + st, ((i / kStages) - 1) & 1);
                FusedStage* d = stages + st;
                const size_t e0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile;
                mbar_expect_tx(full + st, kFusedStageBytes);
#pragma unroll
                for (int j = 0; j < kStencil; j++) bulk_g2s(d->sid[j], t.e.sid + j * S + e0, kTile * 4, full + st, pol);
#pragma unroll
                for (int j = 0; j < kStencil; j++) bulk_g2s(d->sw[j], t.e.sw + j * S + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->cells, t.e.cells + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->grad, t.e.grad + e0, kTile * 16, full + st, pol);
                bulk_g2s(d->dist, t.e.dist + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->fcor, t.e.fcor + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->own, s.vl_in + e0, kTile * 16, full + st, pol);
                bulk_g2s(d->h1, s.h1 + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->h2, s.h2 + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->cmap, t.cmap + e0, kTile * 8, full + st, pol);
            }
        }
        return;
    }
    const int g = (warp - 1) / (kTile / 32);
    const int tl = (int)threadIdx.x - 32 - g * kTile;
    double2* scratch = scratch_all + (size_t)(warp - 1) * 11 * 32;       // this warp's [11][32] {v,l} pairs
    double warp_energy = 0.0;                                            // lane 0: sum over this warp's tiles, in tile order
    for (int i = g; i < my_tiles; i += kGroups) {
        const int st = i % kStages;
        const FusedStage* d = stages + st;
        const int e = (int)(((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile) + tl;
        mbar_wait(full + st, (i / kStages) & 1);
        double e_area = 0.0;
        if (e < t.e.n_edges) {
            // ---- gathers: stencil {v,l}; the two cells' {eta,U}, tendency history, area ----
            double2 nb[kStencil];
#pragma unroll
            for (int j = 0; j < kStencil; j++) {
                const int id = d->sid[j][tl];
                nb[j] = ld_gather(s.vl_in + (id < 0 ? e : id));
            }
            const int2 c = d->cells[tl];
            const int cc[2] = {c.x, c.y};
            double2 eu[2] = {ld_gather(s.eu_in + c.x), ld_gather(s.eu_in + c.y)};
            double b1[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0}, area[2] = {1.0, 1.0};
            if (update_eta) {
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    b1[k] = ld_gather(s.ch1 + cc[k]);
                    b2[k] = ld_gather(s.ch2 + cc[k]);
                    area[k] = ld_gather(t.area + cc[k]);
                }
            }
            const double2 own = d->own[tl];
            const unsigned long long cm = d->cmap[tl];
            // the edge that stores a cell also evaluates its next potential: fetch the table values now
            const int keep = (int)(cm >> 60) & 3;
            TrigVals tv = {};
            if (keep && p.potential != P_NONE) tv = load_trig_at(t, p.potential, (keep & 1) ? c.x : c.y);
            TrigVals tv2 = {};
            if (keep == 3 && p.potential != P_NONE) tv2 = load_trig_at(t, p.potential, c.y);
            // ---- cell part: eta of both cells brought up to the velocities just gathered ----
            if (update_eta) {
                scratch[lane] = own;
#pragma unroll
                for (int j = 0; j < kStencil; j++) scratch[(j + 1) * 32 + lane] = nb[j];
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
                double f0 = 0.0;
                if (update_eta) {
                    double div = 0.0;                                             // updateEta.cpp:39, mesh.cpp:3246
                    const double ra = __drcp_rn(area[k]);
#pragma unroll
                    for (int m = 0; m < kCellEdges; m++) {
                        const unsigned int field = (unsigned int)(cm >> ((k * kCellEdges + m) * 5)) & 31u;
                        const unsigned int idx = field & 15u;
                        if (idx != 15u) {
                            const double2 ed = scratch[idx * 32 + lane];
                            const double ndir = (field & 16u) ? 1.0 : -1.0;       // -dir; dir = -1 when the cell is the edge's outer cell
                            const double coeff = exact_div(ndir * ed.y, area[k], ra);
                            div += (p.h * coeff) * ed.x;
                        }
                    }
                    f0 = div;
                    eu[k].x += ab3_increment(f0, b1[k], b2[k], p.dt, mode_cell);  // temporalOperators.cpp:41,55,64
                }
                if ((cm >> (60 + k)) & 1ull) {                                    // this edge stores the cell
                    const double u_next = p.potential != P_NONE ? tidal_potential_of(p, next, (k == 1 && keep == 3) ? tv2 : tv) : eu[k].y;
                    s.eu_out[cc[k]] = make_double2(eu[k].x, u_next);
                    if (update_eta) s.chw[cc[k]] = f0;
                }
            }
            // ---- edge part (as edge_step_kernel) ----
            const double dd = d->dist[tl], fc = d->fcor[tl];
            double cor = 0.0, vt = 0.0;
            const double rd = __drcp_rn(dd);
#pragma unroll
            for (int j = 0; j < kStencil; j++) {
                const double w = d->sw[j][tl];
                const double coeff = exact_div(fc * w * nb[j].y, dd, rd);         // mesh.cpp:2881
                cor += coeff * nb[j].x;
                vt += nb[j].x * w * nb[j].y;                                      // interpolation.cpp:43
            }
            vt = exact_div(vt, dd, rd);
            e_area = dissipation_flux(p, own.x, vt) * (dd * own.y);
            const double2 G = d->grad[tl];
            const double grad = (-p.g * G.x) * eu[0].x + (-p.g * G.y) * eu[1].x;  // updateMomentum.cpp:42
            const double f0e = grad + cor;
            const double drag = (-p.alpha) * own.x + (G.x * eu[0].y + G.y * eu[1].y);   // timeIntegrator.cpp:219
            double v = own.x + ab3_increment(f0e, d->h1[tl], d->h2[tl], p.dt, mode_edge);
            v += p.dt * drag;                                                     // timeIntegrator.cpp:242
            s.vl_out[e] = make_double2(v, own.y);
            if (mode_edge == AB3_SECOND) s.h1[e] = f0e;
            else s.h2[e] = f0e;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);
        for (int o = 16; o > 0; o >>= 1) e_area += __shfl_down_sync(0xffffffffu, e_area, o);
        warp_energy += e_area;
    }
    // energy diagnostic: warp sums -> CTA sum (fixed order) -> the last CTA to finish adds the CTA sums in index order
    __shared__ double warp_sums[kConsumerWarps];
    __shared__ bool is_last;
    if (lane == 0) warp_sums[warp - 1] = warp_energy;
    asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kTile) : "memory");  // consumers only (the producer warp has left)
    if (threadIdx.x == 32) {
        double tot = 0.0;
        for (int w = 0; w < kConsumerWarps; w++) tot += warp_sums[w];
        s.block_partial[blockIdx.x] = tot;
        __threadfence();
        is_last = (atomicAdd(s.ticket, 1u) == gridDim.x - 1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kTile) : "memory");
    if (is_last && warp == 1) {
        __threadfence();
        double acc = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) acc += ((volatile double*)s.block_partial)[b];
        // fixed order: lane-strided partial sums, then a shuffle tree
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            *s.energy_out = acc;
            *s.ticket = 0u;
        }
    }
}

}  // namespace

static int g_sms_fused = 0;

cudaError_t launch_step_fused(const FusedTables& t, const Physics& p, const FusedState& s, int mode_edge, int mode_cell, int update_eta,
                              const StepScalars& next, cudaStream_t stream) {
    static bool configured_dev[64] = {false};      // the opt-in shared-memory size is a per-device attribute
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 63];
    const size_t smem = kStages * sizeof(FusedStage) + (size_t)kConsumerWarps * 11 * 32 * sizeof(double2) + 2 * kStages * sizeof(uint64_t);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(step_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaDeviceGetAttribute(&g_sms_fused, cudaDevAttrMultiProcessorCount, cur_dev);
        if (g_sms_fused <= 0) g_sms_fused = 148;
        configured = true;
    }
    const int n_tiles = (t.e.n_edges + kTile - 1) / kTile;
    const int grid = n_tiles < g_sms_fused ? n_tiles : g_sms_fused;
    step_fused_kernel<<<grid, kFusedThreads, smem, stream>>>(t, p, s, mode_edge, mode_cell, update_eta, next, n_tiles);
    return cudaGetLastError();
}

}  // namespace odis
