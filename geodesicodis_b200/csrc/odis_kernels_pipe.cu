// Pipelined (bulk-async staged) versions of the two LTE step kernels for sm_100a.
//
// Same arithmetic, operation for operation, as edge_step_kernel / cell_step_kernel in
// odis_kernels.cu (results are bit-identical); what changes is how the streamed tables reach
// the SM. The direct kernels keep every table value of an edge in registers while its loads
// are in flight (112 registers, 25 % occupancy), so only ~1/4 of a warp's lifetime has DRAM
// requests outstanding. Here each CTA is persistent and warp-specialised:
//
//   warp 0 (one elected lane)   issues `cp.async.bulk` global->shared copies of whole table rows
//                               for tile t+1..t+S-1 (TMA engine, no registers, completion counted
//                               on an mbarrier with expect_tx), S stages deep;
//   consumer warps              wait on the stage's "full" mbarrier, read their table values from
//                               shared memory, issue the dependent gathers ({v,l} of the stencil,
//                               {eta,U} of the two cells) from global/L2, do the arithmetic in the
//                               reference's order, store, and release the stage ("empty" mbarrier).
//
// A tile is 128 consecutive edges (or cells); table rows are SoA with a stride padded to 128, so
// every row segment of a tile is one contiguous, 16-byte aligned bulk copy.
#include "odis_kernels.cuh"

namespace odis {
namespace {

constexpr int kTile = 128;            // edges (cells) per tile = one consumer group of 4 warps
constexpr int kGroups = 2;            // consumer groups per CTA
constexpr int kStages = 4;            // tiles in flight per CTA
constexpr int kPipeThreads = 32 + kGroups * kTile;

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
// global -> shared bulk copy (TMA, 1-D), completion bytes counted on `bar`
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
// after the boundary tile's threads have issued their peer stores and met at a barrier: one thread counts the tile
// done (its system-scope fence is cumulative over the stores the barrier ordered before it); the last one publishes
// the new epoch to the neighbours
__device__ __forceinline__ void halo_tile_done(const HaloInline& h, unsigned int n_units) {
    __threadfence_system();
    if (atomicAdd(h.done, 1u) == n_units - 1u) {
        __threadfence_system();
        const unsigned long long epoch = ((volatile unsigned long long*)h.ctl->epoch)[0] + 1ull;
        for (int k = 0; k < h.n_peers; k++) {
            unsigned long long* f = h.remote.flags[k] + h.flag_slot;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
        *h.done = 0u;
    }
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

// ---------------------------------------------------------------- edge step ----
struct __align__(16) EdgeStage {
    int sid[kStencil][kTile];        //  5120 B
    double sw[kStencil][kTile];      // 10240 B
    int2 cells[kTile];               //  1024 B
    double2 grad[kTile];             //  2048 B
    double dist[kTile];              //  1024 B
    double fcor[kTile];              //  1024 B
    double2 own[kTile];              //  2048 B
    double h1[kTile];                //  1024 B
    double h2[kTile];                //  1024 B
};
constexpr uint32_t kEdgeStageBytes = sizeof(EdgeStage);
static_assert(kEdgeStageBytes == 24576, "edge stage layout");

// Narrow stencil ids (opt-in, kIds16): the ten ids of an edge are stored as 16-bit offsets from the edge's own id, tile-major
// ([tile][10][128] shorts = one 2560-byte bulk copy per tile instead of ten 512-byte rows). Along the space-filling-curve numbering
// 99 % of the offsets fit; a tile in which one does not is marked "wide" and its ids come from the ordinary int rows (per-tile flag,
// preloaded into shared memory so that the producer knows each tile's byte count without a global load). 180 B per edge instead of
// 200 for the narrow tiles; the arithmetic is untouched (the id only addresses the gather).
constexpr int kMaxTilesPerCta16 = 4096;
constexpr uint32_t kNarrowIdBytes = kStencil * kTile * sizeof(short);    // 2560

template <bool kIds16>
__device__ __forceinline__ void edge_step_pipe_body(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, int n_tiles,
                                                    const HaloInline& halo, const short* __restrict__ sid16,
                                                    const unsigned char* __restrict__ tile_wide) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    EdgeStage* stages = reinterpret_cast<EdgeStage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(EdgeStage));
    uint64_t* empty = full + kStages;
    unsigned char* wide_s = smem_raw + kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t);     // [my_tiles], kIds16 only
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (kIds16) {
        const int mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int i = threadIdx.x; i < mine; i += kPipeThreads) wide_s[i] = tile_wide[(size_t)blockIdx.x + (size_t)i * gridDim.x];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) {
            mbar_init(full + i, 1);               // one arrive (the producer's expect_tx) + the copies' bytes
            mbar_init(empty + i, kTile / 32);     // one arrive per consumer warp of the group that used the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t S = (size_t)t.stride;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            for (int i = 0; i < my_tiles; i++) {
                const int st = i % kStages;
                if (i >= kStages) mbar_wait(empty + st, ((i / kStages) - 1) & 1);
                EdgeStage* d = stages + st;
                const size_t e0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile;
                const bool narrow = kIds16 && wide_s[i] == 0;
                mbar_expect_tx(full + st, narrow ? kEdgeStageBytes - (uint32_t)sizeof(d->sid) + kNarrowIdBytes : kEdgeStageBytes);
                if (narrow) bulk_g2s(d->sid, sid16 + (e0 / kTile) * (size_t)(kStencil * kTile), kNarrowIdBytes, full + st, pol);
                else {
#pragma unroll
                    for (int j = 0; j < kStencil; j++) bulk_g2s(d->sid[j], t.sid + j * S + e0, kTile * 4, full + st, pol);
                }
#pragma unroll
                for (int j = 0; j < kStencil; j++) bulk_g2s(d->sw[j], t.sw + j * S + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->cells, t.cells + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->grad, t.grad + e0, kTile * 16, full + st, pol);
                bulk_g2s(d->dist, t.dist + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->fcor, t.fcor + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->own, s.vl_in + e0, kTile * 16, full + st, pol);
                bulk_g2s(d->h1, s.h1 + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->h2, s.h2 + e0, kTile * 8, full + st, pol);
            }
        }
        return;
    }
    // ---- consumers: group g takes this CTA's tiles g, g+kGroups, ... ----
    const int g = (warp - 1) / (kTile / 32);
    const int tl = (int)threadIdx.x - 32 - g * kTile;       // 0..127 within the tile
    double warp_energy = 0.0;                               // lane 0: this warp's tiles, in tile order
    for (int i = g; i < my_tiles; i += kGroups) {
        const int st = i % kStages;
        const EdgeStage* d = stages + st;
        const int e = (int)(((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile) + tl;
        // partitioned runs: the first tiles hold the boundary edges (they feed the neighbours)
        const bool bnd_tile = (e - tl) < halo.n_bnd;
        mbar_wait(full + st, (i / kStages) & 1);
        double e_area = 0.0;
        if (e < t.n_edges) {
            // gathers first (their addresses come from shared memory), arithmetic after
            double2 nb[kStencil];
            if (kIds16 && wide_s[i] == 0) {
                const short* off = reinterpret_cast<const short*>(d->sid);        // [10][128] offsets from e; 0 = empty slot (weight 0)
#pragma unroll
                for (int j = 0; j < kStencil; j++) nb[j] = ld_gather(s.vl_in + (e + (int)off[j * kTile + tl]));
            } else {
#pragma unroll
                for (int j = 0; j < kStencil; j++) {
                    const int id = d->sid[j][tl];
                    nb[j] = ld_gather(s.vl_in + (id < 0 ? e : id));
                }
            }
            const int2 c = d->cells[tl];
            const double2 in = ld_gather(s.eu + c.x), out = ld_gather(s.eu + c.y);
            const double dd = d->dist[tl], fc = d->fcor[tl];
            const double2 own = d->own[tl];
            double cor = 0.0, vt = 0.0;
            const double rd = __drcp_rn(dd);
#pragma unroll
            for (int j = 0; j < kStencil; j++) {
                const double w = d->sw[j][tl];
                const double coeff = exact_div(fc * w * nb[j].y, dd, rd);        // mesh.cpp:2881
                cor += coeff * nb[j].x;
                vt += nb[j].x * w * nb[j].y;                                     // interpolation.cpp:43
            }
            vt = exact_div(vt, dd, rd);
            e_area = dissipation_flux(p, own.x, vt) * (dd * own.y);
            const double2 G = d->grad[tl];
            const double grad = (-p.g * G.x) * in.x + (-p.g * G.y) * out.x;      // updateMomentum.cpp:42
            const double f0 = grad + cor;
            const double drag = (-p.alpha) * own.x + (G.x * in.y + G.y * out.y); // timeIntegrator.cpp:219
            const double f1 = d->h1[tl], f2 = d->h2[tl];
            double v = own.x + ab3_increment(f0, f1, f2, p.dt, mode);            // temporalOperators.cpp:41,55,64
            v += p.dt * drag;                                                    // timeIntegrator.cpp:242
            s.vl_out[e] = make_double2(v, own.y);
            if (mode == AB3_SECOND) s.h1[e] = f0;
            else s.h2[e] = f0;
            if (e < halo.n_bnd) {                          // into the neighbours' ghost slots (direct stores over NVLink)
                const int k1 = halo.send_first[e + 1];
                for (int k = halo.send_first[e]; k < k1; k++) halo.remote.data[halo.send_peer[k]][halo.send_remote[k]] = make_double2(v, own.y);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);          // this warp no longer reads the stage
        if (bnd_tile) {
            asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "n"(kTile) : "memory");   // the group's peer stores are issued
            if (tl == 0) halo_tile_done(halo, (unsigned int)((halo.n_bnd + kTile - 1) / kTile));
        }
        for (int o = 16; o > 0; o >>= 1) e_area += __shfl_down_sync(0xffffffffu, e_area, o);
        warp_energy += e_area;
    }
    // energy diagnostic (energy.cpp:36-40 sums serially): warp sums -> CTA sum -> the last CTA to finish adds the
    // CTA sums in index order. Every level has a fixed order, so the result is reproducible.
    constexpr int kConsumerWarps = kGroups * kTile / 32;
    __shared__ double warp_sums[kConsumerWarps];
    __shared__ bool is_last;
    if (lane == 0) warp_sums[warp - 1] = warp_energy;
    asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kTile) : "memory");      // consumers only (the producer warp has left)
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
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (s.ctl != nullptr) {                      // device-side step bookkeeping (StepCtl)
                const unsigned long long k = s.ctl->count;
                s.series[k] = acc;
                s.ctl->cur = s.scal[k];
                s.ctl->count = k + 1ull;
                if (halo.n_bnd > 0) s.ctl->epoch[0] += 1ull;    // every CTA is past its boundary tiles: the v exchange is published
            } else {
                *s.energy_out = acc;
            }
            *s.ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(kPipeThreads, 2) edge_step_pipe_kernel(EdgeTables t, Physics p, EdgeState s, int mode, int n_tiles,
                                                                          HaloInline halo) {
    edge_step_pipe_body<false>(t, p, s, mode, n_tiles, halo, nullptr, nullptr);
}
__global__ void __launch_bounds__(kPipeThreads, 2) edge_step_pipe16_kernel(EdgeTables t, Physics p, EdgeState s, int mode, int n_tiles,
                                                                            HaloInline halo, const short* sid16, const unsigned char* tile_wide) {
    edge_step_pipe_body<true>(t, p, s, mode, n_tiles, halo, sid16, tile_wide);
}

// ---------------------------------------------------------------- cell step ----
struct __align__(16) CellStage {
    int eid[kCellEdges][kTile];      // 3072 B
    double area[kTile];              // 1024 B
    double2 eu[kTile];               // 2048 B
    double h1[kTile];                // 1024 B
    double h2[kTile];                // 1024 B
    double trig[8][kTile];           // 8192 B   rows used depend on the potential
};
constexpr uint32_t kCellStageBytes = sizeof(CellStage);

struct TrigRows {                    // which rows of CellTables.trig / trig_sq a potential reads (tidalPotentials.cpp:80-172)
    int n;
    int row[8];                      // 0..7 = trig rows, 8 = cos^2 lat, 9 = sin^2 lat
};
__host__ __device__ inline TrigRows trig_rows_for(int potential) {
    switch (potential) {
        case P_ECC: return TrigRows{4, {8, 9, 6, 7, 0, 0, 0, 0}};
        case P_OBLIQ: return TrigRows{2, {5, 2, 0, 0, 0, 0, 0, 0}};
        case P_OBLIQ_WEST: return TrigRows{4, {0, 1, 2, 3, 0, 0, 0, 0}};
        case P_FULL: return TrigRows{6, {8, 9, 6, 7, 5, 2, 0, 0}};
        case P_FULL2: return TrigRows{8, {0, 1, 2, 3, 4, 6, 7, 8}};
        default: return TrigRows{0, {0, 0, 0, 0, 0, 0, 0, 0}};
    }
}

__device__ __forceinline__ double tidal_potential_rows(const Physics& p, const StepScalars& m, const double (*T)[kTile], int tl) {
    switch (p.potential) {
        case P_ECC: {
            const double cosSq = T[0][tl], sinSq = T[1][tl], cos2Lon = T[2][tl], sin2Lon = T[3][tl];
            return p.factor * ((1. - 3. * sinSq) * m.cosM + cosSq * (3. * m.cosM * cos2Lon + 4. * m.sinM * sin2Lon));
        }
        case P_OBLIQ: {
            const double sin2Lat = T[0][tl], cosLon = T[1][tl];
            return p.factor * m.cosM * sin2Lat * cosLon;
        }
        case P_OBLIQ_WEST: {
            const double cosLat = T[0][tl], sinLat = T[1][tl], cosLon = T[2][tl], sinLon = T[3][tl];
            return 3 * p.factor * sinLat * cosLat * (cosLon * m.cosM - sinLon * m.sinM);
        }
        case P_FULL: {
            const double cosSq = T[0][tl], sinSq = T[1][tl], cos2Lon = T[2][tl], sin2Lon = T[3][tl], sin2Lat = T[4][tl], cosLon = T[5][tl];
            return p.factor * ((1 - 3 * sinSq) * m.cosM + cosSq * (3 * m.cosM * cos2Lon + 4 * m.sinM * sin2Lon)) +
                   p.factor2 * m.cosM * sin2Lat * cosLon;
        }
        case P_FULL2: {
            const double cosLat = T[0][tl], sinLat = T[1][tl], cosLon = T[2][tl], sinLon = T[3][tl], cos2Lat = T[4][tl], cos2Lon = T[5][tl],
                         sin2Lon = T[6][tl], cosSq = T[7][tl];
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

__global__ void __launch_bounds__(kPipeThreads, 2) cell_step_pipe_kernel(CellTables t, Physics p, CellState s, int mode, StepScalars next,
                                                                          int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CellStage* stages = reinterpret_cast<CellStage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(CellStage));
    uint64_t* empty = full + kStages;
    __shared__ double red[kPipeThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, kTile / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one extra CTA (the last) only finishes the edge kernel's energy sum; the others pipeline tiles
    if (blockIdx.x == gridDim.x - 1) {
        if (s.energy_out == nullptr) return;
        double acc = 0.0;
        for (int i = threadIdx.x; i < s.n_energy_partials; i += kPipeThreads) acc += s.energy_partial[i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < kPipeThreads / 32; w++) tot += red[w];
            *s.energy_out = tot;
        }
        return;
    }
    const int n_workers = (int)gridDim.x - 1;
    const size_t S = (size_t)t.n_cells;                     // SoA stride (padded to the tile size by the host)
    const TrigRows rows = trig_rows_for(p.potential);
    const uint32_t stage_bytes = (uint32_t)(kCellEdges * kTile * 4 + kTile * 8 + kTile * 16 + 2 * kTile * 8 + rows.n * kTile * 8);
    const int my_tiles = (n_tiles - (int)blockIdx.x + n_workers - 1) / n_workers;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            for (int i = 0; i < my_tiles; i++) {
                const int st = i % kStages;
                if (i >= kStages) mbar_wait(empty + st, ((i / kStages) - 1) & 1);
                CellStage* d = stages + st;
                const size_t c0 = ((size_t)blockIdx.x + (size_t)i * n_workers) * kTile;
                mbar_expect_tx(full + st, stage_bytes);
#pragma unroll
                for (int j = 0; j < kCellEdges; j++) bulk_g2s(d->eid[j], t.eid + j * S + c0, kTile * 4, full + st, pol);
                bulk_g2s(d->area, t.area + c0, kTile * 8, full + st, pol);
                bulk_g2s(d->eu, s.eu_in + c0, kTile * 16, full + st, pol);
                bulk_g2s(d->h1, s.h1 + c0, kTile * 8, full + st, pol);
                bulk_g2s(d->h2, s.h2 + c0, kTile * 8, full + st, pol);
                for (int k = 0; k < rows.n; k++) {
                    const int r = rows.row[k];
                    const double* src = r < 8 ? t.trig + (size_t)r * S + c0 : t.trig_sq + (size_t)(r - 8) * S + c0;
                    bulk_g2s(d->trig[k], src, kTile * 8, full + st, pol);
                }
            }
        }
        return;
    }
    const int g = (warp - 1) / (kTile / 32);
    const int tl = (int)threadIdx.x - 32 - g * kTile;
    if (s.next_dev != nullptr) next = *s.next_dev;       // graph replay: time factors left by the edge kernel
    for (int i = g; i < my_tiles; i += kGroups) {
        const int st = i % kStages;
        const CellStage* d = stages + st;
        const int c = (int)(((size_t)blockIdx.x + (size_t)i * n_workers) * kTile) + tl;
        mbar_wait(full + st, (i / kStages) & 1);
        if (c < t.n_active) {
            int packed[kCellEdges];
            double2 ed[kCellEdges];
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) {
                packed[j] = d->eid[j][tl];
                ed[j] = ld_gather(s.vl + (packed[j] == -1 ? 0 : (packed[j] & 0x7fffffff)));
            }
            double2 stt = d->eu[tl];
            const double area = d->area[tl];
            double div = 0.0;                                                     // updateEta.cpp:39, mesh.cpp:3246
            const double ra = __drcp_rn(area);
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) {
                if (packed[j] != -1) {
                    const double ndir = (packed[j] < 0) ? 1.0 : -1.0;
                    const double coeff = exact_div(ndir * ed[j].y, area, ra);
                    div += (p.h * coeff) * ed[j].x;
                }
            }
            const double f0 = div;
            stt.x += ab3_increment(f0, d->h1[tl], d->h2[tl], p.dt, mode);
            s.hw[c] = f0;
            if (p.potential != P_NONE) stt.y = tidal_potential_rows(p, next, d->trig, tl);
            s.eu_out[c] = stt;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);
    }
}

}  // namespace

int pipe_tile() { return kTile; }

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

cudaError_t pipe_configure() {
    cudaError_t e = cudaFuncSetAttribute(edge_step_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t)));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(cell_step_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kStages * sizeof(CellStage) + 2 * kStages * sizeof(uint64_t)));
}

cudaError_t launch_edge_step_pipe(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                  cudaStream_t stream) {
    static bool configured_dev[64] = {false};      // the opt-in shared-memory size is a per-device attribute
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 63];
    const size_t smem = kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(edge_step_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int n_tiles = (t.n_edges + kTile - 1) / kTile;
    const int grid = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
    HaloInline none;
    none.n_bnd = 0;
    edge_step_pipe_kernel<<<grid, kPipeThreads, smem, stream>>>(t, p, s, mode, n_tiles, halo ? *halo : none);
    return cudaGetLastError();
}

bool edge_ids16_fits(int n_edges) {
    const int n_tiles = (n_edges + kTile - 1) / kTile;
    const int grid = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
    return grid > 0 && (n_tiles + grid - 1) / grid <= kMaxTilesPerCta16;
}

cudaError_t launch_edge_step_pipe16(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                    const short* sid16, const unsigned char* tile_wide, cudaStream_t stream) {
    static bool configured_dev[64] = {false};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 63];
    const int n_tiles = (t.n_edges + kTile - 1) / kTile;
    const int grid = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
    if (!edge_ids16_fits(t.n_edges)) return cudaErrorInvalidValue;
    const int per_cta = (n_tiles + grid - 1) / grid;
    const size_t smem_max = kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t) + kMaxTilesPerCta16;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(edge_step_pipe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const size_t smem = kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t) + (size_t)((per_cta + 15) / 16 * 16);
    HaloInline none;
    none.n_bnd = 0;
    edge_step_pipe16_kernel<<<grid, kPipeThreads, smem, stream>>>(t, p, s, mode, n_tiles, halo ? *halo : none, sid16, tile_wide);
    return cudaGetLastError();
}

cudaError_t launch_cell_step_pipe(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                                  cudaStream_t stream) {
    static bool configured_dev[64] = {false};      // the opt-in shared-memory size is a per-device attribute
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 63];
    const size_t smem = kStages * sizeof(CellStage) + 2 * kStages * sizeof(uint64_t);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(cell_step_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int n_tiles = (t.n_active + kTile - 1) / kTile;
    const int workers = n_tiles < 2 * num_sms() ? n_tiles : 2 * num_sms();
    cell_step_pipe_kernel<<<workers + 1, kPipeThreads, smem, stream>>>(t, p, s, mode, next, n_tiles);
    return cudaGetLastError();
}

}  // namespace odis
