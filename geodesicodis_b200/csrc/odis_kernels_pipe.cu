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
#include "odis_sh.cuh"

#include <cstdlib>
#include <mutex>

namespace odis {
namespace {

constexpr int kTile = 128;            // edges (cells) per tile = one consumer group of 4 warps
constexpr int kGroups = 2;            // consumer groups per CTA
constexpr int kStages = 4;            // tiles in flight per CTA
constexpr int kPipeThreads = 32 + kGroups * kTile;
// Partitioned launches of the edge kernel carry one more warp, the halo warp: it meets a consumer group at a named barrier when the
// group has issued the peer stores of a boundary tile and then does the system-scope fence + count (+ flags) of that tile, so that the
// group moves on to its next tile instead of sitting through the fence (measured on 2 and 4 B200s, round 2: up to 8 us per boundary
// tile, which ended the boundary CTAs 4-5 us after all others, every step).
constexpr int kHaloWarp = 1 + kGroups * kTile / 32;          // warp index of the halo warp
constexpr int kPipeThreadsHalo = kPipeThreads + 32;

// Launch of a per-step kernel. cooperative: the grid carries a grid-wide barrier, so all of its CTAs must be resident together — the
// cooperative attribute makes the driver schedule the grid as a whole (two such grids of different streams never interleave
// partially) and refuse a grid that cannot fit. (Programmatic dependent launch between the step's kernels was measured on a B200 in
// round 2: no gain under CUDA-graph replay, so it is not used.)
template <typename T> struct launch_identity { using type = T; };
template <typename... KArgs>
static cudaError_t launch_step_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool cooperative,
                                      typename launch_identity<KArgs>::type... args) {
#ifdef __CUDACC__
    if (cooperative) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    }
#endif
    kernel<<<grid, block, smem, stream>>>(args...);
    return cudaGetLastError();
}

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
// which the next kernel re-reads, out of the 126 MB L2. (Measured on a B200, round 2: without the hint on grids whose tables fit the L2
// as a whole — 40,962 and 163,842 cells — the step is no faster, 13.16 vs 13.21 us, resp. slower, 27.8 vs 26.1 us: those sizes are
// latency-bound, not DRAM-bound, so the hint stays unconditional.)
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
// the same without the evict-first hint: rows that the next kernel reads again (the harmonic basis inputs) should stay in L2
__device__ __forceinline__ void bulk_g2s_keep(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
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
        // ONE system-scope fence (cumulative over every tile's stores: each was fenced before its count), then the flags as relaxed
        // system-scope stores issued back to back. (A release store per neighbour, as until round 2, is a fence each: the second one
        // waits for the first flag's NVLink round trip, and so on — about 3 us per neighbour, serial, measured as +6 / +11 / +18 us
        // per step with 1 / 3 / 5 neighbours.)
        __threadfence_system();
        const unsigned long long epoch = ((volatile unsigned long long*)h.ctl->epoch)[0] + 1ull;
        for (int k = 0; k < h.n_peers; k++) {
            unsigned long long* f = h.remote.flags[k] + h.flag_slot;
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
        // the publisher — one thread per launch — also records the epoch for the cell kernel that follows (which waits for the
        // neighbours' flags to reach it). Not the CTA that happens to finish last: the halo warp publishes while the consumers carry on,
        // and on a small grid every CTA can be through its tiles before the fence above returns.
        ((volatile unsigned long long*)h.ctl->epoch)[0] = epoch;
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

#ifdef ODIS_TRACE
// Tuning aid (variant library only, build_variant("trace", ["ODIS_TRACE"])): per-CTA time stamps of the two staged kernels, written to a
// buffer handed over with odis_debug_trace_enable: [kernel 0 = edge, 1 = cell][launch % slots][CTA][8 events][2: globaltimer ns, clock64].
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned int g_trace_slots = 0, g_trace_ctas = 0;
__device__ unsigned int g_trace_arrivals[2] = {0u, 0u};
struct Trace {
    unsigned long long* row;      // this CTA's 8 x 2 entries of this launch, or nullptr
    __device__ __forceinline__ void begin(int kernel) {     // one thread per CTA
        row = nullptr;
        unsigned long long* base = g_trace;
        if (base == nullptr || blockIdx.x >= g_trace_ctas) return;
        const unsigned int launch = atomicAdd(&g_trace_arrivals[kernel], 1u) / gridDim.x;
        row = base + ((((size_t)kernel * g_trace_slots + launch % g_trace_slots) * g_trace_ctas + blockIdx.x) * 8) * 2;
    }
    __device__ __forceinline__ void mark(int ev) const {
        if (row == nullptr) return;
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        row[2 * ev] = g;
        row[2 * ev + 1] = (unsigned long long)clock64();
    }
};
#define ODIS_TRACE_DECL __shared__ Trace trace_s;
#define ODIS_TRACE_BEGIN(k) do { if (threadIdx.x == 0) { trace_s.begin(k); trace_s.mark(0); } } while (0)
#define ODIS_TRACE_MARK(cond, ev) do { if (cond) trace_s.mark(ev); } while (0)
#else
#define ODIS_TRACE_DECL
#define ODIS_TRACE_BEGIN(k) do { } while (0)
#define ODIS_TRACE_MARK(cond, ev) do { } while (0)
#endif

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
    ODIS_TRACE_DECL
    ODIS_TRACE_BEGIN(0);
    if (kIds16) {
        const int mine = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int i = threadIdx.x; i < mine; i += (int)blockDim.x) wide_s[i] = tile_wide[(size_t)blockIdx.x + (size_t)i * gridDim.x];
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
                // {v,l} of the tile's own edges only: behind the last own edge the array holds ghost slots, which the neighbours write
                // while this kernel runs (the lanes beyond n_edges are idle anyway)
                const int own_n = min(kTile, t.n_edges - (int)e0);
                const uint32_t own_bytes = (uint32_t)own_n * (uint32_t)sizeof(double2);
                mbar_expect_tx(full + st, (narrow ? kEdgeStageBytes - (uint32_t)sizeof(d->sid) + kNarrowIdBytes : kEdgeStageBytes) -
                                              (uint32_t)sizeof(d->own) + own_bytes);
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
                bulk_g2s(d->own, s.vl_in + e0, own_bytes, full + st, pol);
                bulk_g2s(d->h1, s.h1 + e0, kTile * 8, full + st, pol);
                bulk_g2s(d->h2, s.h2 + e0, kTile * 8, full + st, pol);
            }
        }
        return;
    }
    const bool halo_warp_on = blockDim.x > kPipeThreads;    // partitioned launch
    if (warp == kHaloWarp) {
        // ---- halo warp: the boundary tiles are the first tiles; tile i of this CTA belongs to group i % kGroups ----
        const int n_bnd_tiles = (halo.n_bnd + kTile - 1) / kTile;
        for (int i = 0; i < my_tiles && (int)blockIdx.x + i * (int)gridDim.x < n_bnd_tiles; i++) {
            asm volatile("bar.sync %0, %1;" ::"r"(4 + i % kGroups), "r"(kTile + 32) : "memory");   // the group's peer stores are issued
            if (lane == 0) halo_tile_done(halo, (unsigned int)n_bnd_tiles);
            __syncwarp();
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
        ODIS_TRACE_MARK(i == 0 && tl == 0, 1);                      // first tile: its rows have arrived
        double e_area = 0.0;
        // boundary edges: where the new value goes (CSR over the boundary edges), requested now so that the look-up overlaps the gathers
        int k0 = 0, k1 = 0, peer0 = 0, slot0 = 0;
        if (e < halo.n_bnd) {
            k0 = halo.send_first[e]; k1 = halo.send_first[e + 1];
            if (k1 > k0) { peer0 = halo.send_peer[k0]; slot0 = halo.send_remote[k0]; }
        }
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
            if (k1 > k0) {                                 // into the neighbours' ghost slots (direct stores over NVLink)
                halo.remote.data[peer0][slot0] = make_double2(v, own.y);
                for (int k = k0 + 1; k < k1; k++) halo.remote.data[halo.send_peer[k]][halo.send_remote[k]] = make_double2(v, own.y);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);          // this warp no longer reads the stage
        ODIS_TRACE_MARK(i == 0 && tl == 0, 2);                      // first tile: updated and stored
        if (bnd_tile) {
            if (halo_warp_on) {                           // hand the tile to the halo warp (it is waiting there) and move on
                asm volatile("bar.sync %0, %1;" ::"r"(4 + g), "r"(kTile + 32) : "memory");
            } else {
                asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "r"(kTile) : "memory");   // the group's peer stores are issued
                if (tl == 0) halo_tile_done(halo, (unsigned int)((halo.n_bnd + kTile - 1) / kTile));
            }
        }
        ODIS_TRACE_MARK(i == 0 && tl == 0, 3);                      // ... and counted done (boundary tile: fence + flag)
        for (int o = 16; o > 0; o >>= 1) e_area += __shfl_down_sync(0xffffffffu, e_area, o);
        warp_energy += e_area;
    }
    // energy diagnostic (energy.cpp:36-40 sums serially): warp sums -> CTA sum -> the last CTA to finish adds the
    // CTA sums in index order. Every level has a fixed order, so the result is reproducible.
    constexpr int kConsumerWarps = kGroups * kTile / 32;
    __shared__ double warp_sums[kConsumerWarps];
    __shared__ bool is_last;
    if (lane == 0) warp_sums[warp - 1] = warp_energy;
    ODIS_TRACE_MARK(tl == 0 && g == 0, 4);                          // group 0 through its tiles
    ODIS_TRACE_MARK(tl == 0 && g == 1, 5);                          // group 1 through its tiles
    asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kTile) : "memory");      // consumers only (the producer warp has left)
    if (threadIdx.x == 32) {
        double tot = 0.0;
        for (int w = 0; w < kConsumerWarps; w++) tot += warp_sums[w];
        s.block_partial[blockIdx.x] = tot;
        __threadfence();
        is_last = (atomicAdd(s.ticket, 1u) == gridDim.x - 1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kTile) : "memory");
    ODIS_TRACE_MARK(threadIdx.x == 32, 6);                          // CTA done (but for the last CTA's sum)
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
            } else {
                *s.energy_out = acc;
            }
            *s.ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(kPipeThreadsHalo, 2) edge_step_pipe_kernel(EdgeTables t, Physics p, EdgeState s, int mode, int n_tiles,
                                                                          HaloInline halo) {
    edge_step_pipe_body<false>(t, p, s, mode, n_tiles, halo, nullptr, nullptr);
}
__global__ void __launch_bounds__(kPipeThreadsHalo, 2) edge_step_pipe16_kernel(EdgeTables t, Physics p, EdgeState s, int mode, int n_tiles,
                                                                            HaloInline halo, const short* sid16, const unsigned char* tile_wide) {
    edge_step_pipe_body<true>(t, p, s, mode, n_tiles, halo, sid16, tile_wide);
}

// ---------------------------------------------------------------- cell step ----
// Persistent, warp-specialised cell update: one producer lane streams the per-cell rows of the next tiles into shared memory with
// cp.async.bulk (kS stages), kG consumer groups of 128 threads each take whole tiles. A consumer copies its table values out of the
// stage into registers and releases the stage BEFORE it issues the six {v,l} gathers, so the producer refills while the gathers are in
// flight. Same arithmetic, operation for operation, as cell_step_kernel (odis_kernels.cu): bit-identical eta and tendencies.
//
// Harmonic analysis folded in (LSG >= 2; self-gravity term): every consumer thread keeps the (LSG+1)^2 sums
// b_k += Y_k(i) eta_i^{n+1} of ITS cells in registers across all of its tiles (basis values rebuilt from cos/sin of latitude and
// longitude by the recurrences of sh_analysis_mf_kernel, odis_sh.cu) and the CTA reduces ONCE at the end: butterfly per warp, warps in
// order, one partial per row and CTA; the last CTA to finish (ticket) adds the CTAs' partials in CTA order and leaves this rank's b.
// Every level has a fixed order, so the sums do not depend on scheduling. Partitioned solvers: that last CTA also publishes the rank's
// sums for the all-reduce through peer memory which the synthesis kernel completes (protocol of sh_reduce_publish_kernel, odis_sh.cu).
struct __align__(16) CellStage {
    int eid[kCellEdges][kTile];      //  3072 B
    double area[kTile];              //  1024 B
    double2 eu[kTile];               //  2048 B
    double h1[kTile];                //  1024 B
    double h2[kTile];                //  1024 B
    double trig[kCellMaxRows][kTile];  // 10240 B   rows staged depend on the potential (+ the four basis rows with LSG)
};
constexpr uint32_t kCellFixedBytes = kCellEdges * kTile * 4 + kTile * 8 + kTile * 16 + 2 * kTile * 8;

struct TrigRows {                    // which rows of CellTables.trig / trig_sq a potential reads (tidalPotentials.cpp:80-172), in the
    int n;                           // order tidal_potential_in takes them
    int row[8];                      // 0..7 = trig rows, 8 = cos^2 lat, 9 = sin^2 lat
};
inline TrigRows trig_rows_for(int potential) {
    switch (potential) {
        case P_ECC: return TrigRows{4, {8, 9, 6, 7, 0, 0, 0, 0}};
        case P_OBLIQ: return TrigRows{2, {5, 2, 0, 0, 0, 0, 0, 0}};
        case P_OBLIQ_WEST: return TrigRows{4, {0, 1, 2, 3, 0, 0, 0, 0}};
        case P_FULL: return TrigRows{6, {8, 9, 6, 7, 5, 2, 0, 0}};
        case P_FULL2: return TrigRows{8, {0, 1, 2, 3, 4, 6, 7, 8}};
        default: return TrigRows{0, {0, 0, 0, 0, 0, 0, 0, 0}};
    }
}

// tidal potential from its inputs in trig_rows_for() order; expression shapes of tidalPotentials.cpp:80-172 (= tidal_potential(),
// odis_kernels.cu)
template <int kPot>
__device__ __forceinline__ double tidal_potential_in(const Physics& p, const StepScalars& m, const double* in) {
    switch (kPot >= 0 ? kPot : p.potential) {
        case P_ECC: {
            const double cosSq = in[0], sinSq = in[1], cos2Lon = in[2], sin2Lon = in[3];
            return p.factor * ((1. - 3. * sinSq) * m.cosM + cosSq * (3. * m.cosM * cos2Lon + 4. * m.sinM * sin2Lon));
        }
        case P_OBLIQ: {
            const double sin2Lat = in[0], cosLon = in[1];
            return p.factor * m.cosM * sin2Lat * cosLon;
        }
        case P_OBLIQ_WEST: {
            const double cosLat = in[0], sinLat = in[1], cosLon = in[2], sinLon = in[3];
            return 3 * p.factor * sinLat * cosLat * (cosLon * m.cosM - sinLon * m.sinM);
        }
        case P_FULL: {
            const double cosSq = in[0], sinSq = in[1], cos2Lon = in[2], sin2Lon = in[3], sin2Lat = in[4], cosLon = in[5];
            return p.factor * ((1 - 3 * sinSq) * m.cosM + cosSq * (3 * m.cosM * cos2Lon + 4 * m.sinM * sin2Lon)) +
                   p.factor2 * m.cosM * sin2Lat * cosLon;
        }
        case P_FULL2: {
            const double cosLat = in[0], sinLat = in[1], cosLon = in[2], sinLon = in[3], cos2Lat = in[4], cos2Lon = in[5],
                         sin2Lon = in[6], cosSq = in[7];
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

// recurrence coefficients of the normalised Legendre functions (layout of sh_recurrence_table()), this translation unit's copy
__constant__ double c_cell_rec[kShRecDoubles];
struct CellRec {
    __device__ __forceinline__ double a(int m, int l) const { return c_cell_rec[l * kShRecStride + m]; }
    __device__ __forceinline__ double b(int m, int l) const { return c_cell_rec[kShRecStride * kShRecStride + l * kShRecStride + m]; }
    __device__ __forceinline__ double sect(int m) const { return c_cell_rec[2 * kShRecStride * kShRecStride + m]; }
    __device__ __forceinline__ double first(int m) const { return c_cell_rec[2 * kShRecStride * kShRecStride + kShRecStride + m]; }
};
__device__ __forceinline__ double warp_sum_all(double x) {      // butterfly: fixed association, every lane gets the sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double2 ld_gather_cg(const double2* p) {     // ghost slots are written by other GPUs while the kernel runs
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
// the same by a whole warp: lane k polls neighbour k (the waits overlap instead of adding up), then the warp joins
__device__ __forceinline__ void halo_wait_warp(const HaloWait& w, StepCtl* ctl) {
    const int lane = threadIdx.x & 31;
    if (lane < w.n_peers) {
        const unsigned long long ev = ((volatile unsigned long long*)ctl->epoch)[0];
        const long long limit = ctl->spin_cycles > 0 ? ctl->spin_cycles : kHaloSpinCycles;
        const long long t0 = clock64();
        unsigned long long seen;
        const unsigned long long* fv = w.flag[lane];
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fv) : "memory");
        } while (seen < ev && clock64() - t0 < limit);
        if (seen < ev) ctl->pad = 1ull;
    }
    __syncwarp();
}
// the synthesis of one cell: sum over degrees 2..LT of s_k Y_k (recurrences of sh_synthesis_mf_kernel, odis_sh.cu)
template <int LT>
__device__ __forceinline__ double cell_synthesis(const double* ssh, double u, double z, double c1, double s1) {
    const CellRec rc;
    double cm = 1.0, sn = 0.0, pmm = 1.0, acc = 0.0;
#pragma unroll
    for (int m = 0; m <= LT; m++) {
        if (m > 0) {
            const double cn = __fma_rn(cm, c1, -(sn * s1));
            sn = __fma_rn(sn, c1, cm * s1);
            cm = cn;
            pmm = rc.sect(m) * u * pmm;
        }
        double p1 = pmm, p2 = 0.0;
#pragma unroll
        for (int l = m; l <= LT; l++) {
            if (l > m) {
                const double a = l == m + 1 ? rc.first(m) : rc.a(m, l);
                const double b = l == m + 1 ? 0.0 : rc.b(m, l);
                const double pn = a * __fma_rn(z, p1, -(b * p2));
                p2 = p1; p1 = pn;
            }
            if (l >= 2) {
                const int row = l * l + (m ? 2 * m - 1 : 0);
                const double q = m ? __fma_rn(cm, ssh[row], sn * ssh[row + 1]) : ssh[row];
                acc = __fma_rn(p1, q, acc);
            }
        }
    }
    return acc;
}

// Grid-wide barrier of the merged kernel (cooperative launch: every CTA is resident). bar[0] counts arrivals, bar[1] is the
// generation; the caller is the CTA's one arriving thread. Returns true in the LAST CTA to arrive, which must call
// grid_barrier_release afterwards; the others have waited (bounded) for that release when they return.
__device__ __forceinline__ bool grid_barrier_arrive(unsigned int* bar, unsigned int gen0, StepCtl* ctl) {
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1u) return true;
    const long long t0 = clock64(), limit = (ctl != nullptr && ctl->spin_cycles > 0) ? ctl->spin_cycles : kHaloSpinCycles;
    unsigned int gen;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    } while (gen == gen0 && clock64() - t0 < limit);
    if (gen == gen0 && ctl != nullptr) ctl->pad = 1ull;      // reported by the host (ODIS_ERR_STATE), no hang
    return false;
}
__device__ __forceinline__ void grid_barrier_release(unsigned int* bar, unsigned int gen0) {
    bar[0] = 0u;
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen0 + 1u) : "memory");
}

// kPot >= 0: the kernel is compiled for that one potential (P_ECC: the Enceladus / Europa case) — only its four inputs are ever live,
// which keeps the harmonic accumulators of the self-gravity variants in registers; -1: any potential, selected at run time.
template <int kG, int kS, int LSG, bool kPart, bool kMerged, int kPot>
__device__ __forceinline__ void cell_step_pipe_body(const CellTables& t, const Physics& p, const CellState& s, int mode, StepScalars next,
                                                    int n_tiles, const CellRows& rows, const HaloInline& halo, const CellSgAccum& sg,
                                                    const ShExchange& x) {
    static_assert(!(kMerged && kPart), "partitioned solvers keep solve + synthesis in a launch of their own (sh_bsolve_synthesis_mf_kernel)");
    constexpr int kThreads = 32 + kG * kTile;
    constexpr int kConsumerWarps = kG * kTile / 32;
    constexpr int kRows = LSG > 0 ? (LSG + 1) * (LSG + 1) : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CellStage* stages = reinterpret_cast<CellStage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kS * sizeof(CellStage));
    uint64_t* empty = full + kS;
    __shared__ StepScalars nx;                               // time factors of the potential: operands from shared memory, not 12 registers
    __shared__ double ginv_s[kMerged ? kRows * kRows : 1];
    __shared__ double fac_s[kMerged ? kRows : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ODIS_TRACE_DECL
    ODIS_TRACE_BEGIN(1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kS; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, kTile / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        nx = s.next_dev != nullptr ? *s.next_dev : next;     // graph replay: the factors were left by the edge kernel (StepCtl::cur)
    }
    if (kMerged) {                                           // fixed operator of the harmonic solve: in shared memory before it is needed
        for (int k = threadIdx.x; k < kRows * kRows; k += kThreads) ginv_s[k] = sg.ginv[k];
        for (int k = threadIdx.x; k < kRows; k += kThreads) fac_s[k] = sg.factor[k];
    }
    __syncthreads();
    const size_t S = (size_t)t.n_cells;                     // SoA stride (padded to the tile size by the host)
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            const uint32_t stage_bytes = kCellFixedBytes + (uint32_t)rows.n * kTile * 8;
            for (int i = 0; i < my_tiles; i++) {
                const int st = i % kS;
                if (i >= kS) mbar_wait(empty + st, ((i / kS) - 1) & 1);
                CellStage* d = stages + st;
                const size_t c0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile;
                mbar_expect_tx(full + st, stage_bytes);
#pragma unroll
                for (int j = 0; j < kCellEdges; j++) bulk_g2s(d->eid[j], t.eid + j * S + c0, kTile * 4, full + st, pol);
                bulk_g2s(d->area, t.area + c0, kTile * 8, full + st, pol);
                bulk_g2s(d->eu, s.eu_in + c0, kTile * 16, full + st, pol);
                bulk_g2s(d->h1, s.h1 + c0, kTile * 8, full + st, pol);
                bulk_g2s(d->h2, s.h2 + c0, kTile * 8, full + st, pol);
#pragma unroll
                for (int k = 0; k < kCellMaxRows; k++) {
                    if (k < rows.n) {
                        const int r = (int)(rows.src >> (4 * k)) & 15;
                        const double* src = r < 8 ? t.trig + (size_t)r * S + c0 : t.trig_sq + (size_t)(r - 8) * S + c0;
                        // the four basis rows (stage rows 0..3 with LSG) are read again by the synthesis launch that follows
                        if (LSG > 0 && k < 4) bulk_g2s_keep(d->trig[k], src, kTile * 8, full + st);
                        else bulk_g2s(d->trig[k], src, kTile * 8, full + st, pol);
                    }
                }
            }
        }
        return;
    }
    // ---- consumers: group g takes this CTA's tiles g, g+kG, ... ----
    const int g = (warp - 1) / (kTile / 32);
    const int tl = (int)threadIdx.x - 32 - g * kTile;
    double acc[kRows];
#pragma unroll
    for (int k = 0; k < kRows; k++) acc[k] = 0.0;
    bool halo_seen = false;
    for (int i = g; i < my_tiles; i += kG) {
        const int st = i % kS;
        const CellStage* d = stages + st;
        const int c0 = (int)(((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile);
        const int c = c0 + tl;
        // partitioned runs: the last tiles hold the cells that read ghost edges (own boundary cells, then the ghost cells)
        const bool bnd_tile = kPart && (c0 + kTile > halo.wait_from);
        mbar_wait(full + st, (i / kS) & 1);
        ODIS_TRACE_MARK(i == 0 && tl == 0, 1);                      // first tile: its rows have arrived
        int packed[kCellEdges];
#pragma unroll
        for (int j = 0; j < kCellEdges; j++) packed[j] = d->eid[j][tl];
        if (bnd_tile && !halo_seen) {                    // flags only grow: one wait per warp and launch
            ODIS_TRACE_MARK(tl == 0, 2 + 2 * g);                    // group g: before / after the wait for the neighbours' flags
            halo_wait_warp(halo.wait_v, halo.ctl);
            ODIS_TRACE_MARK(tl == 0, 3 + 2 * g);
            halo_seen = true;
        }
        // the six {v,l} gathers go out first; the rest of the stage is read (and the stage released) while they are in flight. A cell's
        // edges fill the slots from 0 (ascending reference id), so only slot 5 can be empty (the 12 pentagons); padded cells gather nothing.
        double2 ed[kCellEdges];
#pragma unroll
        for (int j = 0; j < kCellEdges; j++) ed[j] = make_double2(0.0, 0.0);
        if (c < t.n_active) {
            const int last = packed[kCellEdges - 1] == -1 ? 0 : (packed[kCellEdges - 1] & 0x7fffffff);
            if (!bnd_tile) {
#pragma unroll
                for (int j = 0; j < kCellEdges - 1; j++) ed[j] = ld_gather(s.vl + (packed[j] & 0x7fffffff));
                ed[kCellEdges - 1] = ld_gather(s.vl + last);
            } else {
#pragma unroll
                for (int j = 0; j < kCellEdges - 1; j++) ed[j] = ld_gather_cg(s.vl + (packed[j] & 0x7fffffff));
                ed[kCellEdges - 1] = ld_gather_cg(s.vl + last);
            }
        }
        double2 stt = d->eu[tl];
        const double area = d->area[tl], f1 = d->h1[tl], f2 = d->h2[tl];
        if (kPot >= 0 || p.potential != P_NONE) {
            constexpr int kIn = kPot == P_ECC ? 4 : 8;
            double in[kIn];
#pragma unroll
            for (int k = 0; k < kIn; k++) {
                const int slot = (int)(rows.pot >> (5 * k)) & 31;        // stage row of the potential's k-th input; bit 4: its square
                const double v = (k < rows.n_pot) ? d->trig[slot & 15][tl] : 0.0;
                in[k] = (slot & 16) ? v * v : v;                          // cos^2 / sin^2 lat from the basis rows (the product of mesh.cpp:2144-2145)
            }
            stt.y = tidal_potential_in<kPot>(p, nx, in);
        }
        double bu = 0.0, bz = 0.0, bc1 = 0.0, bs1 = 0.0;
        if (LSG > 0) {
            bu = d->trig[rows.basis & 15][tl]; bz = d->trig[(rows.basis >> 4) & 15][tl];
            bc1 = d->trig[(rows.basis >> 8) & 15][tl]; bs1 = d->trig[(rows.basis >> 12) & 15][tl];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);          // this warp no longer reads the stage
        double e_fit = 0.0;
        if (c < t.n_active) {
            double div = 0.0;                                                     // updateEta.cpp:39, mesh.cpp:3246
            const double ra = __drcp_rn(area);
#pragma unroll
            for (int j = 0; j < kCellEdges; j++) {
                if (j < kCellEdges - 1 || packed[j] != -1) {                      // the 12 pentagons have 5 edges
                    const double ndir = (packed[j] < 0) ? 1.0 : -1.0;             // -dir: dir = -1 for the outer cell
                    const double coeff = exact_div(ndir * ed[j].y, area, ra);
                    div += (p.h * coeff) * ed[j].x;
                }
            }
            const double f0 = div;
            stt.x += ab3_increment(f0, f1, f2, p.dt, mode);
            s.hw[c] = f0;
            s.eu_out[c] = stt;
            if (c < sg.n_fit) e_fit = stt.x;
        }
        if (LSG > 0) {
            // this cell's share of b = Y eta^{n+1}; padded / unfitted cells carry e = 0 and finite (zero) basis inputs
            const CellRec rc;
            double cm = 1.0, sn = 0.0, pmm = 1.0;
#pragma unroll
            for (int m = 0; m <= LSG; m++) {
                if (m > 0) {
                    const double cn = __fma_rn(cm, bc1, -(sn * bs1));       // cos(m lon), sin(m lon) by rotation
                    sn = __fma_rn(sn, bc1, cm * bs1);
                    cm = cn;
                    pmm = rc.sect(m) * bu * pmm;
                }
                const double ec = e_fit * cm, es = e_fit * sn;
                double p1 = pmm, p2 = 0.0;
#pragma unroll
                for (int l = m; l <= LSG; l++) {
                    if (l > m) {
                        const double a = l == m + 1 ? rc.first(m) : rc.a(m, l);
                        const double b = l == m + 1 ? 0.0 : rc.b(m, l);
                        const double pn = a * __fma_rn(bz, p1, -(b * p2));
                        p2 = p1; p1 = pn;
                    }
                    const int row = l * l + (m ? 2 * m - 1 : 0);
                    acc[row] = __fma_rn(p1, ec, acc[row]);
                    if (m > 0) acc[row + 1] = __fma_rn(p1, es, acc[row + 1]);
                }
            }
        }
    }
    ODIS_TRACE_MARK(tl == 0 && g == 0, 6);                          // group 0 through its tiles
    ODIS_TRACE_MARK(tl == 0 && g == 1, 7);
    if (LSG == 0) return;
    // ---- the CTA's harmonic sums: once per CTA ----
    __shared__ double red[kRows * kConsumerWarps];
    __shared__ double bsh[kRows];
    __shared__ double ssh[kRows];
    __shared__ bool is_last;
#pragma unroll
    for (int k = 0; k < kRows; k++) {
        const double tot = warp_sum_all(acc[k]);
        if (lane == 0) red[k * kConsumerWarps + (warp - 1)] = tot;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");      // consumers only (the producer warp has left)
    const int ct = (int)threadIdx.x - 32;
    unsigned int gen0 = 0u;
    if (kMerged && ct == 0) gen0 = *(volatile unsigned int*)(sg.bar + 1);      // before this CTA arrives, so before anyone can release
    if (ct < kRows) {
        double a = red[ct * kConsumerWarps];
#pragma unroll
        for (int q = 1; q < kConsumerWarps; q++) a = a + red[ct * kConsumerWarps + q];
        sg.cta_partial[(size_t)ct * sg.cta_stride + blockIdx.x] = a;
    }
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
    if (ct == 0) is_last = kMerged ? grid_barrier_arrive(sg.bar, gen0, halo.ctl) : (atomicAdd(sg.ticket, 1u) == gridDim.x - 1);
    asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
    unsigned long long epoch = 0ull;
    if (kPart) epoch = x.ctl[0] + (is_last ? 1ull : 0ull);           // the last CTA publishes the next epoch (and records it below)
    if (is_last) {
        // the last CTA to arrive adds the CTAs' partials in CTA order: all of a lane's loads in flight together, then the butterfly
        __threadfence();
        // partitioned: this rank's sums are collected in shared memory and PUSHED as LL lines (odis_sh.cuh) into slot [parity][rank] of
        // every rank's exchange block, its own included — peer stores, no flag, no fence: the readers poll the lines in their own memory
        double* bdst = kPart ? bsh : sg.b_out;
        constexpr int kMaxPerLane = 10;                               // grid <= 2 x 148 CTAs (cell_pipe_grid)
        for (int k = warp - 1; k < kRows; k += kConsumerWarps) {
            const double* row = sg.cta_partial + (size_t)k * sg.cta_stride;
            double v[kMaxPerLane];
#pragma unroll
            for (int q = 0; q < kMaxPerLane; q++) v[q] = (lane + 32 * q < (int)gridDim.x) ? __ldcg(row + lane + 32 * q) : 0.0;
            double a = v[0];
#pragma unroll
            for (int q = 1; q < kMaxPerLane; q++) a = a + v[q];
            for (int q = lane + 32 * kMaxPerLane; q < (int)gridDim.x; q += 32) a = a + __ldcg(row + q);
            a = warp_sum_all(a);
            if (lane == 0) bdst[k] = a;
        }
        if (!kMerged && ct == 0) *sg.ticket = 0u;
        if (kPart) {
            asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
            for (int q = ct; q < x.world * kRows; q += kG * kTile) {
                const int r = q / kRows, k = q - r * kRows;
                sh_ll_store(sh_ll_lines(x.block[r], (int)(epoch & 1ull)) + (size_t)x.rank * kShXSlot + k, bsh[k], (unsigned int)epoch);
            }
            if (ct == 0) x.ctl[0] = epoch;
        }
        if (kMerged) {
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
            if (ct == 0) grid_barrier_release(sg.bar, gen0);
        }
    }
    if (!kMerged) return;
    // ---- merged (unpartitioned solvers): every CTA is past the grid barrier, the sums are complete. Solve, synthesis of the own tiles ----
    if (ct < kRows) bsh[ct] = __ldcg(sg.b_out + ct);
    asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
    for (int j = warp - 1; j < kRows; j += kConsumerWarps) {   // s_j = (g factor_j) sum_k Ginv[j][k] b_k, order of solve_rows (odis_sh.cu)
        const double f = fac_s[j];
        double a = 0.0;
        if (f != 0.0)
            for (int k = lane; k < kRows; k += 32) a = a + ginv_s[j * kRows + k] * bsh[k];
        a = warp_sum_all(a);
        if (lane == 0) {
            ssh[j] = (sg.g * f) * a;
            if (blockIdx.x == 0) sg.s_out[j] = ssh[j];
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kG * kTile) : "memory");
    // U_i += sum_k s_k Y_k(i) on the tiles this thread updated above ({eta,U} as it wrote them; the basis rows from L2), two tiles per trip
    const double* T = t.trig;
    for (int i = g; i < my_tiles; i += 2 * kG) {
        const int ca = (int)(((size_t)blockIdx.x + (size_t)i * gridDim.x) * kTile) + tl;
        const int cb = (i + kG < my_tiles) ? (int)(((size_t)blockIdx.x + (size_t)(i + kG) * gridDim.x) * kTile) + tl : t.n_active;
        const bool oka = ca < t.n_active, okb = cb < t.n_active;
        const int xa = oka ? ca : 0, xb = okb ? cb : 0;
        const double ua = T[xa], za = T[S + xa], c1a = T[2 * S + xa], s1a = T[3 * S + xa];
        const double ub = T[xb], zb = T[S + xb], c1b = T[2 * S + xb], s1b = T[3 * S + xb];
        const double2 sta = s.eu_out[xa], stb = s.eu_out[xb];
        if (oka) s.eu_out[ca] = make_double2(sta.x, sta.y + cell_synthesis<LSG>(ssh, ua, za, c1a, s1a));
        if (okb) s.eu_out[cb] = make_double2(stb.x, stb.y + cell_synthesis<LSG>(ssh, ub, zb, c1b, s1b));
    }
}

// 288 threads (one producer warp, two consumer groups of 128), 4 stages, 2 CTAs per SM: the edge kernel's shape. Measured against a
// deeper pipeline (6 stages: +1.3 us) and one wide CTA per SM (3 or 4 groups, 6-8 stages: +8.5 us) at 655,362 cells, round 2.
constexpr int kCellGroups = 2, kCellStages = 4;
template <int LSG, bool kPart, bool kMerged, int kPot>
__global__ void __launch_bounds__(32 + kCellGroups * kTile, 2) cell_step_pipe_kernel(CellTables t, Physics p, CellState s, int mode, StepScalars next,
                                                                                      int n_tiles, CellRows rows, HaloInline halo, CellSgAccum sg,
                                                                                      ShExchange x) {
    cell_step_pipe_body<kCellGroups, kCellStages, LSG, kPart, kMerged, kPot>(t, p, s, mode, next, n_tiles, rows, halo, sg, x);
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

constexpr size_t kCellSmem = kCellStages * sizeof(CellStage) + 2 * kCellStages * sizeof(uint64_t);
template <int LSG, bool kPart>
static cudaError_t cell_pipe_attrs() {
    cudaError_t e = cudaFuncSetAttribute(cell_step_pipe_kernel<LSG, kPart, false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCellSmem);
    if constexpr (LSG == 0) return e;
    else {
        if (e != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(cell_step_pipe_kernel<LSG, kPart, false, (int)P_ECC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCellSmem)) != cudaSuccess) return e;
        if constexpr (kPart) return e;
        else {
            if ((e = cudaFuncSetAttribute(cell_step_pipe_kernel<LSG, false, true, (int)P_ECC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCellSmem)) != cudaSuccess) return e;
            return cudaFuncSetAttribute(cell_step_pipe_kernel<LSG, false, true, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCellSmem);
        }
    }
}

// once per device, before any stream capture: dynamic shared-memory sizes of the staged kernels, recurrence coefficients of the
// harmonic basis into this translation unit's constant memory (another solver on the same GPU may have kernels in flight that read
// them, and the legacy stream of cudaMemcpyToSymbol does not wait for the solvers' non-blocking streams — hence once, not per solver)
cudaError_t pipe_configure() {
    static std::mutex once;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(once);
    if (done[dev & 63]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(edge_step_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kStages * sizeof(EdgeStage) + 2 * kStages * sizeof(uint64_t)));
    if (e != cudaSuccess) return e;
    if ((e = cell_pipe_attrs<0, false>()) != cudaSuccess || (e = cell_pipe_attrs<0, true>()) != cudaSuccess ||
        (e = cell_pipe_attrs<2, false>()) != cudaSuccess || (e = cell_pipe_attrs<2, true>()) != cudaSuccess ||
        (e = cell_pipe_attrs<3, false>()) != cudaSuccess || (e = cell_pipe_attrs<3, true>()) != cudaSuccess ||
        (e = cell_pipe_attrs<4, false>()) != cudaSuccess || (e = cell_pipe_attrs<4, true>()) != cudaSuccess)
        return e;
    double rec[kShRecDoubles];
    sh_recurrence_table(rec);
    if ((e = cudaMemcpyToSymbol(c_cell_rec, rec, sizeof rec)) != cudaSuccess) return e;
    done[dev & 63] = true;
    return cudaSuccess;
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
    return launch_step_kernel(edge_step_pipe_kernel, grid, halo ? kPipeThreadsHalo : kPipeThreads, smem, stream, false, t, p, s, mode, n_tiles, halo ? *halo : none);
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
    return launch_step_kernel(edge_step_pipe16_kernel, grid, halo ? kPipeThreadsHalo : kPipeThreads, smem, stream, false, t, p, s, mode, n_tiles, halo ? *halo : none, sid16, tile_wide);
}

CellRows cell_rows_for(int potential, bool with_basis) {
    CellRows r;
    r.n = 0; r.n_pot = 0; r.src = 0ull; r.pot = 0ull; r.basis = 0u;
    int src[kCellMaxRows];
    auto slot_of = [&](int row) {
        for (int k = 0; k < r.n; k++)
            if (src[k] == row) return k;
        src[r.n] = row;
        return r.n++;
    };
    int basis[4] = {0, 0, 0, 0};
    if (with_basis)
        for (int k = 0; k < 4; k++) basis[k] = slot_of(k);            // cos lat, sin lat, cos lon, sin lon
    const TrigRows pr = trig_rows_for(potential);
    r.n_pot = pr.n;
    for (int k = 0; k < pr.n; k++) {
        const int row = pr.row[k];
        // cos^2 / sin^2 of the latitude are the squares of rows 0 / 1 (built as cos(lat)*cos(lat), odis_engine.cu / mesh.cpp:2144-2145):
        // with the basis rows staged anyway, the product is formed in the kernel instead of streaming 16 more bytes per cell
        const int slot = (with_basis && row >= 8) ? (basis[row - 8] | 16) : slot_of(row);
        r.pot |= (unsigned long long)slot << (5 * k);
    }
    for (int k = 0; k < r.n; k++) r.src |= (unsigned long long)src[k] << (4 * k);
    for (int k = 0; k < 4; k++) r.basis |= (unsigned)basis[k] << (4 * k);
    return r;
}

bool cell_pipe_supports_sg(int l_max) { return l_max >= 2 && l_max <= 4; }

int cell_pipe_grid(int n_cells) {
    const int n_tiles = (n_cells + kTile - 1) / kTile;
    const int cap = 2 * num_sms();                          // the widest configuration (two CTAs per SM)
    return n_tiles < cap ? (n_tiles > 0 ? n_tiles : 1) : cap;
}

// The merged kernel (cell update + grid barrier + solve + synthesis) needs every CTA resident: cooperative launch. Off with
// ODIS_B200_MERGED_SYNTH=0 (A/B timing) and in host emulation builds unless the emulation runs all CTAs of a launch at once.
bool cell_pipe_merged() {
#ifdef __CUDACC__
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("ODIS_B200_MERGED_SYNTH");
        v = (e && std::atoi(e) == 0) ? 0 : 1;
    }
    return v != 0;
#else
    // host emulation: only when it runs every CTA of a launch at once (tests/simt: ODIS_EMU_CTA_THREADS >= the grid) and says so
    const char* e = std::getenv("ODIS_EMU_MERGED");
    return e && std::atoi(e) != 0;
#endif
}

template <int LSG, bool kPart>
static cudaError_t launch_cell_cfg(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next, int n_tiles,
                                   const CellRows& rows, const HaloInline& halo, const CellSgAccum& sg, const ShExchange& x, cudaStream_t stream) {
    const int cap = 2 * num_sms(), grid = n_tiles < cap ? n_tiles : cap;
    constexpr int kBlock = 32 + kCellGroups * kTile;
    constexpr int kEcc = LSG > 0 ? (int)P_ECC : -1;          // the one-potential build exists for the self-gravity variants (register pressure)
    const bool ecc = LSG > 0 && p.potential == P_ECC;
    if constexpr (LSG > 0 && !kPart) {        // merged solve + synthesis: unpartitioned solvers only (cooperative launch: all CTAs resident)
        if (sg.merged) {
            if (ecc) return launch_step_kernel(cell_step_pipe_kernel<LSG, false, true, kEcc>, grid, kBlock, kCellSmem, stream, true, t, p, s, mode, next, n_tiles, rows, halo, sg, x);
            return launch_step_kernel(cell_step_pipe_kernel<LSG, false, true, -1>, grid, kBlock, kCellSmem, stream, true, t, p, s, mode, next, n_tiles, rows, halo, sg, x);
        }
    }
    if (ecc) return launch_step_kernel(cell_step_pipe_kernel<LSG, kPart, false, kEcc>, grid, kBlock, kCellSmem, stream, false, t, p, s, mode, next, n_tiles, rows, halo, sg, x);
    return launch_step_kernel(cell_step_pipe_kernel<LSG, kPart, false, -1>, grid, kBlock, kCellSmem, stream, false, t, p, s, mode, next, n_tiles, rows, halo, sg, x);
}

#ifdef ODIS_TRACE
cudaError_t trace_enable(unsigned long long* buf, unsigned int slots, unsigned int ctas) {
    cudaError_t e;
    const unsigned int zero[2] = {0u, 0u};
    if ((e = cudaMemcpyToSymbol(g_trace_slots, &slots, sizeof slots)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(g_trace_ctas, &ctas, sizeof ctas)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(g_trace_arrivals, zero, sizeof zero)) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_trace, &buf, sizeof buf);
}
#endif

cudaError_t launch_cell_step_pipe(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                                  const HaloInline* halo, const CellSgAccum* sg, const ShExchange* x, cudaStream_t stream) {
    const int n_tiles = (t.n_active + kTile - 1) / kTile;
    if (n_tiles <= 0) return cudaSuccess;
    const int l_sg = sg ? sg->l_max : 0;
    if (l_sg != 0 && !cell_pipe_supports_sg(l_sg)) return cudaErrorInvalidValue;
    const bool part = halo != nullptr;
    if (part && l_sg != 0 && x == nullptr) return cudaErrorInvalidValue;
    const CellRows rows = cell_rows_for(p.potential, l_sg != 0);
    HaloInline none;
    none.n_bnd = 0;
    none.wait_from = 0x7fffffff;
    none.ctl = nullptr;
    CellSgAccum nosg = {};
    ShExchange nox;
    nox.world = 1; nox.rank = 0; nox.ctl = nullptr;
    for (int r = 0; r < kShMaxWorld; r++) nox.block[r] = nullptr;
    const HaloInline& h = halo ? *halo : none;
    const CellSgAccum& a = sg ? *sg : nosg;
    const ShExchange& xx = x ? *x : nox;
#define ODIS_CELL_LAUNCH(L) (part ? launch_cell_cfg<L, true>(t, p, s, mode, next, n_tiles, rows, h, a, xx, stream) \
                                 : launch_cell_cfg<L, false>(t, p, s, mode, next, n_tiles, rows, h, a, xx, stream))
    switch (l_sg) {
        case 2: return ODIS_CELL_LAUNCH(2);
        case 3: return ODIS_CELL_LAUNCH(3);
        case 4: return ODIS_CELL_LAUNCH(4);
        default: return ODIS_CELL_LAUNCH(0);
    }
#undef ODIS_CELL_LAUNCH
}

}  // namespace odis
