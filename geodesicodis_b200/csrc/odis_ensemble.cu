// Batched parameter-sweep ensembles: M independent LTE runs on one grid advance together ("ensemble" group of
// include/odis_b200.h; BASELINE config 5: ocean thickness x drag coefficient sweeps).
//
// The reference has no ensemble mode: a sweep is M separate `./ODIS` runs, each streaming the same mesh operators.
// Here the members share the device tables and differ only in their scalars (g, h, alpha, the tidal prefactors) and
// state. State is stored member-innermost — v[F][M], {eta,U}[N][M], histories likewise — and a group of 16 lanes owns
// one (edge, chunk of 32 members) or (cell, chunk of 32 members), two members per lane: a block stages the table rows of
// its next 16 edges in shared memory while it updates the current ones, forming the member-independent stencil
// coefficients (with their exact FP64 divisions) once per edge instead of once per member, and every gather is one
// contiguous run of members.
// Per batched step the tables cross HBM once instead of M times: B_alg = M*(40F + 56N) + 248F + 160N.
//
// Arithmetic per member is operation for operation that of edge_step_kernel / cell_step_kernel (odis_kernels.cu), so
// every member is bit-identical to a single run with the same scalars (tests/test_ensemble_gpu.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/odis_b200.h"
#include "odis_error.h"
#include "odis_kernels.cuh"
#include "odis_reorder.h"
#include "odis_sh.cuh"
#include "odis_sh.h"
#include "odis_sphere.h"

using odis::fail;

#define ODIS_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess)                                                                    \
            return fail(ODIS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));       \
    } while (0)

namespace {

constexpr int kEnsThreads = 256;
constexpr int kGroup = 16;            // lanes per group; 2 members per lane
constexpr int kGroups = kEnsThreads / kGroup;   // groups per block

struct MemberPhys {            // per-member scalars (the rest of odis::Physics is shared)
    double g, h, alpha, factor, factor2, ecc, obl, pad;
};

struct EnsTables {
    int F, N, Fs, Ns;          // sizes and SoA strides
    int Mp;                    // members, padded to a multiple of 8 (pad members repeat the last one)
    int chunks;                // groups per entity = ceil(Mp / 32): a group is kGroup lanes, two members per lane
    int cpb;                   // chunks rounded up to a power of two (<= kGroups): groups of a block per entity
    const int2* cells; const double2* grad; const double* fcor; const double* dist; const double* len;
    const int* sid; const double* sw; const double* sl;     // [10][Fs] stencil ids, weights, lengths l_e' of the stencil edges
    const int* eid; const double* area; const double* cl;   // [6][Ns] edge ids, [Ns], [6][Ns] lengths of the cell's edges
    const double* trig; const double* trig_sq;         // [8][Ns], [2][Ns]
    const MemberPhys* phys;                            // [Mp]
    double dt;
    int potential, friction;
};

__device__ __forceinline__ double exact_div(double x, double d, double y) {      // see odis_kernels.cu
    const double q0 = __dmul_rn(x, y);
    const double r0 = __fma_rn(-q0, d, x);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, d, x);
    return __fma_rn(r1, y, q1);
}
__device__ __forceinline__ double ab3_increment(double f0, double f1, double f2, double dt, int mode) {
    const double a = 23. / 12., b = -16. / 12., c = 5. / 12.;
    if (mode == odis::AB3_FULL) return (a * f0 + b * f1 + c * f2) * dt;
    return f0 * dt;
}
__device__ __forceinline__ double dissipation_flux(int friction, double alpha, double h, double vn, double vt) {
    const double sq = vn * vn + vt * vt;
    if (friction == 0) return alpha * 1000.0 * h * sq;           // energy.cpp:34
    return alpha / h * sqrt(sq) * sq;                             // energy.cpp:48-49
}

// Fixed-order sums of the per-thread energy accumulators (two members per thread). A thread works for the same members
// throughout (the total group count is a multiple of `chunks`): the threads of a block with the same member are added in
// thread order, blocks in block order (energy.cpp:36-40 sums serially; only the association differs).
__device__ __forceinline__ void ensemble_energy_sum(double2 acc, int pair, const EnsTables& t, double* block_partial, unsigned int* ticket,
                                                    double* out) {
    __shared__ double sh[2 * kEnsThreads];
    __shared__ int sh_pair[kEnsThreads];
    __shared__ bool is_last;
    sh[2 * threadIdx.x] = acc.x;
    sh[2 * threadIdx.x + 1] = acc.y;
    sh_pair[threadIdx.x] = pair;
    __syncthreads();
    const int Mp = t.Mp;
    for (int m = threadIdx.x; m < Mp; m += kEnsThreads) {
        double sx = 0.0;
        for (int th = (m >> 1) % kGroup; th < kEnsThreads; th += kGroup)
            if (sh_pair[th] == (m >> 1)) sx += sh[2 * th + (m & 1)];
        block_partial[(size_t)blockIdx.x * Mp + m] = sx;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        // P threads per member add interleaved subsets of the block partials (independent L2 loads), then one thread per
        // member adds the P subset sums in order
        for (int m0 = 0; m0 < Mp; m0 += kEnsThreads) {
            const int cols = min(Mp - m0, kEnsThreads);
            const int P = max(1, kEnsThreads / cols);
            const int col = threadIdx.x % cols, sub = threadIdx.x / cols;
            __syncthreads();
            if (sub < P) {
                double tot = 0.0;
                for (unsigned int b = sub; b < gridDim.x; b += P) tot += __ldcg(block_partial + (size_t)b * Mp + m0 + col);
                sh[sub * cols + col] = tot;
            }
            __syncthreads();
            if ((int)threadIdx.x < cols) {
                double tot = 0.0;
                for (int k = 0; k < P; k++) tot += sh[k * cols + threadIdx.x];
                out[m0 + threadIdx.x] = tot;
            }
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// ---- block-cooperative kernels: a block of 16 groups x 16 lanes works through tiles of edges (cells) ----
// A tile is kGroups / cpb consecutive edges, cpb = chunks rounded up to a power of two: group g of the block owns
// (edge g / cpb of the tile, member chunk g % cpb), lane = member pair. While a tile is being updated, the threads of
// the block fetch the table rows of the block's next tile (coalesced SoA reads, one (slot, edge) item per thread), form the
// member-independent coefficients — one exact FP64 division per item instead of one per member — and park everything
// in the other half of a double-buffered shared-memory stage, from where the groups read it back as broadcasts.
struct __align__(16) EdgeStageRow {
    double2 cw[odis::kStencil];        // {Coriolis coefficient (-2 Omega sin(lat) w l_e' / d_e), w} per stencil slot
    double l[odis::kStencil];          // l_e' of the stencil edges
    int id[odis::kStencil];            // stencil edge ids x Mp/2 (pad slots: the edge itself, weight 0)
    int2 c;                            // inner / outer cell x Mp
    double2 G;
    double d, rd, le, pad;
};

__global__ void __launch_bounds__(kEnsThreads, 2) ens_edge_step_kernel(EnsTables t, const double* __restrict__ v_in, double* __restrict__ v_out,
                                                                       const double2* __restrict__ eu, double* h1, double* h2, int mode,
                                                                       double* block_partial, unsigned int* ticket, double* energy_out) {
    __shared__ EdgeStageRow stage[2][kGroups];
    const int Mp = t.Mp, cpb = t.cpb, Et = kGroups / cpb;
    const int lane = threadIdx.x & (kGroup - 1), g = threadIdx.x / kGroup;
    const int chunk = g % cpb, ge = g / cpb;             // this group's member chunk and its edge within a tile
    const int q = chunk * kGroup + lane;                 // member pair of this lane: members 2q, 2q+1
    const bool active = chunk < t.chunks && 2 * q < Mp;
    const int qa = active ? q : 0;
    const double ga = t.phys[2 * qa].g, gb = t.phys[2 * qa + 1].g, ha = t.phys[2 * qa].h, hb = t.phys[2 * qa + 1].h;
    const double ala = t.phys[2 * qa].alpha, alb = t.phys[2 * qa + 1].alpha;
    const int n_tiles = (t.F + Et - 1) / Et;
    const size_t Fs = (size_t)t.Fs;
    // staging role of this thread: item (slot sj, edge sk of the tile) for threads < 10 * Et; scalars of edge sk for the
    // Et threads from 192 on
    const int sj = threadIdx.x / Et, sk = threadIdx.x % Et;
    const bool stage_item = threadIdx.x < odis::kStencil * Et;
    const bool stage_scalar = threadIdx.x >= 192 && threadIdx.x < 192 + Et;
    // software pipeline: the next tile's table values are requested before this tile's gathers and turned into its
    // stage rows (division included) after this tile's arithmetic
    struct Fetched { int raw; double w, l, d, fc, le; int2 c; double2 G; bool ok; };
    auto fetch = [&](int tile) {
        Fetched f;
        const int e = tile * Et + (stage_scalar ? (int)threadIdx.x - 192 : sk);
        f.ok = tile < n_tiles && e < t.F && (stage_item || stage_scalar);
        f.raw = e; f.w = f.l = f.fc = f.le = 0.0; f.d = 1.0; f.c = make_int2(0, 0); f.G = make_double2(0.0, 0.0);
        if (f.ok) {
            f.d = t.dist[e];
            if (stage_item) {
                const size_t at = (size_t)sj * Fs + e;
                const int raw = t.sid[at];
                f.raw = raw < 0 ? e : raw;                 // pad slots (weight 0) gather the edge itself and add an exact zero
                f.w = t.sw[at]; f.l = t.sl[at]; f.fc = t.fcor[e];
            } else {
                f.c = t.cells[e]; f.G = t.grad[e]; f.le = t.len[e];
            }
        }
        return f;
    };
    const int Mp2 = Mp >> 1;
    auto store = [&](const Fetched& f, int buf) {
        if (!f.ok) return;
        const double rd = __drcp_rn(f.d);
        if (stage_item) {
            EdgeStageRow& r = stage[buf][sk];
            r.id[sj] = f.raw * Mp2;                        // premultiplied: double2 index of (edge, pair 0)
            r.l[sj] = f.l;
            r.cw[sj] = make_double2(exact_div(f.fc * f.w * f.l, f.d, rd), f.w);       // mesh.cpp:2881
        } else {
            EdgeStageRow& r = stage[buf][(int)threadIdx.x - 192];
            r.c = make_int2(f.c.x * Mp, f.c.y * Mp); r.G = f.G; r.d = f.d; r.rd = rd; r.le = f.le;
        }
    };
    const double2* vq = reinterpret_cast<const double2*>(v_in) + qa;
    const double2* euq = eu + 2 * qa;
    double2 esum = make_double2(0.0, 0.0);
    int tile = blockIdx.x, buf = 0;
    store(fetch(tile), 0);
    __syncthreads();
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int e = tile * Et + ge;
        const bool work = active && e < t.F;
        const EdgeStageRow& r = stage[buf][ge];
        const Fetched nxt = fetch(tile + gridDim.x);         // requested first: these come from DRAM
        if (work) {
            const size_t o = (size_t)e * Mp2 + qa;                             // double2 index of (e, pair q)
            double2 nb[odis::kStencil];
#pragma unroll
            for (int j = 0; j < odis::kStencil; j++) nb[j] = vq[r.id[j]];
            const double2 own = reinterpret_cast<const double2*>(v_in)[o];
            const double2 f1 = reinterpret_cast<const double2*>(h1)[o], f2 = reinterpret_cast<const double2*>(h2)[o];
            const int2 c = r.c;
            const double2 in_a = euq[c.x], in_b = euq[c.x + 1], out_a = euq[c.y], out_b = euq[c.y + 1];
            double cor_a = 0.0, cor_b = 0.0, vt_a = 0.0, vt_b = 0.0;
#pragma unroll
            for (int j = 0; j < odis::kStencil; j++) {
                const double2 cw = r.cw[j];
                const double l = r.l[j];
                cor_a += cw.x * nb[j].x;
                cor_b += cw.x * nb[j].y;
                vt_a += nb[j].x * cw.y * l;                                // interpolation.cpp:43
                vt_b += nb[j].y * cw.y * l;
            }
            const double d = r.d, rd = r.rd;
            vt_a = exact_div(vt_a, d, rd);
            vt_b = exact_div(vt_b, d, rd);
            const double area_e = d * r.le;                                // A_e = d_e l_e (mesh.cpp:1093)
            esum.x += dissipation_flux(t.friction, ala, ha, own.x, vt_a) * area_e;
            esum.y += dissipation_flux(t.friction, alb, hb, own.y, vt_b) * area_e;
            const double2 G = r.G;
            // dv/dt = -g G eta + C v (updateMomentum.cpp:42); drag + tidal forcing (timeIntegrator.cpp:219)
            const double f0a = ((-ga * G.x) * in_a.x + (-ga * G.y) * out_a.x) + cor_a;
            const double f0b = ((-gb * G.x) * in_b.x + (-gb * G.y) * out_b.x) + cor_b;
            const double drag_a = (-ala) * own.x + (G.x * in_a.y + G.y * out_a.y);
            const double drag_b = (-alb) * own.y + (G.x * in_b.y + G.y * out_b.y);
            double va = own.x + ab3_increment(f0a, f1.x, f2.x, t.dt, mode);    // temporalOperators.cpp:41,55,64
            double vb = own.y + ab3_increment(f0b, f1.y, f2.y, t.dt, mode);
            va += t.dt * drag_a;                                           // timeIntegrator.cpp:242
            vb += t.dt * drag_b;
            reinterpret_cast<double2*>(v_out)[o] = make_double2(va, vb);
            if (mode == odis::AB3_SECOND) reinterpret_cast<double2*>(h1)[o] = make_double2(f0a, f0b);
            else reinterpret_cast<double2*>(h2)[o] = make_double2(f0a, f0b);
        }
        store(nxt, buf ^ 1);
        __syncthreads();                                     // stage[buf ^ 1] is complete, stage[buf] is free
    }
    ensemble_energy_sum(active ? esum : make_double2(0.0, 0.0), active ? q : -1, t, block_partial, ticket, energy_out);
}

struct TrigValues {
    double cosLat, sinLat, cosLon, sinLon, cos2Lat, sin2Lat, cos2Lon, sin2Lon, cosSq, sinSq;
};
// tidalPotentials.cpp:80-172, same expression shapes as tidal_potential() in odis_kernels.cu
__device__ __forceinline__ double member_potential(int potential, const MemberPhys& p, const odis::StepScalars& m, const TrigValues& v) {
    switch (potential) {
        case odis::P_ECC:
            return p.factor * ((1. - 3. * v.sinSq) * m.cosM + v.cosSq * (3. * m.cosM * v.cos2Lon + 4. * m.sinM * v.sin2Lon));
        case odis::P_OBLIQ:
            return p.factor * m.cosM * v.sin2Lat * v.cosLon;
        case odis::P_OBLIQ_WEST:
            return 3 * p.factor * v.sinLat * v.cosLat * (v.cosLon * m.cosM - v.sinLon * m.sinM);
        case odis::P_FULL:
            return p.factor * ((1 - 3 * v.sinSq) * m.cosM + v.cosSq * (3 * m.cosM * v.cos2Lon + 4 * m.sinM * v.sin2Lon)) +
                   p.factor2 * m.cosM * v.sin2Lat * v.cosLon;
        case odis::P_FULL2: {
            const double ecc = p.ecc, obl = p.obl;
            double T1, T2, T3;
            T1 = 3. * ecc * (4. - 7. * obl * obl) * m.cosM + 6 * (obl * obl + ecc * ecc * (3 - 7 * obl * obl)) * m.cos2M;
            T1 += 3 * ecc * obl * obl * (7 * m.cos3M + 17 * ecc * m.cos4M);
            T1 *= -(1 - 3 * v.cos2Lat);
            T2 = (4 + 15 * ecc * ecc + 20 * ecc * m.cosM + 43 * ecc * ecc * m.cos2M) * v.cosLon;
            T2 += 2 * ecc * (4 + 25 * ecc * m.cosM) * m.sinM * v.sinLon;
            T2 *= 24 * obl * v.cosLat * v.sinLat * m.sinM;
            T3 = obl * obl * (2 + 3 * ecc * ecc + 6 * ecc * m.cosM + 9 * ecc * ecc * m.cos2M) * (m.cosM * v.cosLon + m.sinM * v.sinLon);
            T3 += -(obl * obl - 2) * ((6 * ecc * m.cosM + 17 * ecc * ecc * m.cos2M) * v.cos2Lon + 2 * ecc * (4 + 17 * ecc * m.cosM) * m.sinM * v.sin2Lon);
            T3 *= 6 * v.cosSq;
            return p.factor * (T1 + T2 + T3);
        }
        default:
            return 0.0;
    }
}

struct __align__(16) CellStageRow {
    double coeff[odis::kCellEdges];    // D_ie = -dir l_e / A_i per edge slot (mesh.cpp:3246); 0 for the pentagons' missing slot
    int e[odis::kCellEdges];           // edge ids, -1: no edge
    int pad[2];
    double trig[10];                   // cosLat sinLat cosLon sinLon cos2Lat sin2Lat cos2Lon sin2Lon cos^2Lat sin^2Lat
};

// Same structure for the cells: tiles of kGroups / cpb cells per block. flags: odis::CellFlags.
__global__ void __launch_bounds__(kEnsThreads, 2) ens_cell_step_kernel(EnsTables t, const double* __restrict__ v, const double2* __restrict__ eu_in,
                                                                       double2* __restrict__ eu_out, const double* __restrict__ h1,
                                                                       const double* __restrict__ h2, double* __restrict__ hw, int mode,
                                                                       odis::StepScalars next, int flags) {
    __shared__ CellStageRow stage[2][kGroups];
    const int Mp = t.Mp, cpb = t.cpb, Et = kGroups / cpb;
    const int lane = threadIdx.x & (kGroup - 1), g = threadIdx.x / kGroup;
    const int chunk = g % cpb, gc = g / cpb;
    const int q = chunk * kGroup + lane;
    const bool active = chunk < t.chunks && 2 * q < Mp;
    const int qa = active ? q : 0;
    const MemberPhys pa = t.phys[2 * qa], pb = t.phys[2 * qa + 1];
    const int n_tiles = (t.N + Et - 1) / Et;
    const size_t Ns = (size_t)t.Ns;
    const int sj = threadIdx.x / Et, sk = threadIdx.x % Et;           // staging: (edge slot, cell) for threads < 6 * Et,
    const bool stage_item = threadIdx.x < odis::kCellEdges * Et;      // (trig row, cell) for the 10 * Et threads from 96 on
    const int tj = ((int)threadIdx.x - 96) / Et, tk = ((int)threadIdx.x - 96) % Et;
    const bool stage_trig = threadIdx.x >= 96 && tj < 10;
    auto fetch_and_store = [&](int tile, int buf) {
        if (tile >= n_tiles) return;
        if (stage_item) {
            const int i = tile * Et + sk;
            if (i < t.N) {
                const int packed = t.eid[(size_t)sj * Ns + i];
                CellStageRow& r = stage[buf][sk];
                double coeff = 0.0;
                if (packed != -1) {
                    const double area = t.area[i];
                    const double ndir = (packed < 0) ? 1.0 : -1.0;                          // -dir: dir = -1 for the outer cell
                    coeff = exact_div(ndir * t.cl[(size_t)sj * Ns + i], area, __drcp_rn(area));
                }
                r.e[sj] = packed == -1 ? -1 : (packed & 0x7fffffff);
                r.coeff[sj] = coeff;
            }
        }
        if (stage_trig) {
            const int i = tile * Et + tk;
            if (i < t.N) stage[buf][tk].trig[tj] = tj < 8 ? t.trig[(size_t)tj * Ns + i] : t.trig_sq[(size_t)(tj - 8) * Ns + i];
        }
    };
    int tile = blockIdx.x, buf = 0;
    fetch_and_store(tile, 0);
    __syncthreads();
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int i = tile * Et + gc;
        const bool work = active && i < t.N;
        const CellStageRow& r = stage[buf][gc];
        double2 sa, sb, f1, f2, ed[odis::kCellEdges];
        size_t o = 0;
        if (work) {
            o = ((size_t)i * Mp >> 1) + qa;                   // double2 index of (i, pair q) in the plain arrays
            sa = eu_in[(size_t)i * Mp + 2 * qa]; sb = eu_in[(size_t)i * Mp + 2 * qa + 1];
            if (flags & odis::CELL_UPDATE_ETA) {
                f1 = reinterpret_cast<const double2*>(h1)[o];
                f2 = reinterpret_cast<const double2*>(h2)[o];
#pragma unroll
                for (int j = 0; j < odis::kCellEdges; j++) {
                    const int e = r.e[j];
                    ed[j] = reinterpret_cast<const double2*>(v)[((size_t)(e < 0 ? 0 : e) * Mp >> 1) + qa];
                }
            }
        }
        fetch_and_store(tile + gridDim.x, buf ^ 1);
        if (work) {
            if (flags & odis::CELL_UPDATE_ETA) {
                double div_a = 0.0, div_b = 0.0;                                      // d eta/dt = h Div v (updateEta.cpp:39)
#pragma unroll
                for (int j = 0; j < odis::kCellEdges; j++) {
                    if (r.e[j] >= 0) {                                                // the 12 pentagons have 5 edges
                        const double coeff = r.coeff[j];
                        div_a += (pa.h * coeff) * ed[j].x;
                        div_b += (pb.h * coeff) * ed[j].y;
                    }
                }
                sa.x += ab3_increment(div_a, f1.x, f2.x, t.dt, mode);
                sb.x += ab3_increment(div_b, f1.y, f2.y, t.dt, mode);
                reinterpret_cast<double2*>(hw)[o] = make_double2(div_a, div_b);
            }
            if ((flags & odis::CELL_UPDATE_U) && t.potential != odis::P_NONE) {
                TrigValues tv;
                tv.cosLat = r.trig[0]; tv.sinLat = r.trig[1]; tv.cosLon = r.trig[2]; tv.sinLon = r.trig[3];
                tv.cos2Lat = r.trig[4]; tv.sin2Lat = r.trig[5]; tv.cos2Lon = r.trig[6]; tv.sin2Lon = r.trig[7];
                tv.cosSq = r.trig[8]; tv.sinSq = r.trig[9];
                sa.y = member_potential(t.potential, pa, next, tv);
                sb.y = member_potential(t.potential, pb, next, tv);
            }
            eu_out[(size_t)i * Mp + 2 * qa] = sa;
            eu_out[(size_t)i * Mp + 2 * qa + 1] = sb;
        }
        __syncthreads();
    }
}

// ---- member <-> reference-order host arrays ----
// element (i, m) of an [n][Mp] array of `stride` doubles per element (1: plain arrays, 2: {eta,U}), component `offset`
__global__ void ens_scatter_member(int n, int Mp, int m, const int* __restrict__ perm, const double* __restrict__ src, double* dst, int stride,
                                   int offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[((size_t)i * Mp + m) * stride + offset] = src ? src[perm[i]] : 0.0;
}
__global__ void ens_gather_member(int n, int Mp, int m, const int* __restrict__ perm, const double* __restrict__ src, double* dst, int stride,
                                  int offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[perm[i]] = src[((size_t)i * Mp + m) * stride + offset];
}

}  // namespace

// Layout: plain state arrays are [n][Mp] doubles (member innermost), the {eta,U} arrays [n][Mp] double2.
struct odis_ensemble {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int N = 0, F = 0, Ns = 0, Fs = 0, M = 0, Mp = 0;
    std::vector<odis_params> prm;
    EnsTables tab{};
    int grid_edge = 0;
    // device memory
    std::vector<void*> owned;
    double* d_v[2] = {nullptr, nullptr};
    double* d_hv[2] = {nullptr, nullptr};
    double2* d_eu[2] = {nullptr, nullptr};
    double* d_he[3] = {nullptr, nullptr, nullptr};
    double *d_lvl0_v = nullptr, *d_lvl0_e = nullptr;
    double* d_scratch_f = nullptr;     // [F][Mp] sink of the diagnostics-only edge pass
    unsigned grid_cell = 0;
    int *d_edge_perm = nullptr, *d_cell_perm = nullptr;
    double* d_stage = nullptr;
    double* d_block_partial = nullptr;
    unsigned int* d_ticket = nullptr;
    double* d_series = nullptr;        // [cap][Mp]
    size_t series_cap = 0;
    int cur = 0, ecur = 0, hv1 = 0, he1 = 0, he2 = 1, hefree = 2;
    int64_t iter = 0, iter0 = 0;
    int last_mode = -1;
    bool diag_current = false;
    int64_t launches = 0;
    size_t device_bytes = 0;
    // spherical-harmonic self-gravity term (odis_ensemble_enable_self_gravity): batched FP64 tensor-core GEMMs
    bool sh_on = false;
    odis::EnsShTables sh{};
    odis::EnsShWork shw{};
    std::vector<double> sh_ginv_host;
};

namespace {

template <typename T>
int ens_alloc(odis_ensemble* s, T** p, size_t count) {
    if (count == 0) count = 1;
    ODIS_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    ODIS_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), s->stream));
    s->owned.push_back((void*)*p);
    s->device_bytes += count * sizeof(T);
    return ODIS_OK;
}
template <typename T>
int ens_upload(odis_ensemble* s, const T** p, const std::vector<T>& h) {
    T* d = nullptr;
    int rc = ens_alloc(s, &d, h.size());
    if (rc) return rc;
    if (!h.empty()) ODIS_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    *p = d;
    return ODIS_OK;
}

odis::StepScalars step_scalars(double omega, double time) {
    odis::StepScalars m;
    m.cosM = std::cos(omega * time); m.sinM = std::sin(omega * time);
    m.cos2M = std::cos(2 * omega * time); m.sin2M = std::sin(2 * omega * time);
    m.cos3M = std::cos(3 * omega * time); m.cos4M = std::cos(4 * omega * time);
    return m;
}

int ens_mode(const odis_ensemble* s, int64_t iter) {
    if (iter > 1 || s->prm[0].init_load) return odis::AB3_FULL;
    return iter == 0 ? odis::AB3_FIRST : odis::AB3_SECOND;
}

int ens_ensure_series(odis_ensemble* s, size_t need) {
    if (need <= s->series_cap) return ODIS_OK;
    size_t cap = s->series_cap ? s->series_cap : 1024;
    while (cap < need) cap *= 2;
    const size_t Mp = (size_t)s->Mp;
    double* nd = nullptr;
    ODIS_CUDA(cudaMalloc((void**)&nd, cap * Mp * sizeof(double)));
    ODIS_CUDA(cudaMemsetAsync(nd, 0, cap * Mp * sizeof(double), s->stream));
    if (s->d_series) {
        ODIS_CUDA(cudaMemcpyAsync(nd, s->d_series, s->series_cap * Mp * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
        cudaFree(s->d_series);
        s->device_bytes -= s->series_cap * Mp * sizeof(double);
    }
    s->device_bytes += cap * Mp * sizeof(double);
    s->d_series = nd;
    s->series_cap = cap;
    return ODIS_OK;
}

void ens_rotate_cell_history(odis_ensemble* s, int mode) {
    const int l1 = s->he1, l2 = s->he2, fr = s->hefree;
    if (mode == odis::AB3_FIRST) { s->he2 = fr; s->hefree = l2; }
    else if (mode == odis::AB3_SECOND) { s->he1 = fr; s->hefree = l1; }
    else { s->he1 = fr; s->he2 = l1; s->hefree = l2; }
}

// energy of the newest velocities: a diagnostics-only edge pass (dt = 0, start-up mode: velocities and the history
// value it produces go to the inactive velocity buffer and a scratch array)
int ens_run_diagnostics(odis_ensemble* s) {
    if (s->diag_current) return ODIS_OK;
    int rc = ens_ensure_series(s, (size_t)(s->iter - s->iter0) + 1);
    if (rc) return rc;
    EnsTables t = s->tab;
    t.dt = 0.0;
    ens_edge_step_kernel<<<s->grid_edge, kEnsThreads, 0, s->stream>>>(t, s->d_v[s->cur], s->d_v[1 - s->cur], s->d_eu[s->ecur], s->d_scratch_f,
                                                                     s->d_scratch_f, odis::AB3_FIRST, s->d_block_partial, s->d_ticket,
                                                                     s->d_series + (size_t)(s->iter - s->iter0) * s->Mp);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    s->diag_current = true;
    return ODIS_OK;
}

}  // namespace

extern "C" {

int odis_ensemble_create(const odis_mesh_view* mv, const odis_params* params, int32_t n_members, int32_t device, odis_ensemble** out) {
    if (!mv || !params || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (n_members < 1 || n_members > 32 * kGroups) return fail(ODIS_ERR_ARG, "n_members must be in 1..512");
    if (mv->n_cells < 12 || mv->n_edges != 3 * mv->n_cells - 6) return fail(ODIS_ERR_ARG, "mesh sizes are inconsistent (F != 3N-6)");
    const odis_params& p0 = params[0];
    if (!(p0.dt > 0.0) || !(p0.radius > 0.0)) return fail(ODIS_ERR_ARG, "dt and radius must be positive");
    for (int m = 1; m < n_members; m++) {
        const odis_params& p = params[m];
        if (p.dt != p0.dt || p.omega != p0.omega || p.radius != p0.radius || p.potential != p0.potential || p.friction != p0.friction ||
            p.surface != p0.surface || p.shell_thickness != p0.shell_thickness || p.init_load != p0.init_load)
            return fail(ODIS_ERR_ARG, "ensemble members must share dt, omega, radius, shell thickness, potential, friction and surface type, init_load");
    }
    switch (p0.potential) {
        case odis::P_OBLIQ: case odis::P_OBLIQ_WEST: case odis::P_ECC: case odis::P_FULL: case odis::P_FULL2: case odis::P_NONE: break;
        default: return fail(ODIS_ERR_UNSUPPORTED, "potential type has no expression in the reference (tidalPotentials.cpp:80-285) or is outside the hot path");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(ODIS_ERR_CUDA, "no CUDA device available: the LTE solver has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(ODIS_ERR_ARG, "device ordinal out of range");
    ODIS_CUDA(cudaSetDevice(device));

    odis_ensemble* s = new odis_ensemble();
    auto bail = [&](int code) { odis_ensemble_destroy(s); return code; };
    s->device = device;
    s->N = mv->n_cells; s->F = mv->n_edges;
    s->M = n_members; s->Mp = (n_members + 7) / 8 * 8;
    s->prm.assign(params, params + n_members);
    const int N = s->N, F = s->F, Mp = s->Mp;
    s->Ns = (N + 31) / 32 * 32; s->Fs = (F + 31) / 32 * 32;
    const int Ns = s->Ns, Fs = s->Fs;
    if ((long long)F * Mp >= (1ll << 31)) return bail(fail(ODIS_ERR_ARG, "ensemble too large: edges x members must stay below 2^31"));
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "cudaStreamCreate failed"));
    cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);

    // ---- per-member scalars ----
    std::vector<MemberPhys> phys((size_t)Mp);
    for (int m = 0; m < Mp; m++) {
        const odis_params& p = params[std::min(m, n_members - 1)];      // the pad member repeats the last one
        double radius = p.radius;
        if (p.surface == 2 || p.surface == 3) radius += p.shell_thickness;                  // tidalPotentials.cpp:50-53
        const double om2 = p.omega * p.omega, r2 = radius * radius;
        MemberPhys mp{};
        mp.g = p.g; mp.h = p.h; mp.alpha = p.alpha; mp.ecc = p.ecc; mp.obl = p.obl;
        switch (p.potential) {
            case odis::P_ECC: mp.factor = 0.75 * p.love_reduct * om2 * r2 * p.ecc; break;                       // :84
            case odis::P_OBLIQ: mp.factor = -3. / 2. * p.love_reduct * om2 * r2 * p.obl; break;                 // :106
            case odis::P_OBLIQ_WEST: mp.factor = 0.5 * p.love_reduct * om2 * r2 * p.obl; break;                 // :120
            case odis::P_FULL2: mp.factor = 1 / 32. * p.love_reduct * om2 * r2; break;                          // :135
            case odis::P_FULL:                                                                                   // :160-162
                mp.factor = 0.75 * p.love_reduct * om2 * r2 * p.ecc;
                mp.factor2 = -3. / 2. * p.love_reduct * om2 * r2 * p.obl;
                break;
            default: break;
        }
        phys[(size_t)m] = mp;
    }

    // ---- renumbering and tables (as odis_engine.cu, single rank) ----
    const bool identity = p0.reorder == 0;
    std::vector<int> cperm = odis::cell_locality_order(N, mv->node_pos_sph, identity), cinv = odis::invert_permutation(cperm);
    std::vector<int> eperm = odis::edge_locality_order(F, mv->face_nodes, cinv, identity), einv = odis::invert_permutation(eperm);
    std::vector<int2> cells((size_t)Fs, make_int2(0, 0));
    std::vector<double2> grad((size_t)Fs, make_double2(0.0, 0.0));
    std::vector<double> fcor((size_t)Fs, 0.0), dist((size_t)Fs, 1.0), len((size_t)Fs, 0.0), sw((size_t)Fs * odis::kStencil, 0.0),
        sl((size_t)Fs * odis::kStencil, 0.0);
    std::vector<int> sid((size_t)Fs * odis::kStencil, -1);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int en = 0; en < F; en++) {
        const int eo = eperm[en];
        const int c0 = mv->face_nodes[(size_t)eo * 2], c1 = mv->face_nodes[(size_t)eo * 2 + 1];
        cells[en] = make_int2(cinv[c0], cinv[c1]);
        const double d = mv->face_node_dist[eo];
        grad[en] = make_double2((-mv->face_centre_m[(size_t)eo * 2]) / d, (mv->face_centre_m[(size_t)eo * 2 + 1]) / d);   // mesh.cpp:3076-3080
        fcor[en] = -2.0 * p0.omega * std::sin(mv->face_centre_pos_sph[(size_t)eo * 2]);                                    // mesh.cpp:2881
        dist[en] = d;
        len[en] = mv->face_len[eo];
        int cnt = 10;                                                                                                       // mesh.cpp:2866-2872
        if (mv->node_friends[(size_t)c0 * 6 + 5] < 0) cnt--;
        if (mv->node_friends[(size_t)c1 * 6 + 5] < 0) cnt--;
        int ids[10]; double ws[10];
        for (int j = 0; j < cnt; j++) { ids[j] = mv->face_interp_friends[(size_t)eo * 10 + j]; ws[j] = mv->face_interp_weights[(size_t)eo * 10 + j]; }
        for (int a = 1; a < cnt; a++) {                  // CSR column order: ascending reference edge id
            const int id = ids[a]; const double w = ws[a];
            int b = a - 1;
            while (b >= 0 && ids[b] > id) { ids[b + 1] = ids[b]; ws[b + 1] = ws[b]; b--; }
            ids[b + 1] = id; ws[b + 1] = w;
        }
        for (int j = 0; j < cnt; j++) {
            if (ids[j] < 0 || ids[j] >= F) { bad++; continue; }
            sid[(size_t)j * Fs + en] = einv[ids[j]];
            sw[(size_t)j * Fs + en] = ws[j];
            sl[(size_t)j * Fs + en] = mv->face_len[ids[j]];
        }
        for (int j = cnt; j < odis::kStencil; j++) sl[(size_t)j * Fs + en] = mv->face_len[eo];     // pad slots stand for the edge itself
    }
    std::vector<int> eid((size_t)Ns * odis::kCellEdges, -1);
    std::vector<double> area((size_t)Ns, 1.0), trig((size_t)Ns * 8, 0.0), trig_sq((size_t)Ns * 2, 0.0), cl((size_t)Ns * odis::kCellEdges, 0.0);
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int cn = 0; cn < N; cn++) {
        const int co = cperm[cn];
        const int n = (mv->node_friends[(size_t)co * 6 + 5] < 0) ? 5 : 6;
        int ids[6], dirs[6];
        for (int j = 0; j < n; j++) { ids[j] = mv->faces[(size_t)co * 6 + j]; dirs[j] = mv->node_face_dir[(size_t)co * 6 + j]; }
        for (int a = 1; a < n; a++) {                    // CSR column order of operatorDivergence
            const int id = ids[a], dr = dirs[a];
            int b = a - 1;
            while (b >= 0 && ids[b] > id) { ids[b + 1] = ids[b]; dirs[b + 1] = dirs[b]; b--; }
            ids[b + 1] = id; dirs[b + 1] = dr;
        }
        for (int j = 0; j < n; j++) {
            if (ids[j] < 0 || ids[j] >= F) { bad++; continue; }
            eid[(size_t)j * Ns + cn] = einv[ids[j]] | (dirs[j] < 0 ? (int)0x80000000 : 0);
            cl[(size_t)j * Ns + cn] = mv->face_len[ids[j]];
        }
        area[cn] = mv->control_volume_surf_area_map[co];
        const double lat = mv->node_pos_sph[(size_t)co * 2], lon = mv->node_pos_sph[(size_t)co * 2 + 1];
        trig[0 * (size_t)Ns + cn] = std::cos(lat); trig[1 * (size_t)Ns + cn] = std::sin(lat);       // mesh.cpp:2132-2145
        trig[2 * (size_t)Ns + cn] = std::cos(lon); trig[3 * (size_t)Ns + cn] = std::sin(lon);
        trig[4 * (size_t)Ns + cn] = std::cos(2.0 * lat); trig[5 * (size_t)Ns + cn] = std::sin(2.0 * lat);
        trig[6 * (size_t)Ns + cn] = std::cos(2.0 * lon); trig[7 * (size_t)Ns + cn] = std::sin(2.0 * lon);
        trig_sq[cn] = std::cos(lat) * std::cos(lat);
        trig_sq[(size_t)Ns + cn] = std::sin(lat) * std::sin(lat);
    }
    if (bad) return bail(fail(ODIS_ERR_ARG, "mesh tables hold out-of-range ids"));

    EnsTables& t = s->tab;
    t.F = F; t.N = N; t.Fs = Fs; t.Ns = Ns; t.Mp = Mp; t.chunks = (Mp + 2 * kGroup - 1) / (2 * kGroup);
    t.cpb = 1;
    while (t.cpb < t.chunks) t.cpb *= 2;
    t.dt = p0.dt; t.potential = p0.potential; t.friction = p0.friction;
    int rc;
    const int* d_ep = nullptr; const int* d_cp = nullptr;
    if ((rc = ens_upload(s, &t.cells, cells)) || (rc = ens_upload(s, &t.grad, grad)) || (rc = ens_upload(s, &t.fcor, fcor)) ||
        (rc = ens_upload(s, &t.dist, dist)) || (rc = ens_upload(s, &t.len, len)) || (rc = ens_upload(s, &t.sid, sid)) ||
        (rc = ens_upload(s, &t.sw, sw)) || (rc = ens_upload(s, &t.sl, sl)) || (rc = ens_upload(s, &t.cl, cl)) || (rc = ens_upload(s, &t.eid, eid)) || (rc = ens_upload(s, &t.area, area)) ||
        (rc = ens_upload(s, &t.trig, trig)) || (rc = ens_upload(s, &t.trig_sq, trig_sq)) || (rc = ens_upload(s, &t.phys, phys)) ||
        (rc = ens_upload(s, &d_ep, eperm)) || (rc = ens_upload(s, &d_cp, cperm)))
        return bail(rc);
    s->d_edge_perm = const_cast<int*>(d_ep); s->d_cell_perm = const_cast<int*>(d_cp);
    // persistent blocks, two per SM, each working through tiles of kGroups / cpb edges (cells)
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const int per_tile = kGroups / t.cpb;
        s->grid_edge = (int)std::min<long long>(((long long)F + per_tile - 1) / per_tile, (long long)sms * 2);
        s->grid_cell = (unsigned)std::min<long long>(((long long)N + per_tile - 1) / per_tile, (long long)sms * 2);
    }
    const size_t FE = (size_t)F * Mp, NE = (size_t)N * Mp;
    if ((rc = ens_alloc(s, &s->d_v[0], FE)) || (rc = ens_alloc(s, &s->d_v[1], FE)) || (rc = ens_alloc(s, &s->d_hv[0], FE)) ||
        (rc = ens_alloc(s, &s->d_hv[1], FE)) || (rc = ens_alloc(s, &s->d_lvl0_v, FE)) || (rc = ens_alloc(s, &s->d_scratch_f, FE)) ||
        (rc = ens_alloc(s, &s->d_eu[0], NE)) || (rc = ens_alloc(s, &s->d_eu[1], NE)) || (rc = ens_alloc(s, &s->d_he[0], NE)) ||
        (rc = ens_alloc(s, &s->d_he[1], NE)) || (rc = ens_alloc(s, &s->d_he[2], NE)) || (rc = ens_alloc(s, &s->d_lvl0_e, NE)) ||
        (rc = ens_alloc(s, &s->d_stage, (size_t)F * 3)) || (rc = ens_alloc(s, &s->d_block_partial, (size_t)s->grid_edge * Mp)) ||
        (rc = ens_alloc(s, &s->d_ticket, (size_t)1)))
        return bail(rc);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "table upload failed"));
    *out = s;
    rc = odis_ensemble_set_state(s, -1, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc) { *out = nullptr; return bail(rc); }
    return ODIS_OK;
}

int odis_ensemble_set_state(odis_ensemble* s, int32_t member, const double* v, const double* eta, const double* dvdt, const double* detadt,
                            int64_t iter) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL ensemble");
    if (iter < 0) return fail(ODIS_ERR_ARG, "iter must be >= 0");
    if (member < -1 || member >= s->M) return fail(ODIS_ERR_ARG, "member out of range");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int N = s->N, F = s->F, Mp = s->Mp;
    // the buffer rotation restarts from its initial phase
    if (s->cur != 0 || s->ecur != 0) {
        // keep the other members' newest values where the restarted rotation expects them
        if (s->cur != 0) std::swap(s->d_v[0], s->d_v[1]);
        if (s->ecur != 0) std::swap(s->d_eu[0], s->d_eu[1]);
        s->cur = 0; s->ecur = 0;
    }
    if (s->hv1 != 0) { std::swap(s->d_hv[0], s->d_hv[1]); s->hv1 = 0; }
    {
        double* a[3] = {s->d_he[s->he1], s->d_he[s->he2], s->d_he[s->hefree]};
        s->d_he[0] = a[0]; s->d_he[1] = a[1]; s->d_he[2] = a[2];
        s->he1 = 0; s->he2 = 1; s->hefree = 2;
    }
    auto stage = [&](const double* host, size_t n) -> int {
        if (!host) return ODIS_OK;
        ODIS_CUDA(cudaMemcpyAsync(s->d_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        return ODIS_OK;
    };
    const int m0 = member < 0 ? 0 : member, m1 = member < 0 ? Mp : member + 1;
    const int gF = (F + 255) / 256, gN = (N + 255) / 256;
    int rc;
    if ((rc = stage(v, (size_t)F))) return rc;
    for (int m = m0; m < m1; m++)
        ens_scatter_member<<<gF, 256, 0, s->stream>>>(F, Mp, m, s->d_edge_perm, v ? s->d_stage : nullptr, s->d_v[0], 1, 0);
    if ((rc = stage(eta, (size_t)N))) return rc;
    for (int m = m0; m < m1; m++)
        ens_scatter_member<<<gN, 256, 0, s->stream>>>(N, Mp, m, s->d_cell_perm, eta ? s->d_stage : nullptr, (double*)s->d_eu[0], 2, 0);
    // histories [n][3] in reference order: de-interleave on the host side of the staging buffer, one level at a time
    for (int lvl = 0; lvl < 3; lvl++) {
        std::vector<double> tmp;
        if (dvdt) {
            tmp.resize((size_t)F);
            for (int i = 0; i < F; i++) tmp[(size_t)i] = dvdt[(size_t)i * 3 + lvl];
            ODIS_CUDA(cudaMemcpyAsync(s->d_stage, tmp.data(), (size_t)F * sizeof(double), cudaMemcpyHostToDevice, s->stream));
            ODIS_CUDA(cudaStreamSynchronize(s->stream));
        }
        double* dst = lvl == 0 ? s->d_lvl0_v : s->d_hv[lvl - 1];
        for (int m = m0; m < m1; m++)
            ens_scatter_member<<<gF, 256, 0, s->stream>>>(F, Mp, m, s->d_edge_perm, dvdt ? s->d_stage : nullptr, dst, 1, 0);
        if (detadt) {
            ODIS_CUDA(cudaStreamSynchronize(s->stream));
            tmp.resize((size_t)N);
            for (int i = 0; i < N; i++) tmp[(size_t)i] = detadt[(size_t)i * 3 + lvl];
            ODIS_CUDA(cudaMemcpyAsync(s->d_stage, tmp.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
            ODIS_CUDA(cudaStreamSynchronize(s->stream));
        }
        double* dste = lvl == 0 ? s->d_lvl0_e : s->d_he[lvl - 1];
        for (int m = m0; m < m1; m++)
            ens_scatter_member<<<gN, 256, 0, s->stream>>>(N, Mp, m, s->d_cell_perm, detadt ? s->d_stage : nullptr, dste, 1, 0);
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
    }
    s->launches += (int64_t)(m1 - m0) * 8;
    s->iter = iter; s->iter0 = iter;
    s->last_mode = -1;
    s->diag_current = false;
    // potential for the first step: forcing(current_time + dt), timeIntegrator.cpp:187,218
    const double tt = s->prm[0].dt * (double)iter + s->prm[0].dt;
    ens_cell_step_kernel<<<s->grid_cell, kEnsThreads, 0, s->stream>>>(
        s->tab, s->d_v[0], s->d_eu[0], s->d_eu[0], s->d_he[0], s->d_he[1], s->d_he[2], odis::AB3_FULL, step_scalars(s->prm[0].omega, tt),
        odis::CELL_UPDATE_U);
    s->launches++;
    if (s->sh_on) { odis::launch_ens_self_gravity(s->sh, s->shw, s->d_eu[0], s->stream); s->launches += odis::kEnsShLaunches; }
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

int odis_ensemble_enable_self_gravity(odis_ensemble* s, const odis_mesh_view* mv, int32_t l_max, const double* factor) {
    if (!s || !mv || !factor) return fail(ODIS_ERR_ARG, "NULL argument");
    if (mv->n_cells != s->N) return fail(ODIS_ERR_ARG, "mesh does not match the ensemble");
    if (s->sh_on) return fail(ODIS_ERR_STATE, "self-gravity is already enabled");
    const int rows = odis::sh_rows(l_max);
    if (l_max < 2 || rows > odis::kEnsShMaxRows) return fail(ODIS_ERR_ARG, "ensemble self-gravity: sh degree must be in 2..10");
    if (rows >= s->N) return fail(ODIS_ERR_ARG, "sh degree too high for this grid");
    ODIS_CUDA(cudaSetDevice(s->device));
    ODIS_CUDA(odis::sh_configure());
    const int N = s->N, Ns = s->Ns, Mp = s->Mp;
    {
        std::vector<double> Yg((size_t)rows * N);
        odis::sh_basis(N, mv->node_pos_sph, l_max, (size_t)N, Yg.data());
        if (odis::sh_normal_inverse(rows, N, (size_t)N, Yg.data(), 0, s->sh_ginv_host) != 0)
            return fail(ODIS_ERR_ARG, "spherical-harmonic normal matrix is not positive definite");
    }
    std::vector<int> perm((size_t)N);
    ODIS_CUDA(cudaMemcpy(perm.data(), s->d_cell_perm, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<double> pos((size_t)N * 2), Y((size_t)rows * Ns, 0.0), fac((size_t)rows, 0.0), g((size_t)Mp);
    for (int i = 0; i < N; i++) {
        pos[2 * (size_t)i] = mv->node_pos_sph[2 * (size_t)perm[(size_t)i]];
        pos[2 * (size_t)i + 1] = mv->node_pos_sph[2 * (size_t)perm[(size_t)i] + 1];
    }
    odis::sh_basis(N, pos.data(), l_max, (size_t)Ns, Y.data());
    for (int k = odis::kShSkipRows; k < rows; k++) fac[(size_t)k] = factor[odis::sh_row_degree(k)];
    for (int m = 0; m < Mp; m++) g[(size_t)m] = s->prm[(size_t)std::min(m, s->M - 1)].g;
    odis::EnsShTables& t = s->sh;
    t.rows = rows; t.rows_pad = (rows + 7) / 8 * 8; t.stride = Ns; t.n_cells = N; t.Mp = Mp;
    const int blocks = odis::ens_sh_analysis_blocks(N, Mp);
    int rc;
    if ((rc = ens_upload(s, &t.Y, Y)) || (rc = ens_upload(s, &t.Ginv, s->sh_ginv_host)) || (rc = ens_upload(s, &t.factor, fac)) ||
        (rc = ens_upload(s, &t.g, g)) || (rc = ens_alloc(s, &s->shw.partial, (size_t)blocks * t.rows_pad * Mp)) ||
        (rc = ens_alloc(s, &s->shw.b, (size_t)t.rows_pad * Mp)) || (rc = ens_alloc(s, &s->shw.s, (size_t)t.rows_pad * Mp)))
        return rc;
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    s->sh_on = true;
    // the pending step's potential gets the term of the current eta
    odis::launch_ens_self_gravity(s->sh, s->shw, s->d_eu[s->ecur], s->stream);
    s->launches += odis::kEnsShLaunches;
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

int odis_ensemble_get_sh_coefficients(odis_ensemble* s, int32_t member, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->sh_on) return fail(ODIS_ERR_STATE, "self-gravity is not enabled");
    if (member < 0 || member >= s->M) return fail(ODIS_ERR_ARG, "member out of range");
    ODIS_CUDA(cudaSetDevice(s->device));
    const size_t R = (size_t)s->sh.rows;
    std::vector<double> b((size_t)s->sh.rows_pad * s->Mp);
    ODIS_CUDA(cudaMemcpyAsync(b.data(), s->shw.b, b.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t j = 0; j < R; j++) {
        double acc = 0.0;
        for (size_t k = 0; k < R; k++) acc += s->sh_ginv_host[j * R + k] * b[k * (size_t)s->Mp + (size_t)member];
        out[j] = acc;
    }
    return ODIS_OK;
}

int odis_ensemble_step(odis_ensemble* s, int32_t nsteps) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL ensemble");
    if (nsteps < 0) return fail(ODIS_ERR_ARG, "nsteps must be >= 0");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ens_ensure_series(s, (size_t)(s->iter - s->iter0) + (size_t)nsteps + 1);
    if (rc) return rc;
    const unsigned grid_cell = s->grid_cell;
    for (int k = 0; k < nsteps; k++) {
        const int mode = ens_mode(s, s->iter);
        ens_edge_step_kernel<<<s->grid_edge, kEnsThreads, 0, s->stream>>>(s->tab, s->d_v[s->cur], s->d_v[1 - s->cur], s->d_eu[s->ecur],
                                                                         s->d_hv[s->hv1], s->d_hv[1 - s->hv1], mode, s->d_block_partial, s->d_ticket,
                                                                         s->d_series + (size_t)(s->iter - s->iter0) * s->Mp);
        if (mode == odis::AB3_FULL) s->hv1 = 1 - s->hv1;
        const double tnext = s->prm[0].dt * (double)(s->iter + 1) + s->prm[0].dt;
        ens_cell_step_kernel<<<grid_cell, kEnsThreads, 0, s->stream>>>(s->tab, s->d_v[1 - s->cur], s->d_eu[s->ecur], s->d_eu[1 - s->ecur],
                                                                       s->d_he[s->he1], s->d_he[s->he2], s->d_he[s->hefree], mode,
                                                                       step_scalars(s->prm[0].omega, tnext),
                                                                       odis::CELL_UPDATE_ETA | odis::CELL_UPDATE_U);
        ens_rotate_cell_history(s, mode);
        s->ecur = 1 - s->ecur;
        if (s->sh_on) { odis::launch_ens_self_gravity(s->sh, s->shw, s->d_eu[s->ecur], s->stream); s->launches += odis::kEnsShLaunches; }
        s->cur = 1 - s->cur;
        s->iter++;
        s->last_mode = mode;
        s->launches += 2;
    }
    if (nsteps > 0) s->diag_current = false;
    ODIS_CUDA(cudaGetLastError());
    return ODIS_OK;
}

int odis_ensemble_step_timed(odis_ensemble* s, int32_t nsteps, float* elapsed_ms_out) {
    if (!s || !elapsed_ms_out) return fail(ODIS_ERR_ARG, "NULL argument");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ens_ensure_series(s, (size_t)(s->iter - s->iter0) + (size_t)(nsteps > 0 ? nsteps : 0) + 1);
    if (rc) return rc;
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    ODIS_CUDA(cudaEventRecord(s->ev0, s->stream));
    rc = odis_ensemble_step(s, nsteps);
    if (rc) return rc;
    ODIS_CUDA(cudaEventRecord(s->ev1, s->stream));
    ODIS_CUDA(cudaEventSynchronize(s->ev1));
    ODIS_CUDA(cudaEventElapsedTime(elapsed_ms_out, s->ev0, s->ev1));
    return ODIS_OK;
}

int odis_ensemble_get_field(odis_ensemble* s, int32_t member, int32_t field, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (member < 0 || member >= s->M) return fail(ODIS_ERR_ARG, "member out of range");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int N = s->N, F = s->F, Mp = s->Mp;
    const int gF = (F + 255) / 256, gN = (N + 255) / 256;
    size_t count = 0;
    const int which0 = s->last_mode < 0 ? 0 : (s->last_mode == odis::AB3_FIRST ? 2 : 1);     // as odis_get_field
    switch (field) {
        case ODIS_FIELD_VELOCITY:
            count = (size_t)F;
            ens_gather_member<<<gF, 256, 0, s->stream>>>(F, Mp, member, s->d_edge_perm, s->d_v[s->cur], s->d_stage, 1, 0);
            break;
        case ODIS_FIELD_ETA:
        case ODIS_FIELD_POTENTIAL:
            count = (size_t)N;
            ens_gather_member<<<gN, 256, 0, s->stream>>>(N, Mp, member, s->d_cell_perm, (const double*)s->d_eu[s->ecur], s->d_stage, 2,
                                                         field == ODIS_FIELD_ETA ? 0 : 1);
            break;
        case ODIS_FIELD_DVDT:
        case ODIS_FIELD_DETADT: {
            const bool edges = field == ODIS_FIELD_DVDT;
            const int n = edges ? F : N;
            count = (size_t)n * 3;
            std::vector<double> lvl((size_t)n);
            const double* src[3];
            if (edges) { src[1] = s->d_hv[s->hv1]; src[2] = s->d_hv[1 - s->hv1]; src[0] = which0 == 0 ? s->d_lvl0_v : src[which0]; }
            else { src[1] = s->d_he[s->he1]; src[2] = s->d_he[s->he2]; src[0] = which0 == 0 ? s->d_lvl0_e : src[which0]; }
            for (int k = 0; k < 3; k++) {
                ens_gather_member<<<edges ? gF : gN, 256, 0, s->stream>>>(n, Mp, member, edges ? s->d_edge_perm : s->d_cell_perm, src[k],
                                                                          s->d_stage, 1, 0);
                ODIS_CUDA(cudaMemcpyAsync(lvl.data(), s->d_stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
                ODIS_CUDA(cudaStreamSynchronize(s->stream));
                for (int i = 0; i < n; i++) out[(size_t)i * 3 + k] = lvl[(size_t)i];
            }
            s->launches += 3;
            return ODIS_OK;
        }
        default:
            return fail(ODIS_ERR_ARG, "field not available for ensembles (velocity, eta, potential, dvdt, detadt are)");
    }
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaMemcpyAsync(out, s->d_stage, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

int odis_ensemble_get_dissipation_series(odis_ensemble* s, int32_t member, int64_t first, int64_t count, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (member < 0 || member >= s->M) return fail(ODIS_ERR_ARG, "member out of range");
    const int64_t have = s->iter - s->iter0 + 1;
    if (first < 0 || count < 0 || first + count > have) return fail(ODIS_ERR_ARG, "series range exceeds the steps taken since the state was set");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ens_run_diagnostics(s);
    if (rc) return rc;
    const size_t Mp = (size_t)s->Mp;
    ODIS_CUDA(cudaMemcpy2DAsync(out, sizeof(double), s->d_series + (size_t)first * Mp + member, Mp * sizeof(double), sizeof(double), (size_t)count,
                                cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    const double r = s->prm[(size_t)member].radius;
    const double a = 4 * odis::kPi * (r * r);                   // energy.cpp:60
    for (int64_t k = 0; k < count; k++) out[k] /= a;
    return ODIS_OK;
}

int odis_ensemble_get_info(odis_ensemble* s, int32_t* n_members, int64_t* iter, int64_t* launches, int64_t* device_bytes,
                           int64_t* algorithmic_bytes_per_step) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    if (n_members) *n_members = s->M;
    if (iter) *iter = s->iter;
    if (launches) *launches = s->launches;
    if (device_bytes) *device_bytes = (int64_t)s->device_bytes;
    // per batched step: state M*(40F + 56N) (v r/w 16 + history 24 per edge; {eta,U} r/w 32 + history 24 per cell) and the
    // tables once (edge: ids 40, weights 80, stencil lengths 80, cells 8, grad 16, f 8, d 8, l 8 = 248F; cell: ids 24, lengths 48,
    // area 8, trig 80 = 160N)
    if (algorithmic_bytes_per_step) {
        *algorithmic_bytes_per_step = (int64_t)s->M * (40LL * s->F + 56LL * s->N) + 248LL * s->F + 160LL * s->N;
        // self-gravity: Y streamed twice (shared by the members), {eta,U} read by the analysis, read + written by the synthesis
        if (s->sh_on) *algorithmic_bytes_per_step += 8LL * (2 * s->sh.rows - 4) * s->N + 40LL * s->M * s->N;
    }
    return ODIS_OK;
}

void odis_ensemble_destroy(odis_ensemble* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void* p : s->owned) cudaFree(p);
    if (s->d_series) cudaFree(s->d_series);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // extern "C"
