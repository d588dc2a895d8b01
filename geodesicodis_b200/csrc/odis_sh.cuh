// Device side of the spherical-harmonic self-gravity / shell-pressure term (sm_100a, FP64, -fmad=false).
//
// Per time step, after the cell update has produced eta^{n+1} and the tidal potential of the next step:
//   sh_analysis   b = Y eta            dense (rows x cells) matrix-vector product, Y streamed once: the GEMV the reference did
//                                      through SHExpandLSQ (src/extractSHCoeffGG.f95), with the fixed normal-matrix inverse
//                                      hoisted out of the loop; persistent CTAs leave per-CTA partial sums
//   sh_reduce_solve                    adds the partials in CTA order and applies  s = g * factor_l * (Ginv b)_{l >= 2}
//                                      (two launches, reduce then solve, when the basis has more than 128 rows)
//   sh_synthesis  U_i += sum_k Y_ki s_k   the dgemv of pressureGradientSH (src/spatialOperators.cpp:446), fused with the
//                                      add into forcing_potential; rows of degree >= 2 only
// Algorithmic bytes per step: 8*rows*N (analysis) + 8*(rows-4)*N (synthesis) + 8N (eta) + 16N ({eta,U} r/w).
#pragma once
#include <cuda_runtime.h>

namespace odis {

struct ShTables {
    int rows;                 // (l_max+1)^2 basis rows; rows 0..3 (degrees 0, 1) are fitted but never applied
    int stride;               // cells per row of Y / trig (held cells rounded up)
    const double* Y;          // [rows][stride] device cell numbering; nullptr: matrix-free (basis recomputed per cell)
    int l_max;
    const double* trig;       // matrix-free: [4][stride] cos lat, sin lat, cos lon, sin lon (the cell kernel's table)
    const double* rec;        // matrix-free: recurrence coefficients, see sh_recurrence_table()
    const double* Ginv;       // [rows][rows] inverse normal matrix of the least-squares fit
    const double* factor;     // [rows] factor_l of the row's degree (1 - beta_l, or the loading factor); 0 for degrees 0, 1
};

constexpr int kShMaxBlocks = 296;     // analysis CTAs: persistent, two per SM
struct ShWork {
    double* partial;          // [rows][partial_stride] per-CTA sums
    int partial_stride;       // >= kShMaxBlocks
    int n_blocks;             // filled by launch_sh_analysis
    double* b;                // [rows] Y eta
    double* s;                // [rows] g * factor * (Ginv b)
};

// Matrix-free variant (Y == nullptr): both kernels rebuild the basis values of a cell from (sin lat, cos lat, cos lon,
// sin lon) with the Legendre column recurrences and the rotation recurrence for cos/sin(m lon) — 32 B per cell instead of
// 8*rows: FP64-pipe-bound instead of HBM-bound, ~8x less time at l_max = 8. These two kernels use fused multiply-adds
// (the term has no bit-exact reference to follow).
constexpr int kShRecStride = 32;      // l_max <= 31
constexpr int kShRecDoubles = 2 * kShRecStride * kShRecStride + 2 * kShRecStride;
// [a_lm | b_lm | sectoral(m) | first(m)]: Pbar_lm = a_lm (z Pbar_{l-1,m} - b_lm Pbar_{l-2,m}); Pbar_mm = sectoral(m) u Pbar_{m-1,m-1}
// (sign of the Condon-Shortley phase included); Pbar_{m+1,m} = first(m) z Pbar_mm.   Host side; fills kShRecDoubles values.
void sh_recurrence_table(double* rec);
cudaError_t sh_configure();           // dynamic shared-memory opt-in of the matrix-free kernels (before any capture)

// Partitioned runs: every rank fits its own cells, the sums b are all-reduced through peer memory inside the kernels (no host
// call, graph-replayable). Each rank owns one exchange block, mapped by all the others:
//     [kShMaxWorld epoch flags (u64)] [2][kShXRows] doubles: this rank's b of even / odd epochs
//   sh_reduce_publish  this rank's b into its own block (parity of the new epoch); its last CTA then raises this rank's flag in
//                      every rank's block (system-scope release)
//   sh_allsolve        waits until all flags show the epoch (system-scope acquire, bounded), adds the ranks' b in rank order — the
//                      same bits on every rank — and applies the solve
// A rank cannot be more than one epoch ahead of any other (the next allsolve needs everyone's publish), hence two buffers.
constexpr int kShMaxWorld = 8;
constexpr int kShXRows = 1024;
// The folded analysis of the staged cell update (degrees <= 4: at most 25 sums) PUSHES instead: a rank writes its sums into slot
// [parity][rank][kShXSlot] of EVERY rank's block before it raises its flags, so that the readers sum from their own memory.
constexpr int kShXSlot = 32;
static_assert(kShMaxWorld * kShXSlot <= kShXRows, "push layout fits one parity buffer");
constexpr size_t kShXBytes = kShMaxWorld * sizeof(unsigned long long) + 2 * (size_t)kShXRows * sizeof(double);
// ... as "LL" lines (the low-latency protocol of collective libraries): every double travels as ONE 16-byte store {low word, tag, high
// word, tag}, tag = low 32 bits of the epoch. A reader polls the line in its own memory until both tags show the epoch it waits for —
// each 8-byte half {data, tag} arrives atomically, so a matching tag proves its data — and needs neither a flag nor a fence: the
// all-reduce costs ONE NVLink crossing instead of three (data + system fence, release flag, the reader's acquire). Lines live in the
// data region of the block: [2 parities][kShMaxWorld ranks][kShXSlot] x 16 B = 8 KB of its 16 KB. The same solver also runs the
// flag-based pair above once per odis_set_state (the term of the state as set), on the same epoch counter and the same region: its
// doubles (<= 25 rows = the first 200 B of a parity's 8 KB) can only overwrite lines of an epoch two back, which every rank has
// consumed (a rank is never more than one epoch ahead), and a line counts only with BOTH tags equal to the awaited epoch.
struct ShExchange {
    int world, rank;
    unsigned long long* ctl;                 // local: [0] epoch, [1] ticket, [2] set to 1 when a wait gave up, [3] wait limit in clock64 ticks (0: default)
    unsigned char* block[kShMaxWorld];       // every rank's exchange block, own included
};

__device__ __forceinline__ uint4* sh_ll_lines(unsigned char* block, int parity) {
    return reinterpret_cast<uint4*>(block + kShMaxWorld * sizeof(unsigned long long)) + (size_t)parity * (kShMaxWorld * kShXSlot);
}
__device__ __forceinline__ void sh_ll_store(uint4* line, double v, unsigned int tag) {
    const unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
// one look at a line: true (and the value) when both halves carry `tag`
__device__ __forceinline__ bool sh_ll_load(const uint4* line, unsigned int tag, double* v) {
    unsigned int lo, t0, hi, t1;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(line) : "memory");
    *v = __hiloint2double((int)hi, (int)lo);
    return t0 == tag && t1 == tag;
}
// every rank's sums of `epoch` out of this rank's own block into xs[rank * kShXSlot + row] (shared memory), `nthreads` threads with
// index `tid` polling one line each; bounded by `limit` clock64 ticks, a give-up is recorded in ctl[2]
__device__ __forceinline__ void sh_ll_collect(const ShExchange& x, unsigned long long epoch, int rows, double* xs, int tid, int nthreads,
                                              long long limit) {
    const uint4* lines = sh_ll_lines(x.block[x.rank], (int)(epoch & 1ull));
    const unsigned int tag = (unsigned int)epoch;
    for (int q = tid; q < x.world * rows; q += nthreads) {
        const int r = q / rows, k = q - r * rows;
        const long long t0 = clock64();
        double v = 0.0;
        bool ok = sh_ll_load(lines + (size_t)r * kShXSlot + k, tag, &v);
        while (!ok && clock64() - t0 < limit) ok = sh_ll_load(lines + (size_t)r * kShXSlot + k, tag, &v);
        if (!ok) x.ctl[2] = 1ull;
        xs[r * kShXSlot + k] = v;
    }
}

constexpr int kShInlineRows = 128;    // up to here one launch both sums the per-CTA partials and applies the solve
// launches per call of launch_sh_analysis (analysis + reduce/solve)
inline int sh_analysis_launches(int rows, bool partitioned) { return partitioned || rows > kShInlineRows ? 3 : 2; }
// eu: {eta, U} per cell; only the first n_own cells enter the fit. Leaves b and s = g * factor * (Ginv b) in `w`.
// x != nullptr: partitioned run, b is summed over all ranks.
void launch_sh_analysis(const ShTables& t, ShWork w, const double2* eu, int n_own, double g, const ShExchange* x, cudaStream_t stream);
// U of the first n_cells cells (own + halo)
void launch_sh_synthesis(const ShTables& t, const ShWork& w, double2* eu, int n_cells, cudaStream_t stream);

// Default self-gravity step for degrees 2..4 (matrix-free): the harmonic analysis is folded into the staged cell update
// (launch_cell_step_pipe with a CellSgAccum, odis_kernels.cuh), which leaves this rank's sums in w.b — or, partitioned (x != nullptr),
// published in the exchange blocks. This launch completes b (waits for all ranks and adds their sums in rank order), solves and adds
// the term to U of the cells [0, n_cells): edge update, cell update, this = 3 launches per step. Leaves b and s in `w`.
void launch_sh_bsolve_synthesis(const ShTables& t, const ShWork& w, const ShExchange* x, double g, double2* eu, int n_cells, cudaStream_t stream);

}  // namespace odis

// ---- ensembles: M members on one grid, state member-innermost ({eta,U}[cell][Mp]) ---------------------------------------
// The harmonic analysis / synthesis of all members is one FP64 GEMM per direction on the tensor cores
// (mma.sync.aligned.m8n8k4.f64, DMMA):   B[rows x Mp] = Y[rows x N] . Eta[N x Mp]      (analysis, K = cells)
//                                        U[N x Mp]  += Ysel^T[N x rows'] . S[rows' x Mp]  (synthesis, K = harmonic rows >= 4)
// Y is the stored basis (shared by all members). rows <= 128 (degree <= 10).
namespace odis {

struct EnsShTables {
    int rows, rows_pad;       // basis rows, rounded up to a multiple of 8
    int stride;               // cells per row of Y
    int n_cells, Mp;          // Mp: members padded to a multiple of 8
    const double* Y;          // [rows][stride]
    const double* Ginv;       // [rows][rows]
    const double* factor;     // [rows]
    const double* g;          // [Mp] member gravity
};
struct EnsShWork {
    double* partial;          // [blocks][rows_pad][Mp]
    int n_blocks;
    double* b;                // [rows_pad][Mp]
    double* s;                // [rows_pad][Mp]
};
constexpr int kEnsShMaxRows = 128;
constexpr int kEnsShMaxBlocks = 448;
int ens_sh_analysis_blocks(int n_cells, int Mp);
// four launches: analysis GEMM, reduce over its CTAs, per-member solve, synthesis GEMM
constexpr int kEnsShLaunches = 4;
void launch_ens_self_gravity(const EnsShTables& t, EnsShWork w, double2* eu, cudaStream_t stream);

}  // namespace odis
