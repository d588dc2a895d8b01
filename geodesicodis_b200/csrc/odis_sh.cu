// Kernels of the spherical-harmonic self-gravity term: see odis_sh.cuh.
#include "odis_sh.cuh"

#include <cmath>
#include <mutex>

namespace odis {

namespace {

constexpr int kShThreads = 256;
constexpr int kShWarps = kShThreads / 32;
constexpr int kShRowChunk = 32;
constexpr int kShMaxRows = 1024;         // l_max <= 31
constexpr int kShReduceThreads = 1024;
constexpr int kShSkip = 4;               // rows of degree 0 and 1
constexpr int kShMaxPartialsPerLane = (kShMaxBlocks + 31) / 32;

__device__ __forceinline__ double warp_sum(double x) {          // butterfly: every lane ends with the same, order-fixed sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// s_j = (g * factor_j) * sum_k Ginv[j][k] b_k for the rows j this warp owns; b is read through `bsrc`
__device__ __forceinline__ void solve_rows(const ShTables& t, const double* bsrc, double g, double* s, int first, int step) {
    const int lane = threadIdx.x & 31;
    for (int j = first; j < t.rows; j += step) {
        const double f = t.factor[j];
        double acc = 0.0;
        if (f != 0.0) {
            const double* row = t.Ginv + (size_t)j * t.rows;
            for (int k = lane; k < t.rows; k += 32) acc = acc + row[k] * bsrc[k];
        }
        acc = warp_sum(acc);
        if (lane == 0) s[j] = (g * f) * acc;
    }
}

// b_k = sum over the analysis CTAs of partial[k][cta], CTA order fixed: lane-strided partial sums (all loads of a lane in
// flight together), then the butterfly
__device__ __forceinline__ double reduce_partials(const double* __restrict__ row, int n_blocks) {
    const int lane = threadIdx.x & 31;
    double v[kShMaxPartialsPerLane];
#pragma unroll
    for (int q = 0; q < kShMaxPartialsPerLane; q++) {
        const int blk = lane + 32 * q;
        v[q] = blk < n_blocks ? __ldcg(row + blk) : 0.0;
    }
    double a = v[0];
#pragma unroll
    for (int q = 1; q < kShMaxPartialsPerLane; q++) a = a + v[q];
    return warp_sum(a);
}

// small bases: every CTA rebuilds b in shared memory (rows * n_blocks doubles from L2) and its warps each solve one row of s
__global__ void __launch_bounds__(kShReduceThreads) sh_reduce_solve_kernel(ShTables t, ShWork w, double g) {
    __shared__ double bsh[kShInlineRows];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < t.rows; k += kShReduceThreads / 32) {
        const double a = reduce_partials(w.partial + (size_t)k * w.partial_stride, w.n_blocks);
        if (lane == 0) {
            bsh[k] = a;
            if (blockIdx.x == 0) w.b[k] = a;
        }
    }
    __syncthreads();
    solve_rows(t, bsh, g, w.s, blockIdx.x * (kShReduceThreads / 32) + warp, gridDim.x * (kShReduceThreads / 32));
}

// large bases: b first (one warp per row), then sh_solve_kernel
__global__ void __launch_bounds__(kShThreads) sh_reduce_kernel(ShTables t, ShWork w) {
    const int k = blockIdx.x * kShWarps + (threadIdx.x >> 5);
    if (k >= t.rows) return;
    const double a = reduce_partials(w.partial + (size_t)k * w.partial_stride, w.n_blocks);
    if ((threadIdx.x & 31) == 0) w.b[k] = a;
}

// b = Y eta over the cells [0, n_own). One CTA owns a tile of 256*CPT consecutive cells and walks all rows; each thread
// keeps its CPT values of eta in registers, so Y is the only stream.
// cells [first, first + 256*CPT) of the analysis: each thread keeps its CPT values of eta in registers, so Y is the only stream
template <int CPT>
__device__ __forceinline__ void analysis_chunk(const ShTables& t, const double2* __restrict__ eu, int n_own, int first, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double e[CPT];
    int at[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int i = first + c * kShThreads + tid;
        e[c] = i < n_own ? eu[i].x : 0.0;
        at[c] = i < t.stride ? i : t.stride - 1;          // padded cells: finite basis value times an exact zero
    }
#pragma unroll 4
    for (int k = 0; k < t.rows; k++) {
        const double* __restrict__ row = t.Y + (size_t)k * t.stride;
        double y[CPT];
#pragma unroll
        for (int c = 0; c < CPT; c++) y[c] = __ldcs(row + at[c]);        // streamed once per step
        double p = y[0] * e[0];
#pragma unroll
        for (int c = 1; c < CPT; c++) p = p + y[c] * e[c];
        p = warp_sum(p);
        if (lane == 0) red[k * kShWarps + warp] += p;
    }
}

// Range of 256-cell groups CTA `b` of `n` owns: the groups are dealt out evenly, a CTA's groups are consecutive.
__device__ __forceinline__ void cta_groups(int n_own, int& g0, int& g1) {
    const int groups = (n_own + kShThreads - 1) / kShThreads;
    const int per = groups / gridDim.x, extra = groups % gridDim.x;
    g0 = blockIdx.x * per + min((int)blockIdx.x, extra);
    g1 = g0 + per + ((int)blockIdx.x < extra ? 1 : 0);
}

__device__ __forceinline__ void write_partials(const ShTables& t, const ShWork& w, const double* red) {
    __syncthreads();
    for (int k = threadIdx.x; k < t.rows; k += kShThreads) {
        double a = red[k * kShWarps];
#pragma unroll
        for (int q = 1; q < kShWarps; q++) a = a + red[k * kShWarps + q];
        w.partial[(size_t)k * w.partial_stride + blockIdx.x] = a;
    }
}

// b = Y eta over the cells [0, n_own): persistent CTAs, each sums its share of the cells into shared memory
__global__ void __launch_bounds__(kShThreads) sh_analysis_kernel(ShTables t, ShWork w, const double2* __restrict__ eu, int n_own) {
    extern __shared__ double dyn[];
    double* red = dyn;                                   // [rows][kShWarps]
    for (int k = threadIdx.x; k < t.rows * kShWarps; k += kShThreads) red[k] = 0.0;
    __syncthreads();
    int g0, g1;
    cta_groups(n_own, g0, g1);
    int g = g0;
    for (; g + 8 <= g1; g += 8) analysis_chunk<8>(t, eu, n_own, g * kShThreads, red);
    if (g + 4 <= g1) { analysis_chunk<4>(t, eu, n_own, g * kShThreads, red); g += 4; }
    if (g + 2 <= g1) { analysis_chunk<2>(t, eu, n_own, g * kShThreads, red); g += 2; }
    if (g < g1) analysis_chunk<1>(t, eu, n_own, g * kShThreads, red);
    write_partials(t, w, red);
}

__global__ void __launch_bounds__(kShThreads) sh_solve_kernel(ShTables t, ShWork w, double g) {
    solve_rows(t, w.b, g, w.s, blockIdx.x * kShWarps + (threadIdx.x >> 5), gridDim.x * kShWarps);
}

// U_i += sum_{k >= 4} Y_ki s_k, rows ascending
__global__ void __launch_bounds__(kShThreads) sh_synthesis_kernel(ShTables t, ShWork w, double2* __restrict__ eu, int n_cells) {
    __shared__ double ssh[kShMaxRows];
    for (int k = threadIdx.x; k < t.rows; k += kShThreads) ssh[k] = w.s[k];
    __syncthreads();
    const int i = blockIdx.x * kShThreads + threadIdx.x;
    if (i >= n_cells) return;
    const double u0 = eu[i].y;
    double acc = 0.0;
    int k = 4;
    for (; k + 8 <= t.rows; k += 8) {
        double y[8];
#pragma unroll
        for (int q = 0; q < 8; q++) y[q] = __ldcs(t.Y + (size_t)(k + q) * t.stride + i);
#pragma unroll
        for (int q = 0; q < 8; q++) acc = acc + y[q] * ssh[k + q];
    }
    for (; k < t.rows; k++) acc = acc + __ldcs(t.Y + (size_t)k * t.stride + i) * ssh[k];
    eu[i].y = u0 + acc;
}

// ---- partitioned runs: all-reduce of b through peer memory -----------------------------------------------------------
__device__ __forceinline__ unsigned long long* x_flags(unsigned char* block) { return reinterpret_cast<unsigned long long*>(block); }
__device__ __forceinline__ double* x_pub(unsigned char* block, int parity) {
    return reinterpret_cast<double*>(block + kShMaxWorld * sizeof(unsigned long long)) + (size_t)parity * kShXRows;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kShThreads) sh_reduce_publish_kernel(ShTables t, ShWork w, ShExchange x) {
    __shared__ int is_last;
    const unsigned long long epoch = x.ctl[0] + 1;               // bumped by the last CTA, after every CTA has read it
    const int k = blockIdx.x * kShWarps + (threadIdx.x >> 5);
    if (k < t.rows) {
        const double a = reduce_partials(w.partial + (size_t)k * w.partial_stride, w.n_blocks);
        if ((threadIdx.x & 31) == 0) x_pub(x.block[x.rank], (int)(epoch & 1))[k] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(reinterpret_cast<unsigned int*>(x.ctl + 1), 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if (threadIdx.x < x.world) st_release_sys(x_flags(x.block[threadIdx.x]) + x.rank, epoch);
    if (threadIdx.x == 0) {
        *reinterpret_cast<unsigned int*>(x.ctl + 1) = 0u;
        x.ctl[0] = epoch;
    }
}

constexpr long long kShSpinCycles = 20000000000ll;   // ~10 s by default; x.ctl[3] > 0 overrides (ODIS_B200_WAIT_TIMEOUT_S, odis_create)

__global__ void __launch_bounds__(kShThreads) sh_allsolve_kernel(ShTables t, ShWork w, ShExchange x, double g) {
    extern __shared__ double dyn[];
    double* bsh = dyn;                                   // [rows]
    const unsigned long long epoch = x.ctl[0];
    if (threadIdx.x < x.world) {
        const unsigned long long* f = x_flags(x.block[x.rank]) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > ((long long)x.ctl[3] > 0 ? (long long)x.ctl[3] : kShSpinCycles)) { x.ctl[2] = 1ull; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < t.rows; k += kShThreads) {
        double a = 0.0;
        for (int r = 0; r < x.world; r++) a = a + ld_relaxed_sys(x_pub(x.block[r], (int)(epoch & 1)) + k);
        bsh[k] = a;
        if (blockIdx.x == 0) w.b[k] = a;
    }
    __syncthreads();
    solve_rows(t, bsh, g, w.s, blockIdx.x * kShWarps + (threadIdx.x >> 5), gridDim.x * kShWarps);
}

// ---- matrix-free variant -------------------------------------------------------------------------------------------
// Recurrence coefficients (independent of l_max), layout of sh_recurrence_table(). Kernels specialised on l_max unroll the
// (m, l) loops completely, so every coefficient is an immediate constant-bank operand and every row index a literal.
__constant__ double c_rec[kShRecDoubles];
struct RecConst {
    __device__ __forceinline__ double a(int m, int l) const { return c_rec[l * kShRecStride + m]; }
    __device__ __forceinline__ double b(int m, int l) const { return c_rec[kShRecStride * kShRecStride + l * kShRecStride + m]; }
    __device__ __forceinline__ double sect(int m) const { return c_rec[2 * kShRecStride * kShRecStride + m]; }
    __device__ __forceinline__ double first(int m) const { return c_rec[2 * kShRecStride * kShRecStride + kShRecStride + m]; }
};
// Shared-memory copy of the recurrence coefficients the kernel needs, packed [m][l] with stride L+1.
struct RecShared {
    const double *pa, *pb, *psect, *pfirst;
    int ld;
    __device__ __forceinline__ double a(int m, int l) const { return pa[m * ld + l]; }
    __device__ __forceinline__ double b(int m, int l) const { return pb[m * ld + l]; }
    __device__ __forceinline__ double sect(int m) const { return psect[m]; }
    __device__ __forceinline__ double first(int m) const { return pfirst[m]; }
};
__device__ __forceinline__ RecShared load_recurrence(const ShTables& t, double* sm) {
    const int L1 = t.l_max + 1;
    double* a = sm; double* b = a + L1 * L1; double* sect = b + L1 * L1; double* first = sect + L1;
    for (int k = threadIdx.x; k < L1 * L1; k += blockDim.x) {
        const int m = k / L1, l = k % L1;
        a[k] = t.rec[l * kShRecStride + m];
        b[k] = t.rec[kShRecStride * kShRecStride + l * kShRecStride + m];
    }
    for (int m = threadIdx.x; m < L1; m += blockDim.x) {
        sect[m] = t.rec[2 * kShRecStride * kShRecStride + m];
        first[m] = t.rec[2 * kShRecStride * kShRecStride + kShRecStride + m];
    }
    return RecShared{a, b, sect, first, L1};
}
__host__ __device__ inline int rec_shared_doubles(int l_max) { return 2 * (l_max + 1) * (l_max + 1) + 2 * (l_max + 1); }

template <int CPT, int LT, typename Rec>
__device__ __forceinline__ void analysis_mf_chunk(const ShTables& t, const Rec& rc, const double2* __restrict__ eu, int n_own, int first,
                                                  double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, L = LT > 0 ? LT : t.l_max;
    double z[CPT], u[CPT], c1[CPT], s1[CPT], e[CPT], cm[CPT], sn[CPT], pmm[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int i = first + c * kShThreads + tid;
        const int at = i < t.stride ? i : t.stride - 1;
        e[c] = i < n_own ? eu[i].x : 0.0;
        u[c] = t.trig[at]; z[c] = t.trig[(size_t)t.stride + at];
        c1[c] = t.trig[2 * (size_t)t.stride + at]; s1[c] = t.trig[3 * (size_t)t.stride + at];
        cm[c] = 1.0; sn[c] = 0.0; pmm[c] = 1.0;
    }
#pragma unroll (LT > 0 ? LT + 1 : 1)
    for (int m = 0; m <= L; m++) {
        double ec[CPT], es[CPT], p1[CPT], p2[CPT];
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            if (m > 0) {
                const double cn = __fma_rn(cm[c], c1[c], -(sn[c] * s1[c]));       // cos(m lon), sin(m lon) by rotation
                sn[c] = __fma_rn(sn[c], c1[c], cm[c] * s1[c]);
                cm[c] = cn;
                pmm[c] = rc.sect(m) * u[c] * pmm[c];
            }
            ec[c] = e[c] * cm[c]; es[c] = e[c] * sn[c];
            p1[c] = pmm[c]; p2[c] = 0.0;
        }
#pragma unroll (LT > 0 ? LT + 1 : 1)
        for (int l = m; l <= L; l++) {
            if (l > m) {
                const double a = l == m + 1 ? rc.first(m) : rc.a(m, l);
                const double b = l == m + 1 ? 0.0 : rc.b(m, l);
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    const double p = a * __fma_rn(z[c], p1[c], -(b * p2[c]));
                    p2[c] = p1[c]; p1[c] = p;
                }
            }
            double pc = p1[0] * ec[0], ps = p1[0] * es[0];
#pragma unroll
            for (int c = 1; c < CPT; c++) { pc = __fma_rn(p1[c], ec[c], pc); ps = __fma_rn(p1[c], es[c], ps); }
            pc = warp_sum(pc);
            if (m > 0) ps = warp_sum(ps);
            if (lane == 0) {
                const int row = l * l + (m ? 2 * m - 1 : 0);
                red[row * kShWarps + warp] += pc;
                if (m > 0) red[(row + 1) * kShWarps + warp] += ps;
            }
        }
    }
}

template <int LT>
__global__ void __launch_bounds__(kShThreads) sh_analysis_mf_kernel(ShTables t, ShWork w, const double2* __restrict__ eu, int n_own) {
    extern __shared__ double dyn[];
    double* red = dyn;                                   // [rows][kShWarps]
    for (int k = threadIdx.x; k < t.rows * kShWarps; k += kShThreads) red[k] = 0.0;
    int g0, g1;
    cta_groups(n_own, g0, g1);
    int g = g0;
    if (LT > 0) {
        const RecConst rc;
        __syncthreads();
        for (; g + 4 <= g1; g += 4) analysis_mf_chunk<4, LT>(t, rc, eu, n_own, g * kShThreads, red);
        if (g + 2 <= g1) { analysis_mf_chunk<2, LT>(t, rc, eu, n_own, g * kShThreads, red); g += 2; }
        if (g < g1) analysis_mf_chunk<1, LT>(t, rc, eu, n_own, g * kShThreads, red);
    } else {
        const RecShared rc = load_recurrence(t, dyn + (size_t)t.rows * kShWarps);
        __syncthreads();
        for (; g + 4 <= g1; g += 4) analysis_mf_chunk<4, 0>(t, rc, eu, n_own, g * kShThreads, red);
        if (g + 2 <= g1) { analysis_mf_chunk<2, 0>(t, rc, eu, n_own, g * kShThreads, red); g += 2; }
        if (g < g1) analysis_mf_chunk<1, 0>(t, rc, eu, n_own, g * kShThreads, red);
    }
    write_partials(t, w, red);
}

template <int LT, typename Rec>
__device__ __forceinline__ double synthesis_mf_cell(const ShTables& t, const Rec& rc, const double* ssh, double u, double z, double c1, double s1) {
    const int L = LT > 0 ? LT : t.l_max;
    double cm = 1.0, sn = 0.0, pmm = 1.0, acc = 0.0;
#pragma unroll (LT > 0 ? LT + 1 : 1)
    for (int m = 0; m <= L; m++) {
        if (m > 0) {
            const double cn = __fma_rn(cm, c1, -(sn * s1));
            sn = __fma_rn(sn, c1, cm * s1);
            cm = cn;
            pmm = rc.sect(m) * u * pmm;
        }
        double p1 = pmm, p2 = 0.0;
#pragma unroll (LT > 0 ? LT + 1 : 1)
        for (int l = m; l <= L; l++) {
            if (l > m) {
                const double a = l == m + 1 ? rc.first(m) : rc.a(m, l);
                const double b = l == m + 1 ? 0.0 : rc.b(m, l);
                const double p = a * __fma_rn(z, p1, -(b * p2));
                p2 = p1; p1 = p;
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

template <int LT>
__global__ void __launch_bounds__(kShThreads) sh_synthesis_mf_kernel(ShTables t, ShWork w, double2* __restrict__ eu, int n_cells) {
    extern __shared__ double dyn[];
    double* ssh = dyn;                                   // [rows]
    for (int k = threadIdx.x; k < t.rows; k += kShThreads) ssh[k] = w.s[k];
    const int i = blockIdx.x * kShThreads + threadIdx.x;
    const int at = i < n_cells ? i : n_cells - 1;
    // the cell's inputs are requested before the barrier, together with the coefficient vector
    const double u = t.trig[at], z = t.trig[(size_t)t.stride + at], c1 = t.trig[2 * (size_t)t.stride + at], s1 = t.trig[3 * (size_t)t.stride + at];
    const double u0 = eu[at].y;
    double acc = 0.0;
    if (LT > 0) {
        __syncthreads();
        acc = synthesis_mf_cell<LT>(t, RecConst(), ssh, u, z, c1, s1);
    } else {
        const RecShared rc = load_recurrence(t, dyn + t.rows);
        __syncthreads();
        acc = synthesis_mf_cell<0>(t, rc, ssh, u, z, c1, s1);
    }
    if (i < n_cells) eu[i].y = u0 + acc;
}

// Second half of the default self-gravity step (degrees 2..4): the staged cell update (cell_step_pipe_kernel, odis_kernels_pipe.cu) has
// left this rank's harmonic sums b = Y eta^{n+1} — in w.b, or, on a partitioned solver, published in the exchange blocks. Persistent
// CTAs: each one completes b (partitioned: waits for all ranks' epoch flags and adds the ranks' sums in rank order — the same bits on
// every rank and in every CTA; the wait / sum protocol is sh_allsolve_kernel's), solves s = g factor (Ginv b) in shared memory and adds
// the term to the potential of its cells (own + ghost), four cells per thread and trip so that their loads are in flight together.
template <int LT>
__global__ void __launch_bounds__(kShThreads) sh_bsolve_synthesis_mf_kernel(ShTables t, ShWork w, ShExchange x, double g, double2* __restrict__ eu,
                                                                            int n_cells) {
    __shared__ double bsh[kShInlineRows];
    __shared__ double ssh[kShInlineRows];
    const int warp = threadIdx.x >> 5;
    if (x.world > 1) {
        // every rank pushed its sums as LL lines into this rank's own block ([parity][rank][kShXSlot], odis_sh.cuh): poll them there
        __shared__ double xs[kShMaxWorld * kShXSlot];
        const unsigned long long epoch = x.ctl[0];
        sh_ll_collect(x, epoch, t.rows, xs, (int)threadIdx.x, kShThreads, (long long)x.ctl[3] > 0 ? (long long)x.ctl[3] : kShSpinCycles);
        __syncthreads();
        for (int k = threadIdx.x; k < t.rows; k += kShThreads) {
            double a = 0.0;
            for (int r = 0; r < x.world; r++) a = a + xs[r * kShXSlot + k];      // rank order: the same bits on every rank and in every CTA
            bsh[k] = a;
            if (blockIdx.x == 0) w.b[k] = a;
        }
    } else {
        for (int k = threadIdx.x; k < t.rows; k += kShThreads) bsh[k] = __ldcg(w.b + k);
    }
    __syncthreads();
    solve_rows(t, bsh, g, ssh, warp, kShWarps);
    __syncthreads();
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < t.rows; k += kShThreads) w.s[k] = ssh[k];
    const RecConst rc;
    constexpr int kBatch = 4;
    const int stride = gridDim.x * kShThreads;
    for (int i0 = blockIdx.x * kShThreads + threadIdx.x; i0 < n_cells; i0 += kBatch * stride) {
        double u[kBatch], z[kBatch], c1[kBatch], s1[kBatch];
        double2 st[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; q++) {
            const int i = i0 + q * stride, at = i < n_cells ? i : i0;
            u[q] = __ldg(t.trig + at); z[q] = __ldg(t.trig + (size_t)t.stride + at);
            c1[q] = __ldg(t.trig + 2 * (size_t)t.stride + at); s1[q] = __ldg(t.trig + 3 * (size_t)t.stride + at);
            st[q] = eu[at];
        }
#pragma unroll
        for (int q = 0; q < kBatch; q++) {
            const int i = i0 + q * stride;
            if (i < n_cells) eu[i] = make_double2(st[q].x, st[q].y + synthesis_mf_cell<LT>(t, rc, ssh, u[q], z[q], c1[q], s1[q]));
        }
    }
}

// ---- ensembles: FP64 tensor-core GEMMs ------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int kEnsKC = 32;            // cells per K chunk of the analysis / cells per tile of the synthesis
constexpr int kEnsLd = 36;            // shared-memory row stride (doubles): conflict-free fragment reads for both operands
constexpr int kEnsMB = 32;            // members per CTA (4 n-tiles)
constexpr int kEnsMaxMT = kEnsShMaxRows / 8;

// B = Y . Eta for one block of 32 members over this CTA's share of the cells. Warp w owns n-tile w % 4 (8 members) and the
// m-tiles (8 harmonic rows each) w / 4, w / 4 + 2, ...: accumulators stay in registers across the whole K range.
__global__ void __launch_bounds__(kShThreads) ens_sh_analysis_kernel(EnsShTables t, EnsShWork w, const double2* __restrict__ eu) {
    __shared__ double Ys[kEnsShMaxRows * kEnsLd];      // [rows_pad][kEnsLd]: Y[row][cell chunk]
    __shared__ double Es[kEnsKC * kEnsLd];             // [cell][member]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * kEnsMB;
    const int MT = t.rows_pad / 8;
    const int nt = warp & 3, mt0 = warp >> 2;
    double acc[kEnsMaxMT / 2][2];
#pragma unroll
    for (int j = 0; j < kEnsMaxMT / 2; j++) acc[j][0] = acc[j][1] = 0.0;
    // this CTA's chunks of 32 cells: consecutive, dealt out evenly
    const int chunks = (t.n_cells + kEnsKC - 1) / kEnsKC;
    const int per = chunks / gridDim.x, extra = chunks % gridDim.x;
    const int c0 = blockIdx.x * per + min((int)blockIdx.x, extra), c1 = c0 + per + ((int)blockIdx.x < extra ? 1 : 0);
    // software pipeline: the next chunk's values travel into registers while the tensor cores work on the current one
    constexpr int kYPer = kEnsShMaxRows * kEnsKC / kShThreads, kEPer = kEnsKC * kEnsMB / kShThreads;
    double yreg[kYPer], ereg[kEPer];
    auto fetch = [&](int ch) {
        const int cell0 = ch * kEnsKC;
#pragma unroll
        for (int u = 0; u < kYPer; u++) {
            const int q = tid + u * kShThreads, r = q / kEnsKC, c = q % kEnsKC;
            yreg[u] = (ch < c1 && r < t.rows && cell0 + c < t.n_cells) ? __ldcs(t.Y + (size_t)r * t.stride + cell0 + c) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kEPer; u++) {
            const int q = tid + u * kShThreads, c = q / kEnsMB, m = q % kEnsMB;
            ereg[u] = (ch < c1 && cell0 + c < t.n_cells && m0 + m < t.Mp) ? eu[(size_t)(cell0 + c) * t.Mp + m0 + m].x : 0.0;
        }
    };
    fetch(c0);
    for (int ch = c0; ch < c1; ch++) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kYPer; u++) {
            const int q = tid + u * kShThreads, r = q / kEnsKC, c = q % kEnsKC;
            if (r < t.rows_pad) Ys[r * kEnsLd + c] = yreg[u];
        }
#pragma unroll
        for (int u = 0; u < kEPer; u++) {
            const int q = tid + u * kShThreads;
            Es[(q / kEnsMB) * kEnsLd + q % kEnsMB] = ereg[u];
        }
        __syncthreads();
        fetch(ch + 1);
#pragma unroll
        for (int k = 0; k < kEnsKC; k += 4) {
            const double bfrag = Es[(k + (lane & 3)) * kEnsLd + nt * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < kEnsMaxMT / 2; j++) {
                const int mt = mt0 + 2 * j;
                if (mt < MT) dmma_m8n8k4(acc[j][0], acc[j][1], Ys[(mt * 8 + (lane >> 2)) * kEnsLd + k + (lane & 3)], bfrag);
            }
        }
    }
    // C fragment: row = lane / 4, columns 2 * (lane % 4) + {0, 1}
    double* out = w.partial + (size_t)blockIdx.x * t.rows_pad * t.Mp;
#pragma unroll
    for (int j = 0; j < kEnsMaxMT / 2; j++) {
        const int mt = mt0 + 2 * j;
        if (mt >= MT) continue;
        const int row = mt * 8 + (lane >> 2), m = m0 + nt * 8 + 2 * (lane & 3);
        if (m < t.Mp) { out[(size_t)row * t.Mp + m] = acc[j][0]; out[(size_t)row * t.Mp + m + 1] = acc[j][1]; }
    }
}

// b[k][m] = sum over the analysis CTAs, CTA order fixed: one warp per (row, member), lanes over the CTAs, then the butterfly
__global__ void __launch_bounds__(kShThreads) ens_sh_reduce_kernel(EnsShTables t, EnsShWork w) {
    const int item = blockIdx.x * kShWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (item >= t.rows_pad * t.Mp) return;
    const size_t step = (size_t)t.rows_pad * t.Mp;
    constexpr int kPer = (kEnsShMaxBlocks + 31) / 32;
    double v[kPer];
#pragma unroll
    for (int q = 0; q < kPer; q++) {
        const int blk = lane + 32 * q;
        v[q] = blk < w.n_blocks ? __ldcg(w.partial + (size_t)blk * step + item) : 0.0;
    }
    double a = v[0];
#pragma unroll
    for (int q = 1; q < kPer; q++) a = a + v[q];
    a = warp_sum(a);
    if (lane == 0) w.b[item] = a;
}

// s[j][m] = g_m factor_j sum_k Ginv[j][k] b[k][m]: a CTA owns 32 rows j x 8 members; its slice of Ginv and of b goes through
// shared memory so that the inner products run without a global load
__global__ void __launch_bounds__(kShThreads) ens_sh_solve_kernel(EnsShTables t, EnsShWork w) {
    __shared__ double Gs[32][kEnsShMaxRows + 1];
    __shared__ double bsh[kEnsShMaxRows][8];
    const int tid = threadIdx.x, j0 = blockIdx.y * 32, mb = blockIdx.x * 8;
    for (int q = tid; q < 32 * t.rows; q += kShThreads) {
        const int jj = q / t.rows, k = q % t.rows;
        Gs[jj][k] = j0 + jj < t.rows ? t.Ginv[(size_t)(j0 + jj) * t.rows + k] : 0.0;
    }
    for (int q = tid; q < t.rows * 8; q += kShThreads) bsh[q >> 3][q & 7] = w.b[(size_t)(q >> 3) * t.Mp + mb + (q & 7)];
    __syncthreads();
    const int jj = tid >> 3, m = tid & 7, j = j0 + jj;
    if (j >= t.rows_pad) return;
    const double f = j < t.rows ? t.factor[j] : 0.0;
    double a = 0.0;
    for (int k = 0; k < t.rows; k++) a = a + Gs[jj][k] * bsh[k][m];
    w.s[(size_t)j * t.Mp + mb + m] = (t.g[mb + m] * f) * a;
}

// U += Ysel^T . S: a CTA walks tiles of 32 cells for one block of 32 members; warp w owns the cell m-tile w % 4 (8 cells) and the
// member n-tiles w / 4, w / 4 + 2
__global__ void __launch_bounds__(kShThreads) ens_sh_synthesis_kernel(EnsShTables t, EnsShWork w, double2* __restrict__ eu) {
    extern __shared__ double dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * kEnsMB;
    const int K = (t.rows - kShSkip + 3) / 4 * 4;      // harmonic rows of degree >= 2, padded to the MMA depth with zeros
    double* Ss = dyn;                                  // [K][kEnsLd]: S[row - 4][member]
    double* Ys = dyn + (size_t)K * kEnsLd;             // [K][kEnsLd]: Y[row - 4][cell]
    for (int q = tid; q < K * kEnsMB; q += kShThreads) {
        const int r = q / kEnsMB, m = q % kEnsMB;
        Ss[r * kEnsLd + m] = (r + kShSkip < t.rows && m0 + m < t.Mp) ? w.s[(size_t)(r + kShSkip) * t.Mp + m0 + m] : 0.0;
    }
    const int mt = warp & 3, nt0 = warp >> 2;
    const int tiles = (t.n_cells + kEnsKC - 1) / kEnsKC;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int cell0 = tile * kEnsKC;
        __syncthreads();
        for (int q = tid; q < K * kEnsKC; q += kShThreads) {
            const int r = q / kEnsKC, c = q % kEnsKC;
            Ys[r * kEnsLd + c] = (r + kShSkip < t.rows && cell0 + c < t.n_cells) ? __ldcs(t.Y + (size_t)(r + kShSkip) * t.stride + cell0 + c) : 0.0;
        }
        __syncthreads();
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int k = 0; k < K; k += 4) {
            const double a = Ys[(k + (lane & 3)) * kEnsLd + mt * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 2; j++) dmma_m8n8k4(acc[j][0], acc[j][1], a, Ss[(k + (lane & 3)) * kEnsLd + (nt0 + 2 * j) * 8 + (lane >> 2)]);
        }
        const int cell = cell0 + mt * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int m = m0 + (nt0 + 2 * j) * 8 + 2 * (lane & 3);
            if (cell < t.n_cells && m < t.Mp) {
                double2* p = eu + (size_t)cell * t.Mp + m;
                p[0].y = p[0].y + acc[j][0];
                p[1].y = p[1].y + acc[j][1];
            }
        }
    }
}

}  // namespace

static int analysis_grid(int n_cells) {
    const int groups = (n_cells + kShThreads - 1) / kShThreads;
    return groups < kShMaxBlocks ? (groups > 0 ? groups : 1) : kShMaxBlocks;
}

void sh_recurrence_table(double* rec) {
    for (int k = 0; k < kShRecDoubles; k++) rec[k] = 0.0;
    double* a = rec; double* b = rec + kShRecStride * kShRecStride; double* sect = b + kShRecStride * kShRecStride; double* first = sect + kShRecStride;
    for (int m = 0; m < kShRecStride; m++) {
        sect[m] = m == 0 ? 1.0 : -(m == 1 ? sqrt(3.0) : sqrt((2.0 * m + 1.0) / (2.0 * m)));      // (-1)^m accumulates
        first[m] = sqrt(2.0 * m + 3.0);
        for (int l = m + 2; l < kShRecStride; l++) {
            a[l * kShRecStride + m] = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
            b[l * kShRecStride + m] = sqrt(((l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
        }
    }
}

static size_t analysis_smem(int rows) { return (size_t)rows * kShWarps * sizeof(double); }
static size_t analysis_mf_smem(int l_max) { return ((size_t)(l_max + 1) * (l_max + 1) * kShWarps + rec_shared_doubles(l_max)) * sizeof(double); }
static size_t synthesis_mf_smem(int l_max) { return ((size_t)(l_max + 1) * (l_max + 1) + rec_shared_doubles(l_max)) * sizeof(double); }

// once per device: another solver on the same GPU may have kernels in flight that read c_rec, and the legacy stream of
// cudaMemcpyToSymbol does not wait for the solvers' non-blocking streams
cudaError_t sh_configure() {
    static std::mutex once;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(once);
    if (done[dev & 63]) return cudaSuccess;
    cudaError_t e;
    double rec[kShRecDoubles];
    sh_recurrence_table(rec);
    if ((e = cudaMemcpyToSymbol(c_rec, rec, sizeof rec)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sh_analysis_mf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analysis_mf_smem(kShRecStride - 1))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(ens_sh_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * (size_t)kEnsShMaxRows * kEnsLd * sizeof(double)))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(sh_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analysis_smem(kShMaxRows))) != cudaSuccess) return e;
    done[dev & 63] = true;
    return cudaSuccess;
}

// degrees with a fully unrolled specialisation of the matrix-free kernels
#define ODIS_SH_DISPATCH(L, CALL)                  \
    switch (L) {                                   \
        case 2: { constexpr int LT = 2; CALL; } break;   \
        case 3: { constexpr int LT = 3; CALL; } break;   \
        case 4: { constexpr int LT = 4; CALL; } break;   \
        case 5: { constexpr int LT = 5; CALL; } break;   \
        case 6: { constexpr int LT = 6; CALL; } break;   \
        case 8: { constexpr int LT = 8; CALL; } break;   \
        case 10: { constexpr int LT = 10; CALL; } break; \
        case 12: { constexpr int LT = 12; CALL; } break; \
        default: { constexpr int LT = 0; CALL; } break;  \
    }

void launch_sh_analysis(const ShTables& t, ShWork w, const double2* eu, int n_own, double g, const ShExchange* x, cudaStream_t stream) {
    w.n_blocks = analysis_grid(n_own);
    if (!t.Y) {
        ODIS_SH_DISPATCH(t.l_max, (sh_analysis_mf_kernel<LT><<<w.n_blocks, kShThreads, analysis_mf_smem(t.l_max), stream>>>(t, w, eu, n_own)))
    } else sh_analysis_kernel<<<w.n_blocks, kShThreads, analysis_smem(t.rows), stream>>>(t, w, eu, n_own);
    // b = sum of the per-CTA partials, s = g * factor * (Ginv b)
    const int grid = (t.rows + kShWarps - 1) / kShWarps;
    if (x) {
        sh_reduce_publish_kernel<<<grid, kShThreads, 0, stream>>>(t, w, *x);
        sh_allsolve_kernel<<<grid, kShThreads, (size_t)t.rows * sizeof(double), stream>>>(t, w, *x, g);
    } else if (t.rows <= kShInlineRows) sh_reduce_solve_kernel<<<(t.rows + kShReduceThreads / 32 - 1) / (kShReduceThreads / 32), kShReduceThreads, 0, stream>>>(t, w, g);
    else {
        sh_reduce_kernel<<<grid, kShThreads, 0, stream>>>(t, w);
        sh_solve_kernel<<<grid, kShThreads, 0, stream>>>(t, w, g);
    }
}

void launch_sh_synthesis(const ShTables& t, const ShWork& w, double2* eu, int n_cells, cudaStream_t stream) {
    if (!t.Y) {
        ODIS_SH_DISPATCH(t.l_max, (sh_synthesis_mf_kernel<LT><<<(n_cells + kShThreads - 1) / kShThreads, kShThreads, synthesis_mf_smem(t.l_max), stream>>>(
                                       t, w, eu, n_cells)))
        return;
    }
    sh_synthesis_kernel<<<(n_cells + kShThreads - 1) / kShThreads, kShThreads, 0, stream>>>(t, w, eu, n_cells);
}

void launch_sh_bsolve_synthesis(const ShTables& t, const ShWork& w, const ShExchange* x, double g, double2* eu, int n_cells, cudaStream_t stream) {
    int grid = (n_cells + kShThreads - 1) / kShThreads;
    if (grid > kShMaxBlocks) grid = kShMaxBlocks;               // persistent: two CTAs per SM
    ShExchange one;
    one.world = 1; one.rank = 0; one.ctl = nullptr;
    for (int r = 0; r < kShMaxWorld; r++) one.block[r] = nullptr;
    const ShExchange& xx = x ? *x : one;
    switch (t.l_max) {
        case 2: sh_bsolve_synthesis_mf_kernel<2><<<grid, kShThreads, 0, stream>>>(t, w, xx, g, eu, n_cells); break;
        case 3: sh_bsolve_synthesis_mf_kernel<3><<<grid, kShThreads, 0, stream>>>(t, w, xx, g, eu, n_cells); break;
        default: sh_bsolve_synthesis_mf_kernel<4><<<grid, kShThreads, 0, stream>>>(t, w, xx, g, eu, n_cells); break;
    }
}

}  // namespace odis

namespace odis {

int ens_sh_analysis_blocks(int n_cells, int Mp) {
    const int chunks = (n_cells + 31) / 32, yb = (Mp + 31) / 32;
    int bx = 296 / yb;                                // two CTAs per SM over both grid dimensions
    if (bx > kEnsShMaxBlocks) bx = kEnsShMaxBlocks;
    if (bx > chunks) bx = chunks;
    return bx < 1 ? 1 : bx;
}

void launch_ens_self_gravity(const EnsShTables& t, EnsShWork w, double2* eu, cudaStream_t stream) {
    const int yb = (t.Mp + 31) / 32;
    w.n_blocks = ens_sh_analysis_blocks(t.n_cells, t.Mp);
    ens_sh_analysis_kernel<<<dim3((unsigned)w.n_blocks, (unsigned)yb), 256, 0, stream>>>(t, w, eu);
    ens_sh_reduce_kernel<<<(t.rows_pad * t.Mp + 7) / 8, 256, 0, stream>>>(t, w);
    ens_sh_solve_kernel<<<dim3((unsigned)(t.Mp / 8), (unsigned)((t.rows_pad + 31) / 32)), 256, 0, stream>>>(t, w);
    const int tiles = (t.n_cells + 31) / 32;
    int bx = 592 / yb;
    if (bx > tiles) bx = tiles;
    const int K = (t.rows - 4 + 3) / 4 * 4;
    ens_sh_synthesis_kernel<<<dim3((unsigned)(bx < 1 ? 1 : bx), (unsigned)yb), 256, 2 * (size_t)K * 36 * sizeof(double), stream>>>(t, w, eu);
}

}  // namespace odis
