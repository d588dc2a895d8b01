// Kernels of the spherical-harmonic self-gravity term: see odis_sh.cuh.
#include "odis_sh.cuh"

#include <cmath>

namespace odis {

namespace {

constexpr int kShThreads = 256;
constexpr int kShWarps = kShThreads / 32;
constexpr int kShRowChunk = 32;
constexpr int kShMaxRows = 1024;         // l_max <= 31
constexpr int kShReduceThreads = 1024;
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

constexpr long long kShSpinCycles = 20000000000ll;   // ~10 s

__global__ void __launch_bounds__(kShThreads) sh_allsolve_kernel(ShTables t, ShWork w, ShExchange x, double g) {
    extern __shared__ double dyn[];
    double* bsh = dyn;                                   // [rows]
    const unsigned long long epoch = x.ctl[0];
    if (threadIdx.x < x.world) {
        const unsigned long long* f = x_flags(x.block[x.rank]) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < epoch) {
            if (clock64() - t0 > kShSpinCycles) { x.ctl[2] = 1ull; break; }
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

cudaError_t sh_configure() {
    cudaError_t e;
    double rec[kShRecDoubles];
    sh_recurrence_table(rec);
    if ((e = cudaMemcpyToSymbol(c_rec, rec, sizeof rec)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sh_analysis_mf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analysis_mf_smem(kShRecStride - 1))) != cudaSuccess)
        return e;
    return cudaFuncSetAttribute(sh_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analysis_smem(kShMaxRows));
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

}  // namespace odis
