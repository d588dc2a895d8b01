// Kernels of the spherical-harmonic self-gravity term: see odis_sh.cuh.
#include "odis_sh.cuh"

namespace odis {

namespace {

constexpr int kShThreads = 256;
constexpr int kShWarps = kShThreads / 32;
constexpr int kShRowChunk = 32;
constexpr int kShMaxRows = 1024;         // l_max <= 31

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

// b = Y eta over the cells [0, n_own). One CTA owns a tile of 256*CPT consecutive cells and walks all rows; each thread
// keeps its CPT values of eta in registers, so Y is the only stream.
template <int CPT>
__global__ void __launch_bounds__(kShThreads) sh_analysis_kernel(ShTables t, ShWork w, const double2* __restrict__ eu, int n_own, double g,
                                                                  int inline_solve) {
    __shared__ double red[kShRowChunk][kShWarps];
    __shared__ double bsh[kShInlineRows];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = blockIdx.x * (kShThreads * CPT);
    double e[CPT];
    int at[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        const int i = base + c * kShThreads + tid;
        e[c] = i < n_own ? eu[i].x : 0.0;
        at[c] = i < t.stride ? i : t.stride - 1;          // padded cells: finite basis value times an exact zero
    }
    for (int k0 = 0; k0 < t.rows; k0 += kShRowChunk) {
        const int kn = min(kShRowChunk, t.rows - k0);
#pragma unroll 4
        for (int kk = 0; kk < kn; kk++) {
            const double* __restrict__ row = t.Y + (size_t)(k0 + kk) * t.stride;
            double y[CPT];
#pragma unroll
            for (int c = 0; c < CPT; c++) y[c] = __ldcs(row + at[c]);        // streamed once per step
            double p = y[0] * e[0];
#pragma unroll
            for (int c = 1; c < CPT; c++) p = p + y[c] * e[c];
            p = warp_sum(p);
            if (lane == 0) red[kk][warp] = p;
        }
        __syncthreads();
        if (tid < kn) {
            double a = red[tid][0];
#pragma unroll
            for (int q = 1; q < kShWarps; q++) a = a + red[tid][q];
            w.partial[(size_t)blockIdx.x * t.rows + k0 + tid] = a;
        }
        __syncthreads();
    }
    // the last CTA to finish adds the per-CTA sums in CTA order
    __threadfence();
    if (tid == 0) is_last = atomicAdd(w.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int k = warp; k < t.rows; k += kShWarps) {
        double a = 0.0;
        for (unsigned blk = lane; blk < gridDim.x; blk += 32) a = a + __ldcg(w.partial + (size_t)blk * t.rows + k);
        a = warp_sum(a);
        if (lane == 0) {
            w.b[k] = a;
            if (inline_solve) bsh[k] = a;
        }
    }
    if (tid == 0) *w.ticket = 0u;
    if (inline_solve) {
        __syncthreads();
        solve_rows(t, bsh, g, w.s, warp, kShWarps);
    }
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

int cells_per_thread(int n_cells) {
    // the largest tile that still gives every SM two CTAs; the tile fixes the summation order, so it depends on the grid only
    for (int cpt = 8; cpt > 1; cpt >>= 1)
        if ((n_cells + kShThreads * cpt - 1) / (kShThreads * cpt) >= 296) return cpt;
    return 1;
}

}  // namespace

int sh_analysis_blocks(int n_cells) {
    const int tile = kShThreads * cells_per_thread(n_cells);
    return (n_cells + tile - 1) / tile;
}

void launch_sh_analysis(const ShTables& t, const ShWork& w, const double2* eu, int n_own, double g, cudaStream_t stream) {
    const int cpt = cells_per_thread(n_own), blocks = sh_analysis_blocks(n_own);
    const int inl = t.rows <= kShInlineRows ? 1 : 0;
    switch (cpt) {
        case 8: sh_analysis_kernel<8><<<blocks, kShThreads, 0, stream>>>(t, w, eu, n_own, g, inl); break;
        case 4: sh_analysis_kernel<4><<<blocks, kShThreads, 0, stream>>>(t, w, eu, n_own, g, inl); break;
        case 2: sh_analysis_kernel<2><<<blocks, kShThreads, 0, stream>>>(t, w, eu, n_own, g, inl); break;
        default: sh_analysis_kernel<1><<<blocks, kShThreads, 0, stream>>>(t, w, eu, n_own, g, inl); break;
    }
}

void launch_sh_solve(const ShTables& t, const ShWork& w, double g, cudaStream_t stream) {
    sh_solve_kernel<<<(t.rows + kShWarps - 1) / kShWarps, kShThreads, 0, stream>>>(t, w, g);
}

void launch_sh_synthesis(const ShTables& t, const ShWork& w, double2* eu, int n_cells, cudaStream_t stream) {
    sh_synthesis_kernel<<<(n_cells + kShThreads - 1) / kShThreads, kShThreads, 0, stream>>>(t, w, eu, n_cells);
}

}  // namespace odis
