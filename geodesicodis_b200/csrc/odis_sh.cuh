// Device side of the spherical-harmonic self-gravity / shell-pressure term (sm_100a, FP64, -fmad=false).
//
// Per time step, after the cell update has produced eta^{n+1} and the tidal potential of the next step:
//   sh_analysis   b = Y eta            dense (rows x cells) matrix-vector product, Y streamed once: the GEMV the reference did
//                                      through SHExpandLSQ (src/extractSHCoeffGG.f95), with the fixed normal-matrix inverse
//                                      hoisted out of the loop; per-CTA partial sums, last CTA adds them in index order and
//                                      (small bases) applies  s = g * factor_l * (Ginv b)_{l >= 2}
//   sh_solve      the same  s = ...    as its own launch when the basis is too large for one CTA
//   sh_synthesis  U_i += sum_k Y_ki s_k   the dgemv of pressureGradientSH (src/spatialOperators.cpp:446), fused with the
//                                      add into forcing_potential; rows of degree >= 2 only
// Algorithmic bytes per step: 8*rows*N (analysis) + 8*(rows-4)*N (synthesis) + 8N (eta) + 16N ({eta,U} r/w).
#pragma once
#include <cuda_runtime.h>

namespace odis {

struct ShTables {
    int rows;                 // (l_max+1)^2 basis rows; rows 0..3 (degrees 0, 1) are fitted but never applied
    int stride;               // cells per row of Y (held cells rounded up)
    const double* Y;          // [rows][stride] device cell numbering
    const double* Ginv;       // [rows][rows] inverse normal matrix of the least-squares fit
    const double* factor;     // [rows] factor_l of the row's degree (1 - beta_l, or the loading factor); 0 for degrees 0, 1
};

struct ShWork {
    double* partial;          // [blocks][rows] per-CTA sums
    unsigned int* ticket;
    double* b;                // [rows] Y eta
    double* s;                // [rows] g * factor * (Ginv b)
};

constexpr int kShInlineRows = 128;    // up to here the last analysis CTA does the solve itself
int sh_analysis_blocks(int n_cells);
// eu: {eta, U} per cell; only the first n_own cells enter the fit
void launch_sh_analysis(const ShTables& t, const ShWork& w, const double2* eu, int n_own, double g, cudaStream_t stream);
void launch_sh_solve(const ShTables& t, const ShWork& w, double g, cudaStream_t stream);
// U of the first n_cells cells (own + halo)
void launch_sh_synthesis(const ShTables& t, const ShWork& w, double2* eu, int n_cells, cudaStream_t stream);

}  // namespace odis
