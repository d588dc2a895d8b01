// Spherical-harmonic side of the self-gravity / shell-pressure term (host part).
//
// The reference's version of this term is commented out at HEAD; what it did is still readable:
//   * basis  sh_matrix(i, .) = factor_l * Pbar_lm(cos colat_i) * {cos, sin}(m lon_i), l = 2..l_max
//            (/root/reference/src/mesh.cpp:2154-2260), Pbar_lm from SHTOOLS PlmBar, 4-pi normalised, with the
//            Condon-Shortley phase (src/legendre.f95: csphase = -1);
//   * coefficients of eta by an unweighted least-squares fit over the cell centres for all degrees 0..l_max
//            (src/sphericalHarmonics.cpp:16-72 -> src/extractSHCoeffGG.f95: SHExpandLSQ, csphase = -1);
//   * forcing_potential += g * sh_matrix * coefficients  (src/spatialOperators.cpp:387-462: dgemv with alpha = factor = g,
//            beta = 1), factor_l = 1 - beta_l for the LID_* surfaces (src/boundaryConditions.cpp:373), loading_factor[l] for
//            FREE_LOADING.
// Rows of the basis here: degree-major, per degree  m = 0,  then (cos, sin) for m = 1..l  ->  (l_max+1)^2 rows
// (the reference also carried an all-zero sin(0*lon) row per degree). Rows 0..3 are degrees 0 and 1: fitted, never applied.
#pragma once
#include <vector>

namespace odis {

inline int sh_rows(int l_max) { return (l_max + 1) * (l_max + 1); }
constexpr int kShSkipRows = 4;   // degrees 0 and 1

// degree of basis row k
int sh_row_degree(int k);

// Y[k * stride + i] for cells i < n (lat, lon in radians; [n][2] as node_pos_sph), rows k < sh_rows(l_max).
void sh_basis(int n, const double* pos_sph, int l_max, size_t stride, double* Y);

// Inverse of the normal matrix (Y Y^T)^-1 [rows][rows] of the least-squares fit over the first n cells.
// Returns 0, or -1 when the normal matrix is not positive definite (too few cells for l_max).
int sh_normal_inverse(int rows, int n, size_t stride, const double* Y, int threads, std::vector<double>& Ginv);

}  // namespace odis
