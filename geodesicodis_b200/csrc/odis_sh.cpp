// Host side of the spherical-harmonic self-gravity term: basis rows and the least-squares normal inverse.
// See odis_sh.h for what of the reference this follows.
#include "odis_sh.h"

#include <cmath>
#include <cstddef>
#include <omp.h>

namespace odis {

int sh_row_degree(int k) {
    int l = (int)std::floor(std::sqrt((double)k));
    while ((l + 1) * (l + 1) <= k) l++;
    while (l * l > k) l--;
    return l;
}

namespace {

// 4-pi normalised associated Legendre functions with the Condon-Shortley phase, p[l*(l+1)/2 + m] (the index PlmBar
// uses, src/sphericalHarmonics.cpp:146), z = cos(colatitude). Standard column recurrences:
//   Pbar_mm   = sqrt((2m+1)/(2m)) u Pbar_{m-1,m-1}  (Pbar_00 = 1, Pbar_11 = sqrt(3) u),  u = sqrt(1 - z^2)
//   Pbar_lm   = a_lm (z Pbar_{l-1,m} - b_lm Pbar_{l-2,m}),  a_lm = sqrt((4l^2-1)/(l^2-m^2)),  b_lm = sqrt(((l-1)^2-m^2)/(4(l-1)^2-1))
void plm_bar(int l_max, double z, double* p) {
    const double u = std::sqrt((1.0 - z) * (1.0 + z));
    auto at = [](int l, int m) { return l * (l + 1) / 2 + m; };
    p[0] = 1.0;
    for (int m = 1; m <= l_max; m++)
        p[at(m, m)] = (m == 1 ? std::sqrt(3.0) : std::sqrt((2.0 * m + 1.0) / (2.0 * m))) * u * p[at(m - 1, m - 1)];
    for (int m = 0; m <= l_max; m++) {
        if (m + 1 <= l_max) p[at(m + 1, m)] = std::sqrt(2.0 * m + 3.0) * z * p[at(m, m)];
        for (int l = m + 2; l <= l_max; l++) {
            const double a = std::sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
            const double b = std::sqrt(((l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
            p[at(l, m)] = a * (z * p[at(l - 1, m)] - b * p[at(l - 2, m)]);
        }
    }
    for (int l = 1; l <= l_max; l++)
        for (int m = 1; m <= l; m += 2) p[at(l, m)] = -p[at(l, m)];          // csphase = -1: (-1)^m
}

}  // namespace

void sh_basis(int n, const double* pos_sph, int l_max, size_t stride, double* Y) {
    const double half_pi = 0.5 * 3.1415926535897932384626433832795028841971693993751058;
#pragma omp parallel
    {
        std::vector<double> p((size_t)(l_max + 1) * (l_max + 2) / 2);
#pragma omp for schedule(static)
        for (int i = 0; i < n; i++) {
            const double lat = pos_sph[2 * (size_t)i], lon = pos_sph[2 * (size_t)i + 1];
            plm_bar(l_max, std::cos(half_pi - lat), p.data());                // mesh.cpp:2175: cos(pi*0.5 - node_pos_sph(i,0))
            for (int l = 0; l <= l_max; l++) {
                const size_t row = (size_t)l * l;
                Y[row * stride + i] = p[(size_t)l * (l + 1) / 2];
                for (int m = 1; m <= l; m++) {
                    const double plm = p[(size_t)l * (l + 1) / 2 + m];
                    Y[(row + 2 * m - 1) * stride + i] = plm * std::cos((double)m * lon);   // mesh.cpp:2199-2200, :2208-2209
                    Y[(row + 2 * m) * stride + i] = plm * std::sin((double)m * lon);
                }
            }
        }
    }
}

int sh_normal_inverse(int rows, int n, size_t stride, const double* Y, int threads, std::vector<double>& Ginv) {
    const size_t R = (size_t)rows;
    std::vector<double> G(R * R, 0.0);
    if (threads <= 0) threads = omp_get_max_threads();
    // normal matrix, lower triangle; long-double accumulation keeps it independent of the cell order
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int a = 0; a < rows; a++) {
        const double* ya = Y + (size_t)a * stride;
        for (int b = 0; b <= a; b++) {
            const double* yb = Y + (size_t)b * stride;
            long double acc = 0.0L;
            for (int i0 = 0; i0 < n; i0 += 1024) {                            // double inside a chunk, long double across chunks
                const int i1 = i0 + 1024 < n ? i0 + 1024 : n;
                double part = 0.0;
                for (int i = i0; i < i1; i++) part += ya[i] * yb[i];
                acc += (long double)part;
            }
            G[(size_t)a * R + b] = (double)acc;
        }
    }
    // Cholesky G = L L^T in place (lower)
    std::vector<long double> L(R * R, 0.0L);
    for (size_t j = 0; j < R; j++) {
        long double d = G[j * R + j];
        for (size_t k = 0; k < j; k++) d -= L[j * R + k] * L[j * R + k];
        if (!(d > 0.0L)) return -1;
        const long double ljj = sqrtl(d);
        L[j * R + j] = ljj;
        for (size_t i = j + 1; i < R; i++) {
            long double s = G[i * R + j];
            for (size_t k = 0; k < j; k++) s -= L[i * R + k] * L[j * R + k];
            L[i * R + j] = s / ljj;
        }
    }
    // inverse column by column: L L^T x = e_c
    Ginv.assign(R * R, 0.0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int c = 0; c < rows; c++) {
        std::vector<long double> y(R, 0.0L), x(R, 0.0L);
        for (size_t i = (size_t)c; i < R; i++) {
            long double s = (i == (size_t)c) ? 1.0L : 0.0L;
            for (size_t k = (size_t)c; k < i; k++) s -= L[i * R + k] * y[k];
            y[i] = s / L[i * R + i];
        }
        for (size_t ii = R; ii-- > 0;) {
            long double s = y[ii];
            for (size_t k = ii + 1; k < R; k++) s -= L[k * R + ii] * x[k];
            x[ii] = s / L[ii * R + ii];
        }
        for (size_t i = 0; i < R; i++) Ginv[i * R + (size_t)c] = (double)x[i];
    }
    return 0;
}

}  // namespace odis
