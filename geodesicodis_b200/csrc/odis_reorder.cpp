#include "odis_reorder.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>

namespace odis {
namespace {
// 2-D Hilbert index of (x,y) on a 2^order grid
inline uint64_t hilbert_index(uint32_t x, uint32_t y, int order) {
    uint64_t d = 0;
    for (uint32_t s = 1u << (order - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            const uint32_t tmp = x; x = y; y = tmp;
        }
    }
    return d;
}
}  // namespace

std::vector<int> cell_locality_order(int n, const double* pos, bool identity) {
    std::vector<int> perm((size_t)n);
    std::iota(perm.begin(), perm.end(), 0);
    if (identity) return perm;
    const int order = 16;
    std::vector<uint64_t> key((size_t)n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const double lat = pos[(size_t)i * 2], lon = pos[(size_t)i * 2 + 1];
        const double x = std::cos(lat) * std::cos(lon), y = std::cos(lat) * std::sin(lon), z = std::sin(lat);
        // octahedral unfolding of the sphere onto [-1,1]^2
        const double s = std::fabs(x) + std::fabs(y) + std::fabs(z);
        double u = x / s, v = y / s;
        if (z < 0) {
            const double uu = (1.0 - std::fabs(v)) * (u >= 0 ? 1.0 : -1.0);
            const double vv = (1.0 - std::fabs(u)) * (v >= 0 ? 1.0 : -1.0);
            u = uu; v = vv;
        }
        const double scale = (double)((1u << order) - 1);
        const uint32_t qx = (uint32_t)std::min(scale, std::max(0.0, (u * 0.5 + 0.5) * scale));
        const uint32_t qy = (uint32_t)std::min(scale, std::max(0.0, (v * 0.5 + 0.5) * scale));
        key[i] = hilbert_index(qx, qy, order);
    }
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    return perm;
}

std::vector<int> edge_locality_order(int F, const int* face_nodes, const std::vector<int>& cell_new_of_old, bool identity) {
    std::vector<int> perm((size_t)F);
    std::iota(perm.begin(), perm.end(), 0);
    if (identity) return perm;
    std::vector<uint64_t> key((size_t)F);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < F; e++) {
        const uint64_t a = (uint64_t)cell_new_of_old[face_nodes[(size_t)e * 2]];
        const uint64_t b = (uint64_t)cell_new_of_old[face_nodes[(size_t)e * 2 + 1]];
        key[e] = (std::min(a, b) << 32) | std::max(a, b);
    }
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
    return perm;
}

std::vector<int> invert_permutation(const std::vector<int>& perm) {
    std::vector<int> inv(perm.size());
    for (size_t i = 0; i < perm.size(); i++) inv[(size_t)perm[i]] = (int)i;
    return inv;
}
}  // namespace odis
