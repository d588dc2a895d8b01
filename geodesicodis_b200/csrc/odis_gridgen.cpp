// Icosahedral-bisection grid generator (see odis_gridgen.h).
#include "odis_gridgen.h"

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "odis_sphere.h"

namespace odis {
namespace {

struct P3 { double x, y, z; };
inline P3 normalised(P3 p) {
    const double n = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    return P3{p.x / n, p.y / n, p.z / n};
}

// open-addressing map (a,b)->midpoint id, rebuilt per refinement level
struct EdgeMap {
    std::vector<uint64_t> keys;
    std::vector<int> vals;
    uint64_t mask;
    explicit EdgeMap(size_t expected) {
        size_t cap = 16;
        while (cap < expected * 2 + 16) cap <<= 1;
        keys.assign(cap, ~0ull);
        vals.assign(cap, -1);
        mask = cap - 1;
    }
    int* slot(int a, int b) {
        const uint64_t lo = (uint64_t)(a < b ? a : b), hi = (uint64_t)(a < b ? b : a);
        const uint64_t key = (hi << 32) | lo;
        uint64_t h = key * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        size_t i = (size_t)(h & mask);
        while (keys[i] != ~0ull && keys[i] != key) i = (i + 1) & mask;
        keys[i] = key;
        return &vals[i];
    }
};

inline double canon_degrees(double deg) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.16f", deg);
    return std::strtod(buf, nullptr);
}

}  // namespace

int generate_icosahedral_grid(int level, GridFile& out, std::string& err) {
    if (level < 2 || level > 12) {
        err = "grid level must be in 2..12";
        return -1;
    }
    std::vector<P3> pts;
    std::vector<std::array<int, 3>> tris;
    // icosahedron: 0 north pole, 1 south pole, 2-6 upper ring, 7-11 lower ring
    const double ring_lat = std::atan(0.5);
    pts.push_back(P3{0, 0, 1});
    pts.push_back(P3{0, 0, -1});
    for (int k = 0; k < 5; k++) {
        const double lon = (72.0 * k) * kRadPerDeg;
        pts.push_back(P3{std::cos(ring_lat) * std::cos(lon), std::cos(ring_lat) * std::sin(lon), std::sin(ring_lat)});
    }
    for (int k = 0; k < 5; k++) {
        const double lon = (36.0 + 72.0 * k) * kRadPerDeg;
        pts.push_back(P3{std::cos(ring_lat) * std::cos(lon), std::cos(ring_lat) * std::sin(lon), -std::sin(ring_lat)});
    }
    for (int k = 0; k < 5; k++) {
        const int u0 = 2 + k, u1 = 2 + (k + 1) % 5, l0 = 7 + k, l1 = 7 + (k + 1) % 5;
        tris.push_back({0, u0, u1});
        tris.push_back({u0, l0, u1});
        tris.push_back({l0, l1, u1});
        tris.push_back({1, l1, l0});
    }
    for (int it = 0; it < level - 1; it++) {
        EdgeMap mids(tris.size() * 3 / 2);
        std::vector<std::array<int, 3>> next;
        next.reserve(tris.size() * 4);
        for (const auto& t : tris) {
            int m[3];
            for (int s = 0; s < 3; s++) {
                const int a = t[s], b = t[(s + 1) % 3];
                int* v = mids.slot(a, b);
                if (*v < 0) {
                    *v = (int)pts.size();
                    pts.push_back(normalised(P3{pts[a].x + pts[b].x, pts[a].y + pts[b].y, pts[a].z + pts[b].z}));
                }
                m[s] = *v;
            }
            next.push_back({t[0], m[0], m[2]});
            next.push_back({m[0], t[1], m[1]});
            next.push_back({m[2], m[1], t[2]});
            next.push_back({m[0], m[1], m[2]});
        }
        tris.swap(next);
    }
    const int N = (int)pts.size();
    const size_t T = tris.size();

    // circumcentre of each triangle, canonical degrees
    std::vector<double> cc_lat(T), cc_lon(T);
#pragma omp parallel for schedule(static)
    for (long t = 0; t < (long)T; t++) {
        const P3 p = pts[tris[t][0]], q = pts[tris[t][1]], s = pts[tris[t][2]];
        const P3 u{q.x - p.x, q.y - p.y, q.z - p.z}, w{s.x - p.x, s.y - p.y, s.z - p.z};
        P3 c = normalised(P3{u.y * w.z - u.z * w.y, u.z * w.x - u.x * w.z, u.x * w.y - u.y * w.x});
        if (c.x * p.x + c.y * p.y + c.z * p.z < 0) c = P3{-c.x, -c.y, -c.z};
        double lon = std::atan2(c.y, c.x) / kRadPerDeg;
        if (lon < 0) lon += 360.0;
        cc_lat[t] = canon_degrees(std::atan2(c.z, std::sqrt(c.x * c.x + c.y * c.y)) / kRadPerDeg);
        cc_lon[t] = canon_degrees(lon);
        if (cc_lon[t] >= 360.0) cc_lon[t] = 0.0;
    }

    // clockwise successor around each node: in a ccw triangle (a,b,c) the clockwise step
    // around a goes c -> b
    std::vector<int> cnt((size_t)N, 0);
    std::vector<int> from((size_t)N * 6, -1), to((size_t)N * 6, -1), via((size_t)N * 6, -1);
    for (size_t t = 0; t < T; t++) {
        for (int s = 0; s < 3; s++) {
            const int a = tris[t][s], b = tris[t][(s + 1) % 3], c = tris[t][(s + 2) % 3];
            const int k = cnt[a]++;
            if (k >= 6) { err = "internal: node with more than 6 triangles"; return -2; }
            from[(size_t)a * 6 + k] = c;
            to[(size_t)a * 6 + k] = b;
            via[(size_t)a * 6 + k] = (int)t;
        }
    }
    out.n_cells = N;
    out.node_pos_sph.assign((size_t)N * 2, 0.0);
    out.node_friends.assign((size_t)N * 6, -1);
    out.centroid_pos_sph.assign((size_t)N * 12, -1.0 * kRadPerDeg);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int i = 0; i < N; i++) {
        const int n = cnt[i];
        if (n != 5 && n != 6) { bad++; continue; }
        double lon = std::atan2(pts[i].y, pts[i].x) / kRadPerDeg;
        if (lon < 0) lon += 360.0;
        double lat = std::atan2(pts[i].z, std::sqrt(pts[i].x * pts[i].x + pts[i].y * pts[i].y)) / kRadPerDeg;
        out.node_pos_sph[(size_t)i * 2] = canon_degrees(lat) * kRadPerDeg;
        double lon_c = canon_degrees(lon);
        if (lon_c >= 360.0) lon_c = 0.0;
        out.node_pos_sph[(size_t)i * 2 + 1] = lon_c * kRadPerDeg;
        int cur = from[(size_t)i * 6];
        for (int k = 1; k < n; k++) cur = std::min(cur, from[(size_t)i * 6 + k]);   // start at the lowest id
        for (int j = 0; j < n; j++) {
            int k = 0;
            while (k < n && from[(size_t)i * 6 + k] != cur) k++;
            if (k == n) { bad++; break; }
            out.node_friends[(size_t)i * 6 + j] = cur;
            const int t = via[(size_t)i * 6 + k];
            out.centroid_pos_sph[(size_t)i * 12 + 2 * j] = cc_lat[t] * kRadPerDeg;
            out.centroid_pos_sph[(size_t)i * 12 + 2 * j + 1] = cc_lon[t] * kRadPerDeg;
            cur = to[(size_t)i * 6 + k];
        }
    }
    if (bad) { err = "internal: open neighbour ring"; return -3; }
    return 0;
}

}  // namespace odis
