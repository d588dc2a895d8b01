// Host assembly of the nonlinear branch's tables: see odis_mesh_nl.h.
#include "odis_mesh_nl.h"

#include <algorithm>
#include <cmath>
#include <utility>

#include "odis_sphere.h"

namespace odis {
namespace {

// mathRoutines.h:431-440
inline double rbf(double dist, double eps) {
    double phi = std::exp(-std::pow(eps * dist, 2.0));
    if (dist < 1e-8) phi = 1;
    return phi;
}

// mathRoutines.h:79-87
inline void vel_transform(double& cos_a, double& sin_a, double lat1, double lat2, double lon1, double lon2) {
    cos_a = std::cos(lat1) * std::cos(lat2);
    cos_a += (1.0 + std::sin(lat1) * std::sin(lat2)) * std::cos(lon2 - lon1);
    cos_a /= (1.0 + std::sin(lat1) * std::sin(lat2) + std::cos(lat1) * std::cos(lat2) * std::cos(lon2 - lon1));
    sin_a = -(std::sin(lat1) + std::sin(lat2)) * std::sin(lon2 - lon1);
    sin_a /= (1.0 + std::sin(lat1) * std::sin(lat2) + std::cos(lat1) * std::cos(lat2) * std::cos(lon2 - lon1));
}

// Inverse of a small dense matrix (row-major n x n) the way Eigen's dynamic-size inverse() goes about it: partial-pivot LU,
// unblocked and right-looking, then one solve per column of the permuted identity.
std::vector<double> dense_inverse(const std::vector<double>& a, int n) {
    std::vector<double> lu(a);
    std::vector<int> piv((size_t)n);
    for (int i = 0; i < n; i++) piv[(size_t)i] = i;
    auto at = [&](int i, int j) -> double& { return lu[(size_t)i * n + j]; };
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(at(k, k));
        for (int i = k + 1; i < n; i++) {
            const double v = std::fabs(at(i, k));
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {
            for (int j = 0; j < n; j++) std::swap(at(k, j), at(p, j));
            std::swap(piv[(size_t)k], piv[(size_t)p]);
        }
        if (at(k, k) != 0.0)
            for (int i = k + 1; i < n; i++) at(i, k) /= at(k, k);
        for (int j = k + 1; j < n; j++)
            for (int i = k + 1; i < n; i++) at(i, j) -= at(i, k) * at(k, j);
    }
    std::vector<double> inv((size_t)n * n);
    std::vector<double> x((size_t)n);
    for (int col = 0; col < n; col++) {
        for (int i = 0; i < n; i++) x[(size_t)i] = (piv[(size_t)i] == col) ? 1.0 : 0.0;
        for (int j = 0; j < n; j++)
            for (int i = j + 1; i < n; i++) x[(size_t)i] -= at(i, j) * x[(size_t)j];
        for (int j = n - 1; j >= 0; j--) {
            x[(size_t)j] /= at(j, j);
            for (int i = 0; i < j; i++) x[(size_t)i] -= at(i, j) * x[(size_t)j];
        }
        for (int i = 0; i < n; i++) inv[(size_t)i * n + col] = x[(size_t)i];
    }
    return inv;
}

// One output row of a chain of sparse products, as a small (column, value) list. add() follows the accumulation rule of a row-wise
// sparse product: the first contribution to a column is stored, later ones are added, in the order they arrive.
struct RowAcc {
    std::vector<std::pair<int, double>> e;
    void clear() { e.clear(); }
    void add(int col, double v) {
        for (auto& p : e)
            if (p.first == col) { p.second += v; return; }
        e.emplace_back(col, v);
    }
    void sort() { std::sort(e.begin(), e.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; }); }
};

// Rows are independent: chunks of them are filled in parallel into fixed-width buffers (fill(r, row) leaves row r in `row`), then
// appended in order. max_width bounds the entries of one row.
template <typename Fill>
int build_csr(Csr& A, int n_rows, int n_cols, int max_width, bool prune_zeros, Fill fill) {
    A.n_rows = n_rows; A.n_cols = n_cols;
    A.indptr.assign(1, 0); A.indices.clear(); A.data.clear();
    const int chunk = 1 << 18;
    std::vector<int> cols((size_t)chunk * max_width), cnt((size_t)chunk);
    std::vector<double> vals((size_t)chunk * max_width);
    int bad = 0;
    for (int r0 = 0; r0 < n_rows; r0 += chunk) {
        const int r1 = std::min(n_rows, r0 + chunk);
#pragma omp parallel reduction(+ : bad)
        {
            RowAcc row;
#pragma omp for schedule(static)
            for (int r = r0; r < r1; r++) {
                row.clear();
                fill(r, row);
                row.sort();
                int k = 0;
                for (auto& p : row.e) {
                    if (prune_zeros && p.second == 0.0) continue;
                    if (k >= max_width) { bad++; break; }
                    cols[(size_t)(r - r0) * max_width + k] = p.first;
                    vals[(size_t)(r - r0) * max_width + k] = p.second;
                    k++;
                }
                cnt[(size_t)(r - r0)] = k;
            }
        }
        for (int r = r0; r < r1; r++) {
            const int k = cnt[(size_t)(r - r0)];
            A.indices.insert(A.indices.end(), cols.begin() + (size_t)(r - r0) * max_width, cols.begin() + (size_t)(r - r0) * max_width + k);
            A.data.insert(A.data.end(), vals.begin() + (size_t)(r - r0) * max_width, vals.begin() + (size_t)(r - r0) * max_width + k);
            A.indptr.push_back((int)A.indices.size());
        }
    }
    return bad;
}

}  // namespace

int build_nonlinear_tables(const MeshTables& m, double radius, double rbf_eps, NonlinearTables& out, std::string& err) {
    const int N = m.n_cells, F = m.n_edges, V = m.n_vertices;
    const double r = radius;
    if (m.vertex_pos_sph.size() != (size_t)V * 2 || m.face_vertexes.size() != (size_t)F * 2) { err = "mesh has no vertex tables"; return -1; }
    auto node = [&](int i) { return LatLon{m.node_pos_sph[(size_t)i * 2], m.node_pos_sph[(size_t)i * 2 + 1]}; };
    auto vertex = [&](int v) { return LatLon{m.vertex_pos_sph[(size_t)v * 2], m.vertex_pos_sph[(size_t)v * 2 + 1]}; };
    auto face_centre = [&](int e) { return LatLon{m.face_centre_pos_sph[(size_t)e * 2], m.face_centre_pos_sph[(size_t)e * 2 + 1]}; };
    auto sides = [&](int i) { return m.node_friends[(size_t)i * 6 + 5] < 0 ? 5 : 6; };

    // ---- vertices: mesh.cpp:500 (sin lat), :910-1030 (faces around a vertex and their sense), :1107-1147 (area)
    out.vertex_sinlat.resize((size_t)V);
    out.vertex_area.resize((size_t)V);
    out.vertex_faces.assign((size_t)V * 3, -1);
    out.vertex_face_dir.assign((size_t)V * 3, 0);
    for (int v = 0; v < V; v++) out.vertex_sinlat[(size_t)v] = std::sin(m.vertex_pos_sph[(size_t)v * 2]);
    for (int e = 0; e < F; e++) {
        for (int side = 0; side < 2; side++) {
            const int v = m.face_vertexes[(size_t)e * 2 + side];
            for (int j = 0; j < 3; j++) {
                if (out.vertex_faces[(size_t)v * 3 + j] < 0) {
                    out.vertex_faces[(size_t)v * 3 + j] = e;
                    const Vec2 p = map_project(face_centre(e), vertex(v), r);
                    const double fnx = m.face_normal_vec_map[(size_t)e * 2], fny = m.face_normal_vec_map[(size_t)e * 2 + 1];
                    const double cross = p.x * fny - p.y * fnx;
                    out.vertex_face_dir[(size_t)v * 3 + j] = cross > 0 ? 1 : -1;
                    break;
                }
            }
        }
    }
    for (int v = 0; v < V; v++) {
        double area = 0.0;
        for (int j = 0; j < 3; j++) {
            const int n1 = m.vertex_nodes[(size_t)v * 3 + j], n2 = m.vertex_nodes[(size_t)v * 3 + (j + 1) % 3];
            area += std::fabs(spherical_triangle_area(vertex(v), node(n1), node(n2), r));
        }
        out.vertex_area[(size_t)v] = area;
    }

    // ---- operatorCurl, mesh.cpp:3122-3175
    for (int v = 0; v < V; v++)
        for (int j = 0; j < 3; j++)
            if (out.vertex_faces[(size_t)v * 3 + j] < 0) { err = "vertex with fewer than three edges"; return -2; }
    if (build_csr(out.curl, V, F, 3, false, [&](int v, RowAcc& row) {
            const double area = out.vertex_area[(size_t)v];
            for (int j = 0; j < 3; j++) {
                const int e = out.vertex_faces[(size_t)v * 3 + j];
                const double t_ev = out.vertex_face_dir[(size_t)v * 3 + j];
                row.add(e, m.face_node_dist[(size_t)e] * t_ev / area);
            }
        })) { err = "operatorCurl row too wide"; return -3; }

    // ---- per-cell map coordinates of the neighbours (mesh.cpp:1384-1425), per-edge Cartesian normal (:641) and the
    //      velocity-transform angles between the edge and its two cells (:1335-1382)
    std::vector<double> pos_map((size_t)N * 7 * 2, -1.0);
    for (int i = 0; i < N; i++) {
        const Vec2 self = map_project(node(i), node(i), r);
        pos_map[(size_t)i * 14] = self.x; pos_map[(size_t)i * 14 + 1] = self.y;
        for (int j = 1; j < 7; j++) {
            const int f = m.node_friends[(size_t)i * 6 + j - 1];
            if (f < 0) continue;
            const Vec2 p = map_project(node(i), node(f), r);
            pos_map[(size_t)i * 14 + (size_t)j * 2] = p.x; pos_map[(size_t)i * 14 + (size_t)j * 2 + 1] = p.y;
        }
    }
    std::vector<double> normal_xyz((size_t)F * 3), vel_trans((size_t)F * 4);
    for (int i = 0; i < N; i++) {
        const int n = sides(i);
        for (int j = 0; j < n; j++) {
            const int e = m.faces[(size_t)i * 6 + j];
            if (m.face_nodes[(size_t)e * 2] != i) continue;                      // the inner cell names the edge's corners (mesh.cpp:552-555)
            const size_t ca = (size_t)i * 12 + (size_t)((j + n - 1) % n) * 2, cb = (size_t)i * 12 + (size_t)j * 2;
            const LatLon a{m.centroid_pos_sph[ca], m.centroid_pos_sph[ca + 1]}, b{m.centroid_pos_sph[cb], m.centroid_pos_sph[cb + 1]};
            const Vec3 nx = great_circle_normal(a, b);
            normal_xyz[(size_t)e * 3] = nx.x; normal_xyz[(size_t)e * 3 + 1] = nx.y; normal_xyz[(size_t)e * 3 + 2] = nx.z;
        }
    }
    for (int e = 0; e < F; e++) {
        const double lat1 = m.face_intercept_pos_sph[(size_t)e * 2], lon1 = m.face_intercept_pos_sph[(size_t)e * 2 + 1];
        for (int k = 0; k < 2; k++) {
            const LatLon p = node(m.face_nodes[(size_t)e * 2 + k]);
            vel_trans[(size_t)e * 4 + (size_t)k * 2] = 0.0; vel_trans[(size_t)e * 4 + (size_t)k * 2 + 1] = 0.0;
            vel_transform(vel_trans[(size_t)e * 4 + (size_t)k * 2], vel_trans[(size_t)e * 4 + (size_t)k * 2 + 1], lat1, p.lat, lon1, p.lon);
        }
    }

    // ---- operatorRBFinterp = RBF * (Vinv * node2faceAdj), mesh.cpp:359-432 (per-cell matrix and its inverse), :2263-2361
    {
        // per cell: the inverse of its RBF-normal matrix (6 x 6 slots) and the node-to-face RBF values
        std::vector<double> inv_all((size_t)N * 36), phi_all((size_t)N * 6);
#pragma omp parallel
        {
            std::vector<double> mat;
#pragma omp for schedule(static)
            for (int i = 0; i < N; i++) {
                const int n = sides(i);
                mat.assign((size_t)n * n, 0.0);
                for (int j1 = 0; j1 < n; j1++) {
                    const int fj = m.faces[(size_t)i * 6 + j1];
                    for (int j2 = 0; j2 < n; j2++) {
                        const int fi = m.faces[(size_t)i * 6 + j2];
                        const double arc = arc_angle_atan2(face_centre(fj), face_centre(fi));
                        const double phi = rbf(arc, 1.0);
                        mat[(size_t)j1 * n + j2] = phi * (normal_xyz[(size_t)fj * 3] * normal_xyz[(size_t)fi * 3] + normal_xyz[(size_t)fj * 3 + 1] * normal_xyz[(size_t)fi * 3 + 1] +
                                                          normal_xyz[(size_t)fj * 3 + 2] * normal_xyz[(size_t)fi * 3 + 2]);
                    }
                    phi_all[(size_t)i * 6 + j1] = rbf(arc_angle_atan2(node(i), face_centre(fj)), 1.0);      // node_face_RBF
                }
                const std::vector<double> inv = dense_inverse(mat, n);
                for (int x = 0; x < n; x++)
                    for (int y = 0; y < n; y++) inv_all[(size_t)i * 36 + (size_t)x * 6 + y] = inv[(size_t)x * n + y];
            }
        }
        if (build_csr(out.rbf_interp, 3 * N, F, 6, false, [&](int r, RowAcc& row) {
                const int i = r / 3, c = r % 3, n = sides(i);
                for (int x = 0; x < n; x++) {                                      // RBF row entries, ascending slot
                    const double a = phi_all[(size_t)i * 6 + x] * normal_xyz[(size_t)m.faces[(size_t)i * 6 + x] * 3 + c];
                    // (Vinv * adj) row (i, x): one entry per face of the cell, ascending face id
                    RowAcc t;
                    for (int y = 0; y < n; y++) t.add(m.faces[(size_t)i * 6 + y], inv_all[(size_t)i * 36 + (size_t)x * 6 + y] * 1.0);
                    t.sort();
                    for (auto& p : t.e) row.add(p.first, a * p.second);
                }
            })) { err = "operatorRBFinterp row too wide"; return -3; }
    }

    // ---- operatorDirectionalSecondDeriv = N * R * (r^-2 rbfDeriv2 * vandermondeInv * node2nodeAdj), mesh.cpp:280-345, :2364-2719
    {
        const double r_recip = 1.0 / r, s2 = r_recip * r_recip;
        // second-derivative rows (xx, xy, yy) at every cell over the cell and its neighbours
        // stored as fixed-width rows: [3N][7] (column, value), sorted by column
        std::vector<int> sd_col((size_t)3 * N * 7, -1);
        std::vector<double> sd_val((size_t)3 * N * 7, 0.0);
#pragma omp parallel
        {
        std::vector<double> mat, nn_rbf(7);
#pragma omp for schedule(static)
        for (int i = 0; i < N; i++) {
            const int n = sides(i), n1 = n + 1;
            mat.assign((size_t)n1 * n1, 0.0);
            mat[0] = 1.0;
            for (int j1 = 0; j1 < n; j1++) {                                       // mesh.cpp:291-300
                const double x1 = pos_map[(size_t)i * 14 + (size_t)(j1 + 1) * 2] / r, y1 = pos_map[(size_t)i * 14 + (size_t)(j1 + 1) * 2 + 1] / r;
                mat[(size_t)j1 + 1] = rbf(std::sqrt(x1 * x1 + y1 * y1), rbf_eps);
            }
            for (int j1 = 0; j1 < n; j1++) {                                       // :303-343
                const double x1 = pos_map[(size_t)i * 14 + (size_t)(j1 + 1) * 2] / r, y1 = pos_map[(size_t)i * 14 + (size_t)(j1 + 1) * 2 + 1] / r;
                const double arc = std::sqrt(x1 * x1 + y1 * y1);
                mat[(size_t)(j1 + 1) * n1] = rbf(arc, rbf_eps);
                nn_rbf[(size_t)j1 + 1] = rbf(arc, rbf_eps);
                for (int j2 = 0; j2 < n; j2++) {
                    const double x2 = pos_map[(size_t)i * 14 + (size_t)(j2 + 1) * 2] / r, y2 = pos_map[(size_t)i * 14 + (size_t)(j2 + 1) * 2 + 1] / r;
                    mat[(size_t)(j1 + 1) * n1 + j2 + 1] = rbf(std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2)), rbf_eps);
                }
            }
            nn_rbf[0] = 1.0;
            const std::vector<double> inv = dense_inverse(mat, n1);
            // rbfDeriv2 rows of this cell, scaled by r^-2 (mesh.cpp:2481-2545): slot 0 = the cell itself — absent from the xy row
            double d[3][7];
            bool has[3][7];
            for (int c = 0; c < 3; c++)
                for (int x = 0; x < 7; x++) { d[c][x] = 0.0; has[c][x] = false; }
            const double eps = rbf_eps;
            d[0][0] = -2.0 * eps * eps; has[0][0] = true;
            d[2][0] = -2.0 * eps * eps; has[2][0] = true;
            for (int j = 0; j < n; j++) {
                const double x1 = pos_map[(size_t)i * 14 + (size_t)(j + 1) * 2] * r_recip, y1 = pos_map[(size_t)i * 14 + (size_t)(j + 1) * 2 + 1] * r_recip;
                const double arc = std::sqrt(x1 * x1 + y1 * y1);
                const double xx = x1 * x1, yy = y1 * y1, xy = x1 * y1;
                const double arc2 = 1.0 / (arc * arc), arc3 = 1.0 / (arc * arc * arc);
                const double rb = nn_rbf[(size_t)j + 1];
                const double d2phidr2 = 2 * eps * eps * rb * (2 * arc * arc * eps * eps - 1);
                const double dphidr = -2 * arc * eps * eps * rb;
                d[0][j + 1] = (yy * arc3 * dphidr + xx * arc2 * d2phidr2); has[0][j + 1] = true;
                d[1][j + 1] = (-xy * arc3 * dphidr + xy * arc2 * d2phidr2); has[1][j + 1] = true;
                d[2][j + 1] = (xx * arc3 * dphidr + yy * arc2 * d2phidr2); has[2][j + 1] = true;
            }
            for (int c = 0; c < 3; c++) {
                // W = (s2 * rbfDeriv2) * vandermondeInv: columns (i, y), each the sum over the row's slots x in ascending order
                double wv[7];
                bool first[7];
                for (int y = 0; y < n1; y++) { wv[y] = 0.0; first[y] = true; }
                for (int x = 0; x < n1; x++) {
                    if (!has[c][x]) continue;
                    const double a = s2 * d[c][x];
                    for (int y = 0; y < n1; y++) {
                        const double t = a * inv[(size_t)x * n1 + y];
                        if (first[y]) { wv[y] = t; first[y] = false; }
                        else wv[y] += t;
                    }
                }
                // ... * node2nodeAdj: slot y -> the cell itself (y = 0) or its y-th neighbour
                RowAcc row;
                for (int y = 0; y < n1; y++) row.add(y == 0 ? i : m.node_friends[(size_t)i * 6 + y - 1], wv[y] * 1.0);
                row.sort();
                for (size_t q = 0; q < row.e.size(); q++) {
                    sd_col[((size_t)3 * i + c) * 7 + q] = row.e[q].first;
                    sd_val[((size_t)3 * i + c) * 7 + q] = row.e[q].second;
                }
            }
        }
        }
        if (build_csr(out.directional_second_deriv, 2 * F, N, 7, true, [&](int r2, RowAcc& row) {
                const int e = r2 / 2, k = r2 % 2;
                const double nx = m.face_normal_vec_map[(size_t)e * 2], ny = m.face_normal_vec_map[(size_t)e * 2 + 1];
                const double nc[3] = {nx * nx, 2 * nx * ny, ny * ny};              // N rows, mesh.cpp:2690-2698
                const int cell = m.face_nodes[(size_t)e * 2 + k];
                const double cosa = vel_trans[(size_t)e * 4 + (size_t)k * 2], sina = vel_trans[(size_t)e * 4 + (size_t)k * 2 + 1];
                const double R[3][3] = {{cosa * cosa, -2 * cosa * sina, sina * sina},          // mesh.cpp:2585-2680
                                        {cosa * sina, cosa * cosa - sina * sina, -sina * cosa},
                                        {sina * sina, 2 * cosa * sina, cosa * cosa}};
                double nr[3];                                                      // (N * R) row: ascending q
                for (int c = 0; c < 3; c++) {
                    nr[c] = nc[0] * R[0][c];
                    nr[c] += nc[1] * R[1][c];
                    nr[c] += nc[2] * R[2][c];
                }
                for (int c = 0; c < 3; c++)
                    for (int q = 0; q < 7; q++) {
                        const int col = sd_col[((size_t)3 * cell + c) * 7 + q];
                        if (col >= 0) row.add(col, nr[c] * sd_val[((size_t)3 * cell + c) * 7 + q]);
                    }
            })) { err = "operatorDirectionalSecondDeriv row too wide"; return -3; }                  // prune(0.0), mesh.cpp:2717
    }
    return 0;
}

}  // namespace odis
