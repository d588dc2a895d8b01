// Grid reader and C-grid table builder (see odis_mesh.h). Citations are to
// /root/reference/src/mesh.cpp unless another file is named.
#include "odis_mesh.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "odis_sphere.h"

namespace odis {

// ------------------------------------------------------------------ grid file I/O --------
namespace {

// Pulls the next numeric token out of a grid line. The format's separators are spaces,
// braces, brackets and commas (mesh.cpp:4038-4076 splits on exactly those).
inline const char* next_number(const char* p) {
    while (*p && !((*p >= '0' && *p <= '9') || *p == '-' || *p == '+' || *p == '.')) ++p;
    return p;
}

}  // namespace

int read_grid_file(const std::string& path, GridFile& out, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        err = "ERROR: GRID FILE NOT FOUND AT " + path;      // wording of mesh.cpp:4095
        return -1;
    }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (sz > 0 && std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
        std::fclose(f);
        err = "short read on " + path;
        return -2;
    }
    std::fclose(f);
    buf[(size_t)sz] = 0;

    // line starts, skipping the header line (mesh.cpp:4033)
    std::vector<const char*> lines;
    const char* p = buf.data();
    const char* end = buf.data() + sz;
    while (p < end && *p != '\n') ++p;
    if (p < end) ++p;
    while (p < end) {
        const char* q = p;
        bool blank = true;
        while (q < end && *q != '\n') {
            if (*q != ' ' && *q != '\t' && *q != '\r') blank = false;
            ++q;
        }
        if (!blank) lines.push_back(p);
        p = (q < end) ? q + 1 : q;
    }
    const int n = (int)lines.size();
    if (n < 12) {
        err = "grid file " + path + " holds fewer than 12 cells";
        return -3;
    }
    for (char* c = buf.data(); c < buf.data() + sz; ++c)
        if (*c == '\n') *c = 0;

    out.n_cells = n;
    out.node_pos_sph.assign((size_t)n * 2, 0.0);
    out.node_friends.assign((size_t)n * 6, -1);
    out.centroid_pos_sph.assign((size_t)n * 12, 0.0);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int l = 0; l < n; l++) {
        const char* s = lines[l];
        char* e = nullptr;
        s = next_number(s);
        const long id = std::strtol(s, &e, 10);               // COL 0: id (mesh.cpp:4040)
        if (e == s || id < 0 || id >= n) { bad++; continue; }
        s = e;
        double vals[2];
        for (int k = 0; k < 2; k++) {                          // lat, lon in degrees (mesh.cpp:4043-4046)
            s = next_number(s);
            vals[k] = std::strtod(s, &e);
            if (e == s) bad++;
            s = e;
        }
        out.node_pos_sph[(size_t)id * 2 + 0] = vals[0] * kRadPerDeg;
        out.node_pos_sph[(size_t)id * 2 + 1] = vals[1] * kRadPerDeg;
        for (int k = 0; k < 6; k++) {                          // neighbour ids (mesh.cpp:4049-4055)
            s = next_number(s);
            const long fr = std::strtol(s, &e, 10);
            if (e == s) bad++;
            s = e;
            out.node_friends[(size_t)id * 6 + k] = (int)fr;
        }
        for (int k = 0; k < 12; k++) {                         // corner lat, lon (mesh.cpp:4059-4074)
            s = next_number(s);
            const double v = std::strtod(s, &e);
            if (e == s) bad++;
            s = e;
            out.centroid_pos_sph[(size_t)id * 12 + k] = v * kRadPerDeg;
        }
    }
    if (bad) {
        err = "malformed line(s) in " + path;
        return -4;
    }
    return 0;
}

namespace {
// Degrees whose "%.16f" text parses back (x pi/180) to exactly `rad`, when such a value exists within a
// few ulps of rad*180/pi (true for every grid that was itself read from / generated for this format).
double degrees_for_text(double rad) {
    const double d0 = rad * (180.0 / kPi);
    char buf[64];
    for (int k = 0; k < 9; k++) {
        double cand = d0;
        const int steps = (k + 1) / 2;
        for (int s = 0; s < steps; s++) cand = std::nextafter(cand, (k % 2) ? INFINITY : -INFINITY);
        std::snprintf(buf, sizeof buf, "%.16f", cand);
        const double back = std::strtod(buf, nullptr);
        if (back * kRadPerDeg == rad) return back;
    }
    return d0;
}
}  // namespace

int write_grid_file(const std::string& path, const GridFile& g, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        err = "cannot open " + path + " for writing";
        return -1;
    }
    std::fprintf(f, "ID    NODE LAT     NODE LON     FRIENDS LIST                           CENTROID COORD LIST \n");
    for (int i = 0; i < g.n_cells; i++) {
        std::fprintf(f, "%-5d %.16f %.16f {", i, degrees_for_text(g.node_pos_sph[(size_t)i * 2]), degrees_for_text(g.node_pos_sph[(size_t)i * 2 + 1]));
        for (int j = 0; j < 6; j++) std::fprintf(f, "%5d%s", g.node_friends[(size_t)i * 6 + j], j < 5 ? "," : "}, {");
        for (int j = 0; j < 6; j++) {
            const bool pad = g.node_friends[(size_t)i * 6 + j] < 0;
            const double la = pad ? -1.0 : degrees_for_text(g.centroid_pos_sph[(size_t)i * 12 + 2 * j]);
            const double lo = pad ? -1.0 : degrees_for_text(g.centroid_pos_sph[(size_t)i * 12 + 2 * j + 1]);
            std::fprintf(f, "( %.16f, %.16f)%s", la, lo, j < 5 ? ", " : "} \n");
        }
    }
    std::fclose(f);
    return 0;
}

void quantise_time_step(double period, double target_dt, double* dt_out, int* steps_out) {
    // mesh.cpp:1601-1618
    double dt = period;
    int dt_num = 100;
    while (dt > target_dt) {
        dt = period / dt_num;
        dt_num += 100;
    }
    dt_num -= 100;
    *dt_out = dt;
    *steps_out = dt_num;
}

// ------------------------------------------------------------------ table builder --------
namespace {

struct Builder {
    const GridFile& g;
    MeshTables& m;
    const double r;
    const int N;
    std::vector<int> sides;        // 5 or 6 per cell

    Builder(const GridFile& grid, MeshTables& out, double radius)
        : g(grid), m(out), r(radius), N(grid.n_cells), sides((size_t)grid.n_cells) {}

    LatLon node(int i) const { return LatLon{m.node_pos_sph[(size_t)i * 2], m.node_pos_sph[(size_t)i * 2 + 1]}; }
    LatLon corner(int i, int j) const {
        return LatLon{m.centroid_pos_sph[(size_t)i * 12 + 2 * j], m.centroid_pos_sph[(size_t)i * 12 + 2 * j + 1]};
    }
    LatLon vertex(int v) const { return LatLon{m.vertex_pos_sph[(size_t)v * 2], m.vertex_pos_sph[(size_t)v * 2 + 1]}; }
    LatLon intercept(int e) const {
        return LatLon{m.face_intercept_pos_sph[(size_t)e * 2], m.face_intercept_pos_sph[(size_t)e * 2 + 1]};
    }
    int fr(int i, int j) const { return m.node_friends[(size_t)i * 6 + j]; }
    int slot_of_friend(int c, int who) const {
        for (int j = 0; j < sides[c]; j++)
            if (fr(c, j) == who) return j;
        return -1;
    }
    int slot_of_edge(int c, int e) const {
        for (int j = 0; j < sides[c]; j++)
            if (m.faces[(size_t)c * 6 + j] == e) return j;
        return -1;
    }

    int check_input(std::string& err) {
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int i = 0; i < N; i++) {
            const int n = (fr(i, 5) < 0) ? 5 : 6;               // mesh.cpp:452-453 and every loop after
            sides[i] = n;
            for (int j = 0; j < n; j++) {
                const int f = fr(i, j);
                if (f < 0 || f >= N || f == i) { bad++; continue; }
                bool back = false;
                const int nf = (fr(f, 5) < 0) ? 5 : 6;
                for (int k = 0; k < nf; k++) back |= (fr(f, k) == i);
                if (!back) bad++;
            }
        }
        if (bad) {
            err = "grid neighbour lists are not symmetric / in range";
            return -10;
        }
        if (3L * N - 6 > 2147483647L) {
            err = "grid too large for 32-bit edge ids";
            return -11;
        }
        return 0;
    }

    // Cell-centred stereographic maps of the Voronoi corners and the planar cell area.
    // mesh.cpp:1453-1481 (CalcMappingCoords, corner part), :1828-1867 (CalcControlVolumeArea)
    void cell_maps_and_areas() {
        m.centroid_pos_map.assign((size_t)N * 12, -1.0);
        m.control_volume_surf_area_map.assign((size_t)N, 0.0);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) {
            const LatLon c = node(i);
            double* cm = &m.centroid_pos_map[(size_t)i * 12];
            for (int j = 0; j < 6; j++) {
                if (fr(i, j) == -1) continue;                      // pentagon pad keeps (-1,-1)
                const Vec2 q = map_project(c, corner(i, j), r);
                cm[2 * j] = q.x;
                cm[2 * j + 1] = q.y;
            }
            const Vec2 centre = map_project(c, c, r);              // node_pos_map(i,0,:), mesh.cpp:1398-1403
            const int n = sides[i];
            double area = 0.0;
            for (int j = 0; j < n; j++) {
                const int j2 = (j + 1) % n;
                area += planar_triangle_area(centre.x, centre.y, cm[2 * j], cm[2 * j2], cm[2 * j + 1], cm[2 * j2 + 1]);
            }
            m.control_volume_surf_area_map[i] = area;
        }
    }

    // Vertex numbering. The reference numbers a Voronoi corner the first time the cell loop
    // meets it and finds the two other owners by coordinate matching (mesh.cpp:451-497);
    // "first met" is the lowest-numbered of the three cells, so ids follow a prefix sum.
    int number_vertices(std::string& err) {
        std::vector<int> first((size_t)N + 1, 0);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) {
            const int n = sides[i];
            int cnt = 0;
            for (int j = 0; j < n; j++) cnt += (i < fr(i, j) && i < fr(i, (j + 1) % n));
            first[(size_t)i + 1] = cnt;
        }
        for (int i = 0; i < N; i++) first[(size_t)i + 1] += first[i];
        const int V = first[N];
        if (V != 2 * N - 4) {
            err = "grid is not a closed triangulation (vertex count " + std::to_string(V) + ", expected " + std::to_string(2 * N - 4) + ")";
            return -12;
        }
        m.n_vertices = V;
        m.vertexes.assign((size_t)N * 6, -1);
        m.vertex_pos_sph.assign((size_t)V * 2, 0.0);
        m.vertex_nodes.assign((size_t)V * 3, -1);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int i = 0; i < N; i++) {
            const int n = sides[i];
            int id = first[i];
            for (int j = 0; j < n; j++) {
                const int f1 = fr(i, j), f2 = fr(i, (j + 1) % n);
                if (!(i < f1 && i < f2)) continue;
                m.vertexes[(size_t)i * 6 + j] = id;
                m.vertex_pos_sph[(size_t)id * 2] = m.centroid_pos_sph[(size_t)i * 12 + 2 * j];
                m.vertex_pos_sph[(size_t)id * 2 + 1] = m.centroid_pos_sph[(size_t)i * 12 + 2 * j + 1];
                m.vertex_nodes[(size_t)id * 3 + 0] = i;
                m.vertex_nodes[(size_t)id * 3 + 1] = f1;
                m.vertex_nodes[(size_t)id * 3 + 2] = f2;
                // the same corner in the two other cells: the slot whose two flanking
                // neighbours are the remaining pair
                const int others[2][3] = {{f1, i, f2}, {f2, i, f1}};
                for (int t = 0; t < 2; t++) {
                    const int c = others[t][0], a = others[t][1], b = others[t][2];
                    const int nc = sides[c];
                    int hit = -1;
                    for (int k = 0; k < nc; k++) {
                        const int p = fr(c, k), q = fr(c, (k + 1) % nc);
                        if ((p == a && q == b) || (p == b && q == a)) hit = k;
                    }
                    if (hit < 0) { bad++; continue; }
                    m.vertexes[(size_t)c * 6 + hit] = id;
                }
                id++;
            }
        }
        if (bad) {
            err = "inconsistent corner ordering between neighbouring cells";
            return -13;
        }
        return 0;
    }

    // Edge numbering + per-edge metric terms. First-seen numbering (mesh.cpp:533-692) means
    // the edge between i and f belongs to min(i,f) and ids follow a prefix sum over cells.
    int number_edges(std::string& err) {
        std::vector<int> first((size_t)N + 1, 0);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) {
            int cnt = 0;
            for (int j = 0; j < sides[i]; j++) cnt += (fr(i, j) > i);
            first[(size_t)i + 1] = cnt;
        }
        for (int i = 0; i < N; i++) first[(size_t)i + 1] += first[i];
        const int F = first[N];
        if (F != 3 * N - 6) {
            err = "edge count " + std::to_string(F) + " != 3N-6";
            return -14;
        }
        m.n_edges = F;
        m.faces.assign((size_t)N * 6, -1);
        m.node_face_dir.assign((size_t)N * 6, 0);
        m.face_nodes.assign((size_t)F * 2, -1);
        m.face_len.assign((size_t)F, 0.0);
        m.face_node_dist.assign((size_t)F, 0.0);
        m.face_centre_m.assign((size_t)F * 2, 0.0);
        m.face_centre_pos_sph.assign((size_t)F * 2, 0.0);
        m.face_intercept_pos_sph.assign((size_t)F * 2, 0.0);
        m.face_normal_vec_map.assign((size_t)F * 2, 0.0);
        m.face_area.assign((size_t)F, 0.0);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int i = 0; i < N; i++) {
            const int n = sides[i];
            int e = first[i];
            for (int j = 0; j < n; j++) {
                const int f = fr(i, j);
                if (f < i) continue;
                m.faces[(size_t)i * 6 + j] = e;
                m.node_face_dir[(size_t)i * 6 + j] = 1;
                const int j2 = slot_of_friend(f, i);
                if (j2 < 0) { bad++; continue; }
                m.faces[(size_t)f * 6 + j2] = e;
                m.node_face_dir[(size_t)f * 6 + j2] = -1;
                m.face_nodes[(size_t)e * 2] = i;
                m.face_nodes[(size_t)e * 2 + 1] = f;

                // the edge joins corners j-1 and j of its inner cell (mesh.cpp:552-555)
                const LatLon a = corner(i, (j + n - 1) % n), b = corner(i, j);
                const LatLon ni = node(i), nf = node(f);
                m.face_len[e] = std::fabs(arc_length_acos(a, b, r));                    // :559-561
                LatLon c = chord_midpoint(a, b);                                        // :564-569
                if (c.lon < 0.0) c.lon += 2 * kPi;
                m.face_centre_pos_sph[(size_t)e * 2] = c.lat;
                m.face_centre_pos_sph[(size_t)e * 2 + 1] = c.lon;
                m.face_centre_m[(size_t)e * 2] = map_factor(c, ni);                     // :604-613
                m.face_centre_m[(size_t)e * 2 + 1] = map_factor(c, nf);
                m.face_node_dist[e] = arc_angle_atan2(ni, nf) * r;                      // :617-623
                const Vec2 nrm = edge_normal_in_map(a, b);                              // :634-637
                m.face_normal_vec_map[(size_t)e * 2] = nrm.x;
                m.face_normal_vec_map[(size_t)e * 2 + 1] = nrm.y;
                const LatLon x = great_circle_intersection(a, b, ni, nf);               // :661-669
                m.face_intercept_pos_sph[(size_t)e * 2] = x.lat;
                m.face_intercept_pos_sph[(size_t)e * 2 + 1] = x.lon;
                m.face_area[e] = m.face_node_dist[e] * m.face_len[e];                   // :1093
                e++;
            }
        }
        if (bad) {
            err = "asymmetric neighbour list while numbering edges";
            return -15;
        }
        return 0;
    }

    // The reference orders a cell's edges and corners around the cell by sorting map angles
    // measured from the current edge, descending, after shifting positive angles by -360 deg
    // (mesh.cpp:734-782). With the neighbour lists running clockwise seen from outside (the
    // only orientation for which the reference's intersection points land on the near side
    // of the sphere) that order is: edges k, k-1, k-2, ...; corners k-1, k-2, ... . One angle
    // per cell verifies the orientation instead of sorting 2F small lists.
    int check_orientation(std::string& err) {
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int c = 0; c < N; c++) {
            const int e0 = m.faces[(size_t)c * 6], e1 = m.faces[(size_t)c * 6 + 1];
            const double deg = map_angle_between(intercept(e0), intercept(e1), node(c)) * 180. / kPi;
            if (!(deg > 0.0 + 1e-8)) bad++;
        }
        if (bad) {
            err = "neighbour lists must run clockwise (seen from outside the sphere) around every cell";
            return -16;
        }
        return 0;
    }

    // TRiSK tangential-reconstruction stencil and weights (mesh.cpp:717-910), plus the
    // edge->vertex and vertex kite-fraction tables the same loop fills (:826-829, :858-866).
    void trisk_weights() {
        const int F = m.n_edges, V = m.n_vertices;
        m.face_interp_friends.assign((size_t)F * 10, 0);
        m.face_interp_weights.assign((size_t)F * 10, 0.0);
        m.face_vertexes.assign((size_t)F * 2, -1);
        m.vertex_R.assign((size_t)V * 3, 0.0);
        std::vector<int> last_edge((size_t)N);
#pragma omp parallel for schedule(static)
        for (int c = 0; c < N; c++) {
            int mx = -1;
            for (int j = 0; j < sides[c]; j++) mx = std::max(mx, m.faces[(size_t)c * 6 + j]);
            last_edge[c] = mx;
        }
#pragma omp parallel for schedule(static)
        for (int e = 0; e < F; e++) {
            int j_add = 0;
            for (int side = 0; side < 2; side++) {
                const int c = m.face_nodes[(size_t)e * 2 + side];
                const int n = sides[c];
                const int k = slot_of_edge(c, e);
                const LatLon nc = node(c);
                int ering[6], vring[6];
                for (int j = 0; j < n; j++) {
                    ering[j] = m.faces[(size_t)c * 6 + (k - j + n) % n];
                    vring[j] = m.vertexes[(size_t)c * 6 + (k - 1 - j + 2 * n) % n];
                }
                double area_cv = 0.0;                                            // :786-813
                for (int j = 0; j < n; j++)
                    area_cv += spherical_triangle_area(nc, vertex(vring[j]), vertex(vring[(j + 1) % n]), r);
                double R[6];                                                     // :815-870
                for (int j = 0; j < n; j++) {
                    const LatLon vj = vertex(vring[j]);
                    const double a1 = spherical_triangle_area(nc, vj, intercept(ering[j]), r);
                    const double a2 = spherical_triangle_area(nc, vj, intercept(ering[(j + 1) % n]), r);
                    R[j] = (a1 + a2) / area_cv;
                }
                // edge -> vertex: every (edge, cell) visit writes the entry of ring edge 1;
                // the reference's loop order makes the lexicographically last visit win.
                {
                    const int f = ering[1];
                    const int c2 = (m.face_nodes[(size_t)f * 2] == c) ? m.face_nodes[(size_t)f * 2 + 1] : m.face_nodes[(size_t)f * 2];
                    const int n2 = sides[c2];
                    const int p2 = slot_of_edge(c2, f);
                    const int e2 = m.faces[(size_t)c2 * 6 + (p2 + 1) % n2];
                    const int side2 = (m.face_nodes[(size_t)e2 * 2] == c2) ? 0 : 1;
                    if (e > e2 || (e == e2 && side > side2)) {
                        m.face_vertexes[(size_t)f * 2] = vring[1];
                        m.face_vertexes[(size_t)f * 2 + 1] = vring[0];
                    }
                }
                if (e == last_edge[c]) {                                         // last visit of cell c
                    for (int j = 0; j < n; j++) {
                        const int v = vring[j];
                        for (int s = 0; s < 3; s++)
                            if (m.vertex_nodes[(size_t)v * 3 + s] == c) { m.vertex_R[(size_t)v * 3 + s] = R[j]; break; }
                    }
                }
                const int t_ev = (side == 0) ? 1 : -1;                           // :898-900
                for (int j = 1; j < n; j++) {                                    // :873-908
                    double w = 0.0;
                    for (int j2 = 0; j2 < j; j2++) w += R[j2];
                    w -= 0.5;
                    w *= m.node_face_dir[(size_t)c * 6 + (k - j + n) % n];
                    w *= t_ev;
                    m.face_interp_friends[(size_t)e * 10 + j_add] = ering[j];
                    m.face_interp_weights[(size_t)e * 10 + j_add] = w;
                    j_add++;
                }
            }
        }
    }
};

}  // namespace

int build_mesh_tables(const GridFile& grid, double radius, MeshTables& out, std::string& err, int threads) {
#ifdef _OPENMP
    const int prev = omp_get_max_threads();
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    out = MeshTables();
    out.n_cells = grid.n_cells;
    out.radius = radius;
    out.node_pos_sph = grid.node_pos_sph;
    out.node_friends = grid.node_friends;
    out.centroid_pos_sph = grid.centroid_pos_sph;
    Builder b(grid, out, radius);
    int rc = b.check_input(err);
    if (!rc) { b.cell_maps_and_areas(); rc = b.number_vertices(err); }
    if (!rc) rc = b.number_edges(err);
    if (!rc) rc = b.check_orientation(err);
    if (!rc) b.trisk_weights();
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(prev);
#endif
    return rc;
}

}  // namespace odis
