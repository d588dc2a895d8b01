// Host-side C-grid tables for the LTE hot path: grid_lN.txt reader and the table builder.
//
// Replaces the parts of the reference `Mesh` class (/root/reference/include/mesh.h:29-249,
// /root/reference/src/mesh.cpp) that feed ab3Explicit. Differences by design:
//   * sizes come from the grid file at run time (the reference bakes NODE_NUM/FACE_NUM in at
//     compile time, constants/gridConstants.h:17-46);
//   * only tables the solver reads are built, each as one flat std::vector in the
//     reference's row-major layout so they can cross the C ABI as plain pointers;
//   * cell->edge / cell->vertex connectivity is derived combinatorially from the neighbour
//     lists (two counting passes + fills that run in parallel) instead of by first-seen
//     insertion and floating-point angle sorts; the resulting integer tables are identical
//     to the reference's (tests/test_mesh_parity.py) and the FP64 tables use the same
//     expressions in the same order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace odis {

struct GridFile {                       // contents of input_files/grid_l<L>.txt (mesh.cpp:4016-4101)
    int n_cells = 0;
    std::vector<double> node_pos_sph;       // [N][2] lat, lon (rad)
    std::vector<int> node_friends;          // [N][6] neighbour ids, -1 pad for the 12 pentagons
    std::vector<double> centroid_pos_sph;   // [N][6][2] Voronoi corner lat, lon (rad)
};

// Parses the text grid format. Returns 0 on success, <0 on error (message in err).
int read_grid_file(const std::string& path, GridFile& out, std::string& err);
// Writes the text grid format (degrees, "%.16f"), the inverse of read_grid_file.
int write_grid_file(const std::string& path, const GridFile& g, std::string& err);

struct MeshTables {
    int n_cells = 0, n_edges = 0, n_vertices = 0;
    double radius = 0.0;                    // sphere radius the metric tables were built for

    // ---- per cell (reference names; include/mesh.h:81-150)
    std::vector<double> node_pos_sph;       // [N][2]
    std::vector<int> node_friends;          // [N][6]
    std::vector<double> centroid_pos_sph;   // [N][6][2]
    std::vector<double> centroid_pos_map;   // [N][6][2]  corners in the cell-centred stereographic map
    std::vector<double> control_volume_surf_area_map;  // [N]  planar (mapped) cell area
    std::vector<int> faces;                 // [N][6] cell -> edge ids (-1 pad)
    std::vector<int> node_face_dir;         // [N][6] +1 if the cell is the edge's inner cell, -1 outer
    std::vector<int> vertexes;              // [N][6] cell -> vertex ids (-1 pad)
    // ---- per edge (include/mesh.h:154-172)
    std::vector<int> face_nodes;            // [F][2] inner, outer cell
    std::vector<int> face_vertexes;         // [F][2]
    std::vector<int> face_interp_friends;   // [F][10] TRiSK stencil (unused tail entries 0)
    std::vector<double> face_interp_weights;// [F][10]
    std::vector<double> face_len;           // [F] arc length of the Voronoi edge (m)
    std::vector<double> face_node_dist;     // [F] arc distance between the two cell centres (m)
    std::vector<double> face_centre_m;      // [F][2] map factors of the edge centre seen from each cell
    std::vector<double> face_centre_pos_sph;    // [F][2]
    std::vector<double> face_intercept_pos_sph; // [F][2]
    std::vector<double> face_area;          // [F] d_e * l_e
    std::vector<double> face_normal_vec_map;// [F][2]
    // ---- per vertex (include/mesh.h:174-183)
    std::vector<double> vertex_pos_sph;     // [V][2]
    std::vector<int> vertex_nodes;          // [V][3]
    std::vector<double> vertex_R;           // [V][3] kite-area fractions
};

// Builds every table above from a parsed grid file for a sphere of the given radius.
// threads <= 0 uses the OpenMP default. Returns 0, or <0 with a message in err.
int build_mesh_tables(const GridFile& grid, double radius, MeshTables& out, std::string& err, int threads = 0);

// Time step quantisation: smallest n = 100k with period/n <= target (mesh.cpp:1601-1618).
void quantise_time_step(double period, double target_dt, double* dt_out, int* steps_per_period_out);

}  // namespace odis
