// Locality renumbering of cells and edges for the device tables.
//
// The reference numbers cells hierarchically (each refinement level appends its new cells), so a
// cell's neighbours sit O(N) apart in memory and every stencil gather is a cache
// miss. The device numbering walks a Hilbert curve over an octahedral unfolding of the sphere, so
// consecutive threads touch a compact patch of cells and edges and the gathers hit L1/L2.
// Host code keeps both permutations; every field crossing the C ABI is in reference numbering.
#pragma once
#include <vector>
namespace odis {
// perm[new] = old. identity == true returns 0..n-1.
std::vector<int> cell_locality_order(int n_cells, const double* node_pos_sph /*[N][2] rad*/, bool identity);
// Edges follow their lower-ranked cell: sorted by (min new cell id, max new cell id).
std::vector<int> edge_locality_order(int n_edges, const int* face_nodes /*[F][2] old ids*/,
                                     const std::vector<int>& cell_new_of_old, bool identity);
std::vector<int> invert_permutation(const std::vector<int>& perm);
}  // namespace odis
