// Domain decomposition of the C-grid for multi-GPU runs (one rank per GPU).
//
// The reference is a single process (no MPI/NCCL anywhere); this is new. Cells are cut into
// `world` contiguous ranges of the device (space-filling-curve) numbering, so every part is a compact patch
// of the sphere; an edge belongs to the rank of its lower-numbered cell, which makes the edge ranges
// contiguous too. A rank stores its own cells/edges first and then a one-ring halo:
//   ghost cells  = the other-rank cells of its own edges                     (eta, U for the pressure gradient)
//   ghost edges  = the other-rank edges of all cells it touches              (v, l for the TRiSK stencil and
//                  (own cells and ghost cells)                                the divergence of its own cells)
// Per step there is ONE exchange: v of the boundary edges, pushed by the edge update itself. Ghost cells are
// not exchanged: every rank holds all edges of its ghost cells, so it repeats their (cheap, bit-identical)
// eta update locally from the exchanged velocities. The cell send lists below describe ownership (who holds
// the authoritative value of a ghost cell) for set/get and for tests. Every rank derives all halos deterministically from the same global tables, so
// send lists need no negotiation: the sender knows the receiver's ghost slot of every entity.
#pragma once
#include <vector>

namespace odis {

struct HaloPeer {
    int rank = -1;
    std::vector<int> send_edge_local;   // my local edge ids whose v the peer needs
    std::vector<int> send_edge_remote;  // the peer's local (ghost) ids for them
    std::vector<int> send_cell_local;
    std::vector<int> send_cell_remote;
    int recv_edges = 0, recv_cells = 0; // how many of my ghosts this peer fills
};

struct Partition {
    int world = 1, rank = 0;
    std::vector<int> cell_begin, edge_begin;   // [world+1] ranges in global device numbering
    int n_own_cells = 0, n_own_edges = 0;
    // Placement of the boundary inside the own ranges. Boundary edges (sent to a neighbour, or reading a ghost value: a
    // non-own cell, or a non-own edge of their two cells) come FIRST, [0, n_bnd_edges): the edge kernel updates and pushes
    // them before the interior, so the transfer overlaps the interior update. Boundary cells (a non-own edge in their
    // divergence) come LAST among the own cells, [n_own_cells - n_bnd_cells, n_own_cells), right before the ghost cells:
    // everything in the cell kernel that has to wait for the neighbours' velocities sits at the end of its iteration space.
    int n_bnd_cells = 0, n_bnd_edges = 0;
    std::vector<int> local_cells;              // global device ids: own (interior, then boundary), then ghosts ascending
    std::vector<int> local_edges;
    std::vector<HaloPeer> peers;               // ascending rank, only ranks that share a boundary
    int cell_owner(int c) const;
    int edge_owner(int e) const;
};

// edge_cells: [F][2] cells of each edge, cell_edges: [N][6] edges of each cell (-1 pad), both in global
// device numbering with edges sorted by (lower cell, higher cell).
Partition build_partition(int n_cells, int n_edges, const int* edge_cells, const int* cell_edges, int world, int rank);

// Everything a rank needs to know about numbering: the global device order and its own local order.
struct LocalNumbering {
    Partition part;
    std::vector<int> g_cell_perm, g_cell_inv, g_edge_perm, g_edge_inv;   // global: perm[device id] = reference id
    std::vector<int> cell_perm, edge_perm;                                // local id -> reference id (own first, then halo)
    std::vector<int> loc_cell, loc_edge;                                  // global device id -> local id, -1 if not held (world > 1 only)
    int local_cell_of_ref(int ref_id) const { const int g = g_cell_inv[(size_t)ref_id]; return part.world == 1 ? g : loc_cell[(size_t)g]; }
    int local_edge_of_ref(int ref_id) const { const int g = g_edge_inv[(size_t)ref_id]; return part.world == 1 ? g : loc_edge[(size_t)g]; }
};
// reorder == false keeps the reference order on a single rank; partitions always use the locality order.
void build_local_numbering(int n_cells, int n_edges, const double* node_pos_sph, const int* face_nodes, const int* faces, bool reorder,
                           int rank, int world, LocalNumbering& out);

}  // namespace odis
