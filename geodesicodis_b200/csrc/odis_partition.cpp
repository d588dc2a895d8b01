#include "odis_partition.h"

#include "odis_reorder.h"

#include <algorithm>

namespace odis {

int Partition::cell_owner(int c) const {
    return (int)(std::upper_bound(cell_begin.begin(), cell_begin.end(), c) - cell_begin.begin()) - 1;
}
int Partition::edge_owner(int e) const {
    return (int)(std::upper_bound(edge_begin.begin(), edge_begin.end(), e) - edge_begin.begin()) - 1;
}

namespace {
// ghost sets of rank q: cells and edges it reads but does not own (ascending global ids)
void ghosts_of(const Partition& P, int q, int n_cells, const int* edge_cells, const int* cell_edges, std::vector<int>& gcells,
               std::vector<int>& gedges) {
    gcells.clear();
    gedges.clear();
    const int c0 = P.cell_begin[q], c1 = P.cell_begin[q + 1], e0 = P.edge_begin[q], e1 = P.edge_begin[q + 1];
    std::vector<int> touched;                                   // cells whose edges rank q reads
    for (int e = e0; e < e1; e++)
        for (int k = 0; k < 2; k++) {
            const int c = edge_cells[(size_t)e * 2 + k];
            if (c < c0 || c >= c1) gcells.push_back(c);
        }
    std::sort(gcells.begin(), gcells.end());
    gcells.erase(std::unique(gcells.begin(), gcells.end()), gcells.end());
    auto add_edges_of = [&](int c) {
        for (int j = 0; j < 6; j++) {
            const int e = cell_edges[(size_t)c * 6 + j];
            if (e >= 0 && (e < e0 || e >= e1)) gedges.push_back(e);
        }
    };
    for (int c = c0; c < c1; c++) add_edges_of(c);
    for (int c : gcells) add_edges_of(c);
    std::sort(gedges.begin(), gedges.end());
    gedges.erase(std::unique(gedges.begin(), gedges.end()), gedges.end());
    (void)n_cells;
}
}  // namespace

Partition build_partition(int n_cells, int n_edges, const int* edge_cells, const int* cell_edges, int world, int rank) {
    Partition P;
    P.world = world;
    P.rank = rank;
    P.cell_begin.resize((size_t)world + 1);
    P.edge_begin.resize((size_t)world + 1);
    for (int r = 0; r <= world; r++) P.cell_begin[r] = (int)((long long)n_cells * r / world);
    // edges are sorted by their lower cell, so a rank's edges are those whose lower cell lies in its range
    P.edge_begin[0] = 0;
    {
        int e = 0;
        for (int r = 1; r <= world; r++) {
            while (e < n_edges && std::min(edge_cells[(size_t)e * 2], edge_cells[(size_t)e * 2 + 1]) < P.cell_begin[r]) e++;
            P.edge_begin[r] = e;
        }
        P.edge_begin[world] = n_edges;
    }
    P.n_own_cells = P.cell_begin[rank + 1] - P.cell_begin[rank];
    P.n_own_edges = P.edge_begin[rank + 1] - P.edge_begin[rank];

    std::vector<int> gcells, gedges;
    ghosts_of(P, rank, n_cells, edge_cells, cell_edges, gcells, gedges);
    for (int c = P.cell_begin[rank]; c < P.cell_begin[rank + 1]; c++) P.local_cells.push_back(c);
    P.local_cells.insert(P.local_cells.end(), gcells.begin(), gcells.end());
    for (int e = P.edge_begin[rank]; e < P.edge_begin[rank + 1]; e++) P.local_edges.push_back(e);
    P.local_edges.insert(P.local_edges.end(), gedges.begin(), gedges.end());

    // who fills my ghosts
    std::vector<int> recv_e((size_t)world, 0), recv_c((size_t)world, 0);
    for (int c : gcells) recv_c[P.cell_owner(c)]++;
    for (int e : gedges) recv_e[P.edge_owner(e)]++;
    // what I fill in every other rank's halo (same derivation, run for that rank)
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        std::vector<int> qc, qe;
        ghosts_of(P, q, n_cells, edge_cells, cell_edges, qc, qe);
        HaloPeer peer;
        peer.rank = q;
        const int q_own_c = P.cell_begin[q + 1] - P.cell_begin[q], q_own_e = P.edge_begin[q + 1] - P.edge_begin[q];
        for (size_t k = 0; k < qc.size(); k++)
            if (qc[k] >= P.cell_begin[rank] && qc[k] < P.cell_begin[rank + 1]) {
                peer.send_cell_local.push_back(qc[k] - P.cell_begin[rank]);
                peer.send_cell_remote.push_back(q_own_c + (int)k);
            }
        for (size_t k = 0; k < qe.size(); k++)
            if (qe[k] >= P.edge_begin[rank] && qe[k] < P.edge_begin[rank + 1]) {
                peer.send_edge_local.push_back(qe[k] - P.edge_begin[rank]);
                peer.send_edge_remote.push_back(q_own_e + (int)k);
            }
        peer.recv_cells = recv_c[q];
        peer.recv_edges = recv_e[q];
        if (!peer.send_cell_local.empty() || !peer.send_edge_local.empty() || peer.recv_cells || peer.recv_edges) P.peers.push_back(std::move(peer));
    }

    // ---- boundary placement inside the own ranges (purely local: ghost slots keep their positions) ----
    // edges: boundary FIRST (sent to a neighbour, or touching a ghost) - the edge kernel updates and pushes them before
    // the interior; cells: boundary LAST (a non-own edge in their divergence), next to the ghost cells, so that every cell
    // whose update has to wait for the neighbours' velocities sits at the end of the cell kernel's iteration space
    const int c0 = P.cell_begin[rank], c1 = P.cell_begin[rank + 1], e0 = P.edge_begin[rank], e1 = P.edge_begin[rank + 1];
    std::vector<char> bnd_c((size_t)P.n_own_cells, 0), bnd_e((size_t)P.n_own_edges, 0);
    for (const HaloPeer& peer : P.peers)
        for (int l : peer.send_edge_local) bnd_e[(size_t)l] = 1;
    auto cell_has_foreign_edge = [&](int c) {
        for (int j = 0; j < 6; j++) {
            const int e = cell_edges[(size_t)c * 6 + j];
            if (e >= 0 && (e < e0 || e >= e1)) return true;
        }
        return false;
    };
    for (int c = c0; c < c1; c++)
        if (cell_has_foreign_edge(c)) bnd_c[(size_t)(c - c0)] = 1;
    for (int e = e0; e < e1; e++)
        for (int k = 0; k < 2; k++) {
            const int c = edge_cells[(size_t)e * 2 + k];
            if (c < c0 || c >= c1 || cell_has_foreign_edge(c)) bnd_e[(size_t)(e - e0)] = 1;
        }
    auto place = [](const std::vector<char>& bnd, bool boundary_first, std::vector<int>& new_of_old) {
        const int n = (int)bnd.size();
        new_of_old.resize((size_t)n);
        int nb = 0;
        for (int i = 0; i < n; i++) nb += bnd[(size_t)i] ? 1 : 0;
        int kb = boundary_first ? 0 : n - nb, ki = boundary_first ? nb : 0;
        for (int i = 0; i < n; i++) new_of_old[(size_t)i] = bnd[(size_t)i] ? kb++ : ki++;
        return nb;
    };
    std::vector<int> cnew, enew;
    P.n_bnd_cells = place(bnd_c, false, cnew);
    P.n_bnd_edges = place(bnd_e, true, enew);
    for (int i = 0; i < P.n_own_cells; i++) P.local_cells[(size_t)cnew[(size_t)i]] = c0 + i;
    for (int i = 0; i < P.n_own_edges; i++) P.local_edges[(size_t)enew[(size_t)i]] = e0 + i;
    for (HaloPeer& peer : P.peers) {
        for (int& l : peer.send_cell_local) l = cnew[(size_t)l];
        for (int& l : peer.send_edge_local) l = enew[(size_t)l];
    }
    return P;
}

void build_local_numbering(int Ng, int Fg, const double* node_pos_sph, const int* face_nodes, const int* faces, bool reorder, int rank,
                           int world, LocalNumbering& L) {
    const bool identity = !reorder && world == 1;
    L.g_cell_perm = cell_locality_order(Ng, node_pos_sph, identity);
    L.g_cell_inv = invert_permutation(L.g_cell_perm);
    L.g_edge_perm = edge_locality_order(Fg, face_nodes, L.g_cell_inv, identity);
    L.g_edge_inv = invert_permutation(L.g_edge_perm);
    L.loc_cell.clear();
    L.loc_edge.clear();
    if (world == 1) {
        L.part = Partition();
        L.part.world = 1; L.part.rank = 0;
        L.part.n_own_cells = Ng; L.part.n_own_edges = Fg;
        L.cell_perm = L.g_cell_perm;
        L.edge_perm = L.g_edge_perm;
        return;
    }
    std::vector<int> edge_cells((size_t)Fg * 2), cell_edges((size_t)Ng * 6, -1);
#pragma omp parallel for schedule(static)
    for (int en = 0; en < Fg; en++) {
        const int eo = L.g_edge_perm[en];
        edge_cells[(size_t)en * 2] = L.g_cell_inv[face_nodes[(size_t)eo * 2]];
        edge_cells[(size_t)en * 2 + 1] = L.g_cell_inv[face_nodes[(size_t)eo * 2 + 1]];
    }
#pragma omp parallel for schedule(static)
    for (int cn = 0; cn < Ng; cn++) {
        const int co = L.g_cell_perm[cn];
        for (int j = 0; j < 6; j++) {
            const int e = faces[(size_t)co * 6 + j];
            cell_edges[(size_t)cn * 6 + j] = e < 0 ? -1 : L.g_edge_inv[e];
        }
    }
    L.part = build_partition(Ng, Fg, edge_cells.data(), cell_edges.data(), world, rank);
    L.loc_cell.assign((size_t)Ng, -1);
    L.loc_edge.assign((size_t)Fg, -1);
    L.cell_perm.resize(L.part.local_cells.size());
    L.edge_perm.resize(L.part.local_edges.size());
    for (size_t i = 0; i < L.part.local_cells.size(); i++) {
        L.loc_cell[(size_t)L.part.local_cells[i]] = (int)i;
        L.cell_perm[i] = L.g_cell_perm[(size_t)L.part.local_cells[i]];
    }
    for (size_t i = 0; i < L.part.local_edges.size(); i++) {
        L.loc_edge[(size_t)L.part.local_edges[i]] = (int)i;
        L.edge_perm[i] = L.g_edge_perm[(size_t)L.part.local_edges[i]];
    }
}

}  // namespace odis
