// C ABI, host-only groups (config, grid, mesh) of include/odis_b200.h.
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "odis_analytic.h"
#include "../../include/odis_b200.h"
#include "odis_config.h"
#include "odis_error.h"
#include "odis_gridgen.h"
#include "odis_mesh.h"
#include "odis_mesh_nl.h"
#include "odis_partition.h"
#include "odis_sh.h"

struct odis_config {
    odis::Config cfg;
    bool finalized = false;
};
struct odis_mesh {
    odis::MeshTables t;
};
struct odis_nonlinear {
    odis::NonlinearTables t;
};

namespace odis {
thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
}  // namespace odis

using odis::fail;

extern "C" {

const char* odis_last_error(void) { return odis::g_last_error.c_str(); }
const char* odis_version(void) { return "odis_b200 0.1 sm_100a"; }

// ------------------------------------------------------------------ nonlinear-branch tables ----
int odis_nonlinear_create(const odis_mesh* mesh, double rbf_eps, odis_nonlinear** out) {
    if (!mesh || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    odis_nonlinear* n = new (std::nothrow) odis_nonlinear();
    if (!n) return fail(ODIS_ERR_ARG, "out of memory");
    std::string err;
    if (odis::build_nonlinear_tables(mesh->t, mesh->t.radius, rbf_eps, n->t, err) != 0) {
        delete n;
        return fail(ODIS_ERR_ARG, err);
    }
    *out = n;
    return ODIS_OK;
}

int odis_nonlinear_get_view(const odis_nonlinear* nl, odis_nonlinear_view* view) {
    if (!nl || !view) return fail(ODIS_ERR_ARG, "NULL argument");
    auto csr = [](const odis::Csr& A) {
        odis_csr_view v;
        v.n_rows = A.n_rows; v.n_cols = A.n_cols; v.indptr = A.indptr.data(); v.indices = A.indices.data(); v.data = A.data.data();
        return v;
    };
    view->curl = csr(nl->t.curl);
    view->rbf_interp = csr(nl->t.rbf_interp);
    view->directional_second_deriv = csr(nl->t.directional_second_deriv);
    view->vertex_sinlat = nl->t.vertex_sinlat.data();
    view->vertex_area = nl->t.vertex_area.data();
    return ODIS_OK;
}

void odis_nonlinear_free(odis_nonlinear* nl) { delete nl; }

// ------------------------------------------------------------------ spherical harmonics (host helpers) ----
int odis_sh_basis(int32_t n, const double* pos_sph, int32_t l_max, double* Y_out) {
    if (!pos_sph || !Y_out || n <= 0) return fail(ODIS_ERR_ARG, "NULL or empty argument");
    if (l_max < 0 || l_max > 31) return fail(ODIS_ERR_ARG, "sh degree must be in 0..31");
    odis::sh_basis(n, pos_sph, l_max, (size_t)n, Y_out);
    return ODIS_OK;
}

int odis_sh_normal_inverse(int32_t n, const double* pos_sph, int32_t l_max, double* Ginv_out) {
    if (!pos_sph || !Ginv_out || n <= 0) return fail(ODIS_ERR_ARG, "NULL or empty argument");
    if (l_max < 0 || l_max > 31) return fail(ODIS_ERR_ARG, "sh degree must be in 0..31");
    const int rows = odis::sh_rows(l_max);
    std::vector<double> Y((size_t)rows * n), G;
    odis::sh_basis(n, pos_sph, l_max, (size_t)n, Y.data());
    if (odis::sh_normal_inverse(rows, n, (size_t)n, Y.data(), 0, G) != 0) return fail(ODIS_ERR_ARG, "normal matrix is not positive definite");
    std::memcpy(Ginv_out, G.data(), G.size() * sizeof(double));
    return ODIS_OK;
}

// ------------------------------------------------------------------ config ----
int odis_config_create(odis_config** out) {
    if (!out) return fail(ODIS_ERR_ARG, "out is NULL");
    *out = new (std::nothrow) odis_config();
    return *out ? ODIS_OK : fail(ODIS_ERR_ARG, "out of memory");
}

int odis_config_load(const char* run_dir, odis_config** out) {
    if (!run_dir || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    odis_config* c = new (std::nothrow) odis_config();
    if (!c) return fail(ODIS_ERR_ARG, "out of memory");
    std::string err;
    if (c->cfg.load(run_dir, err) != 0) {
        delete c;
        return fail(ODIS_ERR_IO, err);
    }
    if (c->cfg.finalize(err) != 0) {
        delete c;
        return fail(ODIS_ERR_CONFIG, err);
    }
    c->finalized = true;
    *out = c;
    return ODIS_OK;
}

int odis_config_set(odis_config* cfg, const char* key, const char* value_text) {
    if (!cfg || !key || !value_text) return fail(ODIS_ERR_ARG, "NULL argument");
    if (cfg->cfg.set(key, value_text) != 0) return fail(ODIS_ERR_ARG, std::string("unknown input.in key: ") + key);
    return ODIS_OK;
}

int odis_config_finalize(odis_config* cfg) {
    if (!cfg) return fail(ODIS_ERR_ARG, "NULL argument");
    std::string err;
    if (cfg->cfg.finalize(err) != 0) return fail(ODIS_ERR_CONFIG, err);
    cfg->finalized = true;
    return ODIS_OK;
}

static const odis::ConfigEntry* lookup(const odis_config* cfg, const char* key, odis::ConfigEntry::Type t, int* rc) {
    if (!cfg || !key) { *rc = fail(ODIS_ERR_ARG, "NULL argument"); return nullptr; }
    const odis::ConfigEntry* e = cfg->cfg.find(key);
    if (!e) { *rc = fail(ODIS_ERR_ARG, std::string("unknown input.in key: ") + key); return nullptr; }
    if (e->type != t) { *rc = fail(ODIS_ERR_ARG, std::string("wrong type requested for key: ") + key); return nullptr; }
    *rc = ODIS_OK;
    return e;
}

int odis_config_get_double(const odis_config* cfg, const char* key, double* out) {
    int rc;
    const odis::ConfigEntry* e = lookup(cfg, key, odis::ConfigEntry::DOUBLE, &rc);
    if (e && out) *out = e->d;
    return rc;
}
int odis_config_get_int(const odis_config* cfg, const char* key, int32_t* out) {
    int rc;
    const odis::ConfigEntry* e = lookup(cfg, key, odis::ConfigEntry::INT, &rc);
    if (e && out) *out = e->i;
    return rc;
}
int odis_config_get_bool(const odis_config* cfg, const char* key, int32_t* out) {
    int rc;
    const odis::ConfigEntry* e = lookup(cfg, key, odis::ConfigEntry::BOOL, &rc);
    if (e && out) *out = e->b ? 1 : 0;
    return rc;
}
int odis_config_get_string(const odis_config* cfg, const char* key, char* buf, int32_t buflen) {
    int rc;
    const odis::ConfigEntry* e = lookup(cfg, key, odis::ConfigEntry::STRING, &rc);
    if (e && buf && buflen > 0) {
        std::strncpy(buf, e->s.c_str(), (size_t)buflen - 1);
        buf[buflen - 1] = 0;
    }
    return rc;
}
int odis_config_get_enum(const odis_config* cfg, int32_t which, int32_t* out) {
    if (!cfg || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!cfg->finalized) return fail(ODIS_ERR_STATE, "odis_config_finalize has not been called");
    switch (which) {
        case 0: *out = cfg->cfg.fric_type; break;
        case 1: *out = cfg->cfg.surface_type; break;
        case 2: *out = cfg->cfg.solver_type; break;
        case 3: *out = cfg->cfg.tide_type; break;
        case 4: *out = cfg->cfg.initial_condition; break;
        default: return fail(ODIS_ERR_ARG, "which must be 0..4");
    }
    return ODIS_OK;
}
void odis_config_free(odis_config* cfg) { delete cfg; }

int odis_analytical_state(const odis_mesh_view* mv, const odis_params* prm, double* v, double* dvdt, double* eta, double* detadt) {
    if (!mv || !prm || !v || !dvdt || !eta || !detadt) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!mv->node_pos_sph || !mv->face_centre_pos_sph || !mv->face_normal_vec_map) return fail(ODIS_ERR_ARG, "mesh view lacks the position / normal tables");
    if (prm->potential != 1 /*OBLIQ_WEST*/)
        return fail(ODIS_ERR_UNSUPPORTED, "the analytical solution exists for potential OBLIQ_WEST only (analyticalLTE.cpp:133; other types return nothing there)");
    if (!(prm->omega != 0.0) || !(prm->g * prm->h > 0.0)) return fail(ODIS_ERR_ARG, "omega, g and h must be non-zero");
    const odis::AnalyticParams ap{prm->radius, prm->omega, prm->g, prm->h, prm->alpha, prm->obl, prm->dt};
    odis::analytical_state_obliq_west(ap, mv->n_cells, mv->n_edges, mv->node_pos_sph, mv->face_centre_pos_sph, mv->face_normal_vec_map, v, dvdt, eta,
                                      detadt);
    return ODIS_OK;
}

int odis_quantise_time_step(double period, double target_dt, double* dt_out, int32_t* steps_out) {
    if (!dt_out || !steps_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int n = 0;
    odis::quantise_time_step(period, target_dt, dt_out, &n);
    *steps_out = n;
    return ODIS_OK;
}

// ------------------------------------------------------------------ mesh ----
static int build_from(const odis::GridFile& g, double radius, int threads, odis_mesh** out) {
    odis_mesh* m = new (std::nothrow) odis_mesh();
    if (!m) return fail(ODIS_ERR_ARG, "out of memory");
    std::string err;
    if (odis::build_mesh_tables(g, radius, m->t, err, threads) != 0) {
        delete m;
        return fail(ODIS_ERR_GRID, err);
    }
    *out = m;
    return ODIS_OK;
}

int odis_mesh_from_file(const char* grid_path, double radius, int32_t threads, odis_mesh** out) {
    if (!grid_path || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    odis::GridFile g;
    std::string err;
    if (odis::read_grid_file(grid_path, g, err) != 0) return fail(ODIS_ERR_IO, err);
    return build_from(g, radius, threads, out);
}

int odis_mesh_from_arrays(int32_t n_cells, const double* node_pos_sph, const int32_t* node_friends,
                          const double* centroid_pos_sph, double radius, int32_t threads, odis_mesh** out) {
    if (!node_pos_sph || !node_friends || !centroid_pos_sph || !out || n_cells < 12) return fail(ODIS_ERR_ARG, "bad argument");
    odis::GridFile g;
    g.n_cells = n_cells;
    g.node_pos_sph.assign(node_pos_sph, node_pos_sph + (size_t)n_cells * 2);
    g.node_friends.assign(node_friends, node_friends + (size_t)n_cells * 6);
    g.centroid_pos_sph.assign(centroid_pos_sph, centroid_pos_sph + (size_t)n_cells * 12);
    return build_from(g, radius, threads, out);
}

int odis_mesh_get_view(const odis_mesh* mesh, odis_mesh_view* v) {
    if (!mesh || !v) return fail(ODIS_ERR_ARG, "NULL argument");
    const odis::MeshTables& t = mesh->t;
    v->n_cells = t.n_cells; v->n_edges = t.n_edges; v->n_vertices = t.n_vertices; v->radius = t.radius;
    v->node_pos_sph = t.node_pos_sph.data();
    v->node_friends = t.node_friends.data();
    v->centroid_pos_sph = t.centroid_pos_sph.data();
    v->control_volume_surf_area_map = t.control_volume_surf_area_map.data();
    v->faces = t.faces.data();
    v->node_face_dir = t.node_face_dir.data();
    v->vertexes = t.vertexes.data();
    v->face_nodes = t.face_nodes.data();
    v->face_vertexes = t.face_vertexes.data();
    v->face_interp_friends = t.face_interp_friends.data();
    v->face_interp_weights = t.face_interp_weights.data();
    v->face_len = t.face_len.data();
    v->face_node_dist = t.face_node_dist.data();
    v->face_centre_m = t.face_centre_m.data();
    v->face_centre_pos_sph = t.face_centre_pos_sph.data();
    v->face_intercept_pos_sph = t.face_intercept_pos_sph.data();
    v->face_area = t.face_area.data();
    v->face_normal_vec_map = t.face_normal_vec_map.data();
    v->vertex_pos_sph = t.vertex_pos_sph.data();
    v->vertex_nodes = t.vertex_nodes.data();
    v->vertex_R = t.vertex_R.data();
    return ODIS_OK;
}
void odis_mesh_free(odis_mesh* mesh) { delete mesh; }

// ------------------------------------------------------------------ grid ----
int odis_grid_generate(int32_t level, int32_t* n_cells_out, double** node_pos_sph_out, int32_t** node_friends_out,
                       double** centroid_pos_sph_out) {
    if (!n_cells_out || !node_pos_sph_out || !node_friends_out || !centroid_pos_sph_out) return fail(ODIS_ERR_ARG, "NULL argument");
    odis::GridFile g;
    std::string err;
    if (odis::generate_icosahedral_grid(level, g, err) != 0) return fail(ODIS_ERR_ARG, err);
    const size_t n = (size_t)g.n_cells;
    double* pos = (double*)std::malloc(n * 2 * sizeof(double));
    int32_t* fr = (int32_t*)std::malloc(n * 6 * sizeof(int32_t));
    double* cen = (double*)std::malloc(n * 12 * sizeof(double));
    if (!pos || !fr || !cen) {
        std::free(pos); std::free(fr); std::free(cen);
        return fail(ODIS_ERR_ARG, "out of memory");
    }
    std::memcpy(pos, g.node_pos_sph.data(), n * 2 * sizeof(double));
    std::memcpy(fr, g.node_friends.data(), n * 6 * sizeof(int32_t));
    std::memcpy(cen, g.centroid_pos_sph.data(), n * 12 * sizeof(double));
    *n_cells_out = g.n_cells;
    *node_pos_sph_out = pos; *node_friends_out = fr; *centroid_pos_sph_out = cen;
    return ODIS_OK;
}

int odis_grid_write_file(const char* path, int32_t n_cells, const double* node_pos_sph, const int32_t* node_friends,
                         const double* centroid_pos_sph) {
    if (!path || !node_pos_sph || !node_friends || !centroid_pos_sph || n_cells < 12) return fail(ODIS_ERR_ARG, "bad argument");
    odis::GridFile g;
    g.n_cells = n_cells;
    g.node_pos_sph.assign(node_pos_sph, node_pos_sph + (size_t)n_cells * 2);
    g.node_friends.assign(node_friends, node_friends + (size_t)n_cells * 6);
    g.centroid_pos_sph.assign(centroid_pos_sph, centroid_pos_sph + (size_t)n_cells * 12);
    std::string err;
    if (odis::write_grid_file(path, g, err) != 0) return fail(ODIS_ERR_IO, err);
    return ODIS_OK;
}

void odis_free(void* p) { std::free(p); }

// ------------------------------------------------------------------ partition plan (host only) ----
static int32_t* dup_ints(const std::vector<int>& v) {
    int32_t* p = (int32_t*)std::malloc((v.size() + 1) * sizeof(int32_t));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(int32_t));
    return p;
}

int odis_partition_plan(const odis_mesh_view* mv, int32_t reorder, int32_t rank, int32_t world, odis_partition_plan_t* plan) {
    if (!mv || !plan) return fail(ODIS_ERR_ARG, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ODIS_ERR_ARG, "rank/world out of range");
    odis::LocalNumbering L;
    odis::build_local_numbering(mv->n_cells, mv->n_edges, mv->node_pos_sph, mv->face_nodes, mv->faces, reorder != 0, rank, world, L);
    std::memset(plan, 0, sizeof *plan);
    plan->rank = rank; plan->world = world;
    plan->own_cells = L.part.n_own_cells; plan->own_edges = L.part.n_own_edges;
    plan->local_cells = (int32_t)L.cell_perm.size(); plan->local_edges = (int32_t)L.edge_perm.size();
    plan->local_cell_ref = dup_ints(L.cell_perm);
    plan->local_edge_ref = dup_ints(L.edge_perm);
    plan->n_peers = (int32_t)L.part.peers.size();
    std::vector<int> ranks, counts, se_ref, se_slot, sc_ref, sc_slot;
    for (const odis::HaloPeer& p : L.part.peers) {
        ranks.push_back(p.rank);
        counts.push_back((int)p.send_edge_local.size()); counts.push_back((int)p.send_cell_local.size());
        counts.push_back(p.recv_edges); counts.push_back(p.recv_cells);
        for (size_t i = 0; i < p.send_edge_local.size(); i++) { se_ref.push_back(L.edge_perm[(size_t)p.send_edge_local[i]]); se_slot.push_back(p.send_edge_remote[i]); }
        for (size_t i = 0; i < p.send_cell_local.size(); i++) { sc_ref.push_back(L.cell_perm[(size_t)p.send_cell_local[i]]); sc_slot.push_back(p.send_cell_remote[i]); }
    }
    plan->peer_rank = dup_ints(ranks);
    plan->peer_counts = dup_ints(counts);
    plan->send_edge_ref = dup_ints(se_ref); plan->send_edge_slot = dup_ints(se_slot);
    plan->send_cell_ref = dup_ints(sc_ref); plan->send_cell_slot = dup_ints(sc_slot);
    return ODIS_OK;
}

void odis_partition_plan_free(odis_partition_plan_t* plan) {
    if (!plan) return;
    std::free(plan->local_cell_ref); std::free(plan->local_edge_ref); std::free(plan->peer_rank); std::free(plan->peer_counts);
    std::free(plan->send_edge_ref); std::free(plan->send_edge_slot); std::free(plan->send_cell_ref); std::free(plan->send_cell_slot);
    std::memset(plan, 0, sizeof *plan);
}

}  // extern "C"
