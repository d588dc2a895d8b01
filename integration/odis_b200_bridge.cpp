// See odis_b200_bridge.h. Compiled inside a GeodesicODIS build (reference headers on the include path).
#include "odis_b200_bridge.h"

#include "gridConstants.h"
#include "outFiles.h"

#include <cstdlib>
#include <sstream>
#include <vector>

namespace odis_bridge {

namespace {
odis_solver* g_solver = nullptr;
Globals* g_globals = nullptr;
Mesh* g_grid = nullptr;
Group g_group;                       // world > 1: the partitioned solvers (g_solver stays null)
std::vector<double> g_part;          // one rank's share of a field while the whole is summed
}  // namespace

void check(Globals* globals, int rc, const char* what) {
    if (rc == ODIS_OK) return;
    std::ostringstream msg;
    msg << "ERROR: " << what << " failed in libodis_b200 (" << rc << "): " << odis_last_error() << std::endl;
    globals->Output->Write(ERR_MESSAGE, &msg);
    globals->Output->TerminateODIS();
}

odis_mesh_view mesh_view(Globals* globals, Mesh* grid) {
    odis_mesh_view mv{};
    mv.n_cells = NODE_NUM;
    mv.n_edges = FACE_NUM;
    mv.n_vertices = VERTEX_NUM;
    mv.radius = globals->radius.Value();
    mv.node_pos_sph = &grid->node_pos_sph(0, 0);
    mv.node_friends = &grid->node_friends(0, 0);
    mv.centroid_pos_sph = &grid->centroid_pos_sph(0, 0, 0);
    mv.control_volume_surf_area_map = &grid->control_volume_surf_area_map(0);
    mv.faces = &grid->faces(0, 0);
    mv.node_face_dir = &grid->node_face_dir(0, 0);
    mv.vertexes = &grid->vertexes(0, 0);
    mv.face_nodes = &grid->face_nodes(0, 0);
    mv.face_vertexes = &grid->face_vertexes(0, 0);
    mv.face_interp_friends = &grid->face_interp_friends(0, 0);
    mv.face_interp_weights = &grid->face_interp_weights(0, 0);
    mv.face_len = &grid->face_len(0);
    mv.face_node_dist = &grid->face_node_dist(0);
    mv.face_centre_m = &grid->face_centre_m(0, 0);
    mv.face_centre_pos_sph = &grid->face_centre_pos_sph(0, 0);
    mv.face_intercept_pos_sph = &grid->face_intercept_pos_sph(0, 0);
    mv.face_area = &grid->face_area(0);
    mv.face_normal_vec_map = &grid->face_normal_vec_map(0, 0);
    mv.vertex_pos_sph = &grid->vertex_pos_sph(0, 0);
    mv.vertex_nodes = &grid->vertex_nodes(0, 0);
    mv.vertex_R = &grid->vertex_R(0, 0);
    return mv;
}

odis_params params(Globals* globals) {
    odis_params p{};
    p.g = globals->g.Value();
    p.h = globals->h.Value();
    p.alpha = globals->alpha.Value();
    p.dt = globals->timeStep.Value();
    p.radius = globals->radius.Value();
    p.omega = globals->angVel.Value();
    p.love_reduct = globals->loveReduct.Value();
    p.ecc = globals->e.Value();
    p.obl = globals->theta.Value();
    p.shell_thickness = globals->shell_thickness.Value();
    p.semimajor_axis = globals->a.Value();
    p.potential = (int32_t)globals->tide_type;
    p.friction = (int32_t)globals->fric_type;
    p.surface = (int32_t)globals->surface_type;
    p.init_load = globals->initial_condition == INIT_LOAD ? 1 : 0;      // temporalOperators.cpp:36
    p.reorder = 1;
    return p;
}

odis_solver* solver(Globals* globals, Mesh* grid) {
    if (g_solver && g_globals == globals && (grid == nullptr || g_grid == grid)) return g_solver;
    if (grid == nullptr) check(globals, ODIS_ERR_STATE, "looking up the device solver (no Mesh seen yet)");
    release();
    const odis_mesh_view mv = mesh_view(globals, grid);
    const odis_params p = params(globals);
    odis_solver* s = nullptr;
    check(globals, odis_create(&mv, &p, /*device*/ 0, &s), "odis_create");
    if (globals->advection.Value()) {
        // the three operators only the nonlinear branch reads, handed over as Eigen stores them (row-major compressed)
        auto csr = [](const SpMat& A) {
            return odis_csr_view{(int32_t)A.rows(), (int32_t)A.cols(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr()};
        };
        odis_nonlinear_view nv{csr(grid->operatorCurl), csr(grid->operatorRBFinterp), csr(grid->operatorDirectionalSecondDeriv),
                               &grid->vertex_sinlat(0), &grid->vertex_area(0)};
        check(globals, odis_enable_advection(s, &mv, &nv), "odis_enable_advection");
    }
    g_solver = s;
    g_globals = globals;
    g_grid = grid;
    return s;
}

void release() {
    if (g_solver) odis_destroy(g_solver);
    for (int r = 0; r < g_group.world; r++)
        if (g_group.rank[r] && g_group.rank[r] != g_solver) odis_destroy(g_group.rank[r]);
    g_group = Group{};
    g_solver = nullptr;
    g_globals = nullptr;
    g_grid = nullptr;
}

const Group& group(Globals* globals, Mesh* grid) {
    if (g_group.rank[0] && g_globals == globals && (grid == nullptr || g_grid == grid)) return g_group;
    const char* e = std::getenv("ODIS_B200_GPUS");
    const int world = e ? std::atoi(e) : 1;
    if (world <= 1) {                                   // one GPU: the solver of solver()
        odis_solver* s = solver(globals, grid);
        g_group = Group{};
        g_group.rank[0] = s;
        return g_group;
    }
    if (world > 8) check(globals, ODIS_ERR_ARG, "ODIS_B200_GPUS: at most 8 GPUs");
    if (globals->advection.Value()) check(globals, ODIS_ERR_UNSUPPORTED, "ODIS_B200_GPUS > 1 with `advection; true` (the nonlinear branch runs on one GPU)");
    if (grid == nullptr) check(globals, ODIS_ERR_STATE, "looking up the device solvers (no Mesh seen yet)");
    release();
    const odis_mesh_view mv = mesh_view(globals, grid);
    const odis_params p = params(globals);
    Group g;
    g.world = world;
    for (int r = 0; r < world; r++) check(globals, odis_create_partitioned(&mv, &p, /*device*/ r, r, world, &g.rank[r]), "odis_create_partitioned");
    const size_t bs = (size_t)odis_halo_blob_size();
    std::vector<unsigned char> blobs(bs * (size_t)world);
    for (int r = 0; r < world; r++) check(globals, odis_halo_export(g.rank[r], blobs.data() + bs * (size_t)r), "odis_halo_export");
    for (int r = 0; r < world; r++) check(globals, odis_halo_connect(g.rank[r], blobs.data()), "odis_halo_connect");
    g_group = g;
    g_globals = globals;
    g_grid = grid;
    return g_group;
}

void set_state_all(Globals* globals, const Group& g, const double* v, const double* eta, const double* dvdt, const double* detadt, int64_t iter) {
    for (int r = 0; r < g.world; r++) check(globals, odis_set_state(g.rank[r], v, eta, dvdt, detadt, iter), "odis_set_state");
}

void step_all(Globals* globals, const Group& g, int32_t nsteps) {
    // every rank's steps are enqueued before any rank is waited for: the steps in flight wait inside the kernels for their neighbours
    for (int r = 0; r < g.world; r++) check(globals, odis_step(g.rank[r], nsteps), "odis_step");
    for (int r = 0; r < g.world; r++) check(globals, odis_synchronize(g.rank[r]), "odis_synchronize");
}

void get_field_all(Globals* globals, const Group& g, int32_t field, double* out, size_t count) {
    check(globals, odis_get_field(g.rank[0], field, out), "odis_get_field");
    if (g.world == 1) return;
    g_part.resize(count);
    for (int r = 1; r < g.world; r++) {
        check(globals, odis_get_field(g.rank[r], field, g_part.data()), "odis_get_field");
        for (size_t i = 0; i < count; i++) out[i] += g_part[i];       // own entries of rank r, zeros elsewhere
    }
}

double dissipation_avg_all(Globals* globals, const Group& g) {
    double tot = 0.0;
    for (int r = 0; r < g.world; r++) {
        double x = 0.0;
        check(globals, odis_get_dissipation_avg(g.rank[r], &x), "odis_get_dissipation_avg");
        tot += x;
    }
    return tot;
}

}  // namespace odis_bridge
