// Whole-run driver: what `./ODIS` does from a run directory, on the GPU engine.
//
// Replaces main() (/root/reference/src/main.cpp:46-68), solveODIS (src/solver.cpp:15-61), the
// non-loop parts of ab3Explicit (src/timeIntegrator.cpp:57-204,276-322: output cadence, the
// "DUMPING DATA" log line, SIGINT handling, restart files) and OutFiles (src/outFiles.cpp: OUTPUT.txt /
// ERROR.txt, data.h5 creation :138-462, grid dump :464-520, DumpData :522-684). The loop body itself
// is odis_step(). Same files in, same files out:
//   <run_dir>/input.in, <run_dir>/input_files/grid_l<L>.txt [, InitialConditions/*.txt]
//   <run_dir>/DATA/OUTPUT.txt, ERROR.txt, data.h5 ; <run_dir>/InitialConditions/{vel,pres}_init.txt
#include <cmath>
#include <csignal>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/odis_b200.h"
#include "odis_config.h"
#include "odis_error.h"
#include "odis_h5lite.h"
#include "odis_mesh.h"
#include "odis_mesh_nl.h"
#include "odis_sphere.h"

using odis::fail;

struct odis_h5 {
    odis::H5LiteWriter w;
};

namespace {

volatile std::sig_atomic_t g_sigint = 0;
void on_sigint(int) { g_sigint = 1; }     // CatchExit, timeIntegrator.cpp:34-37

struct Logs {
    std::string out_path, err_path;
    bool echo = false;
    void out(const std::string& s) const {
        std::ofstream f(out_path, std::ofstream::out | std::ofstream::app);
        f << s << std::endl;              // OutFiles::WriteMessage appends a newline (outFiles.cpp:72)
        if (echo) std::printf("%s\n", s.c_str());
    }
    void err(const std::string& s) const {
        std::ofstream f(err_path, std::ofstream::out | std::ofstream::app);
        f << s << std::endl;
    }
};

std::string fmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    std::vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

// Globals::OutputConsts, globals.cpp:573-597 (default ostream formatting = %g with 6 digits)
void write_model_parameters(const odis::Config& cfg, const Logs& log) {
    log.out("Model parameters: ");
    for (const std::string& key : cfg.keys()) {
        const odis::ConfigEntry* e = cfg.find(key);
        std::ostringstream o;
        o << "\t\t " << std::left;
        o.width(30);
        o << key;
        switch (e->type) {
            case odis::ConfigEntry::DOUBLE: o << e->d; break;
            case odis::ConfigEntry::INT: o << e->i; break;
            case odis::ConfigEntry::BOOL: o << e->b; break;
            case odis::ConfigEntry::STRING: o << e->s; break;
        }
        log.out(o.str());
    }
}

// loadInitialConditions, initialConditions.cpp:19-144: "<a>, <b>, <c>, <d>" per line, std::stod on each token
int load_restart(const std::string& path, size_t n, std::vector<double>& s, std::vector<double>& hist) {
    std::ifstream f(path);
    if (!f.is_open()) return -1;
    std::string line;
    size_t i = 0;
    while (i < n && std::getline(f, line)) {
        std::istringstream ls(line);
        std::string tok;
        double vals[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; k++) {
            if (!std::getline(ls >> std::ws, tok, ' ')) break;
            try { vals[k] = std::stod(tok); } catch (...) { return -2; }
        }
        s[i] = vals[0];
        hist[i * 3] = vals[1]; hist[i * 3 + 1] = vals[2]; hist[i * 3 + 2] = vals[3];
        i++;
    }
    return 0;
}

// writeInitialConditions, initialConditions.cpp:209-276 ("%1.6E, %1.6E, %1.6E, %1.6E\n")
void write_restart(const std::string& path, size_t n, const std::vector<double>& s, const std::vector<double>& hist) {
    std::remove(path.c_str());
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return;
    for (size_t i = 0; i < n; i++) std::fprintf(f, "%1.6E, %1.6E, %1.6E, %1.6E\n", s[i], hist[i * 3], hist[i * 3 + 1], hist[i * 3 + 2]);
    std::fclose(f);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ HDF5 output ----
int odis_h5_create(const char* path, odis_h5** out) {
    if (!path || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    odis_h5* h = new odis_h5();
    std::string err;
    if (h->w.create(path, err) != 0) { delete h; return fail(ODIS_ERR_IO, err); }
    *out = h;
    return ODIS_OK;
}
int odis_h5_add_dataset(odis_h5* h, const char* name, int32_t rank, const uint64_t* dims, int32_t* id_out) {
    if (!h || !name || !dims || !id_out) return fail(ODIS_ERR_ARG, "NULL argument");
    std::string err;
    const int id = h->w.add_dataset(name, rank, dims, err);
    if (id < 0) return fail(ODIS_ERR_ARG, err);
    *id_out = id;
    return ODIS_OK;
}
int odis_h5_write_rows(odis_h5* h, int32_t dataset, uint64_t first_row, uint64_t nrows, const float* data) {
    if (!h || !data) return fail(ODIS_ERR_ARG, "NULL argument");
    std::string err;
    if (h->w.write_rows(dataset, first_row, nrows, data, err) != 0) return fail(ODIS_ERR_IO, err);
    return ODIS_OK;
}
int odis_h5_close(odis_h5* h) {
    if (!h) return ODIS_OK;
    std::string err;
    const int rc = h->w.close(err);
    delete h;
    return rc ? fail(ODIS_ERR_IO, err) : ODIS_OK;
}

// ------------------------------------------------------------------ whole run ----
int odis_run(const char* run_dir_c, const odis_run_options* opt_in, odis_run_result* res) {
    if (!run_dir_c) return fail(ODIS_ERR_ARG, "NULL run_dir");
    odis_run_options opt{};
    opt.reorder = 1;
    if (opt_in) opt = *opt_in;
    odis_run_result local{};
    if (!res) res = &local;
    *res = odis_run_result{};
    const std::string dir(run_dir_c);
    const std::string data = dir + "/DATA";
    ::mkdir(data.c_str(), 0770);                                              // outFiles.cpp:160
    Logs log{data + "/OUTPUT.txt", data + "/ERROR.txt", opt.echo != 0};
    std::remove(log.out_path.c_str());
    std::remove(log.err_path.c_str());
    log.out("\n\n        | ODIS time-step engine for NVIDIA B200 (GeodesicODIS-compatible run directory) |\n\n");
    log.err("ODIS error file. Warnings and model termination errors will be written here.\n");   // outFiles.cpp:64
    auto terminate = [&](int code, const std::string& msg) {                  // OutFiles::TerminateODIS, outFiles.cpp:123-130
        log.err(msg);
        log.err("TERMINATING ODIS.");
        return fail(code, msg);
    };
    // a failed write (full disk, quota) must end the run with an error instead of leaving a silently truncated data.h5
    int io_rc = ODIS_OK;
    std::string io_err, err;
    auto wr = [&](int r) { if (r != 0 && io_rc == ODIS_OK) { io_rc = ODIS_ERR_IO; io_err = err; } };

    // ---- Globals(0) ----
    odis::Config cfg;
    if (cfg.load(dir, err) != 0) return terminate(ODIS_ERR_IO, err);
    log.out("Found input file.");                                             // globals.cpp:351
    for (const std::string& k : cfg.unassigned())                             // globals.cpp:445-466
        log.err("WARNING: Unassigned global constant: " + k + ". \nAssigning default value could be risky...\nUsing Titan value\n");
    if (cfg.finalize(err) != 0) return terminate(ODIS_ERR_CONFIG, err);
    if (cfg.solver_type != odis::AB3) return terminate(ODIS_ERR_UNSUPPORTED, "only solver type AB3 has a live implementation (solver.cpp:25-50)");
    if (cfg.initial_condition == odis::INIT_ANALYTICAL && cfg.tide_type != 1 /*OBLIQ_WEST*/)
        return terminate(ODIS_ERR_UNSUPPORTED, "initial conditions; ANALYTICAL exists for potential OBLIQ_WEST only (analyticalLTE.cpp:133)");
    for (const char* k : {"pressure output", "kinetic output", "dummy2 output"})
        if (cfg.get_bool(k)) log.err(std::string("WARNING: '") + k + "' would create a second dataset named 'displacement' (outFiles.cpp:286,308,338); ignored.");

    // ---- Mesh ----
    const std::string grid_path = dir + "/input_files/grid_l" + std::to_string(cfg.get_int("geodesic grid level")) + ".txt";   // mesh.cpp:4023
    odis::GridFile grid;
    if (odis::read_grid_file(grid_path, grid, err) != 0) return terminate(ODIS_ERR_IO, err);
    log.out("\nFound mesh file: " + grid_path);                               // mesh.cpp:4030
    odis::MeshTables mesh;
    if (odis::build_mesh_tables(grid, cfg.get_double("radius"), mesh, err, 0) != 0) return terminate(ODIS_ERR_GRID, err);
    double dt = 0.0;
    int total_iter = 0;
    odis::quantise_time_step(cfg.get_double("orbital period"), cfg.get_double("time step"), &dt, &total_iter);   // mesh.cpp:1601-1618
    cfg.set_double("time step", dt);
    cfg.set_int("total iterations", total_iter);
    const int N = mesh.n_cells, F = mesh.n_edges;
    const double end_time = cfg.get_double("simulation end time");
    const int output_time = cfg.get_int("output time");
    if (total_iter <= 0 || output_time <= 0 || total_iter / output_time <= 0)
        return terminate(ODIS_ERR_CONFIG, "time step / output time give no usable output cadence");
    const int out_freq = total_iter / output_time;                            // timeIntegrator.cpp:188

    // ---- data.h5 (CreateHDF5Framework, outFiles.cpp:138-462) ----
    odis::H5LiteWriter h5;
    if (h5.create(data + "/data.h5", err) != 0) return terminate(ODIS_ERR_IO, err);
    const uint64_t T = (uint64_t)((int)end_time * output_time + 1);            // outFiles.cpp:149-150,179
    const uint64_t dims_f[2] = {T, (uint64_t)F}, dims_n[2] = {T, (uint64_t)N}, dims_t[1] = {T}, dims_fp[1] = {(uint64_t)F};
    int ds_u = -1, ds_v = -1, ds_eta = -1, ds_diss = -1, ds_avg = -1, ds_kin = -1, ds_d1 = -1;
    int ds_ux = -1, ds_uy = -1, ds_uz = -1;
    if (cfg.get_bool("velocity output")) { ds_u = h5.add_dataset("east velocity", 2, dims_f, err); ds_v = h5.add_dataset("north velocity", 2, dims_f, err); }
    if (cfg.get_bool("velocity cartesian output")) {                           // outFiles.cpp:255-272
        ds_ux = h5.add_dataset("x velocity", 2, dims_n, err); ds_uy = h5.add_dataset("y velocity", 2, dims_n, err);
        ds_uz = h5.add_dataset("z velocity", 2, dims_n, err);
    }
    if (cfg.get_bool("displacement output")) ds_eta = h5.add_dataset("displacement", 2, dims_n, err);
    if (cfg.get_bool("dissipation output")) ds_diss = h5.add_dataset("dissipated energy", 2, dims_f, err);
    if (cfg.get_bool("dissipation avg output")) ds_avg = h5.add_dataset("dissipation avg output", 1, dims_t, err);
    if (cfg.get_bool("kinetic avg output")) ds_kin = h5.add_dataset("kinetic avg output", 1, dims_t, err);
    if (cfg.get_bool("dummy1 output")) ds_d1 = h5.add_dataset("dummy1 output", 2, dims_n, err);
    const int ds_lon = h5.add_dataset("face longitude", 1, dims_fp, err), ds_lat = h5.add_dataset("face latitude", 1, dims_fp, err);
    if (h5.finalize(err) != 0) return terminate(ODIS_ERR_IO, err);
    {   // DumpGridData, outFiles.cpp:492-515
        std::vector<float> lons((size_t)F), lats((size_t)F);
        for (int i = 0; i < F; i++) {
            lats[i] = (float)((float)mesh.face_centre_pos_sph[(size_t)i * 2] * 180. / odis::kPi);
            lons[i] = (float)((float)mesh.face_centre_pos_sph[(size_t)i * 2 + 1] * 180. / odis::kPi);
        }
        wr(h5.write_rows(ds_lon, 0, (uint64_t)F, lons.data(), err));
        wr(h5.write_rows(ds_lat, 0, (uint64_t)F, lats.data(), err));
    }
    write_model_parameters(cfg, log);                                         // main.cpp:59

    // ---- solveODIS -> ab3Explicit ----
    log.out("Identifying using selected solver method... " + cfg.get_string("solver type") + "\n");
    log.out("Entering time solver " + cfg.get_string("solver type") + "...\n");
    odis_mesh_view mv{};
    mv.n_cells = N; mv.n_edges = F; mv.n_vertices = mesh.n_vertices; mv.radius = mesh.radius;
    mv.node_pos_sph = mesh.node_pos_sph.data(); mv.node_friends = mesh.node_friends.data();
    mv.centroid_pos_sph = mesh.centroid_pos_sph.data();
    mv.control_volume_surf_area_map = mesh.control_volume_surf_area_map.data();
    mv.faces = mesh.faces.data(); mv.node_face_dir = mesh.node_face_dir.data(); mv.vertexes = mesh.vertexes.data();
    mv.face_nodes = mesh.face_nodes.data(); mv.face_vertexes = mesh.face_vertexes.data();
    mv.face_interp_friends = mesh.face_interp_friends.data(); mv.face_interp_weights = mesh.face_interp_weights.data();
    mv.face_len = mesh.face_len.data(); mv.face_node_dist = mesh.face_node_dist.data(); mv.face_centre_m = mesh.face_centre_m.data();
    mv.face_centre_pos_sph = mesh.face_centre_pos_sph.data(); mv.face_intercept_pos_sph = mesh.face_intercept_pos_sph.data();
    mv.face_area = mesh.face_area.data(); mv.face_normal_vec_map = mesh.face_normal_vec_map.data();
    mv.vertex_pos_sph = mesh.vertex_pos_sph.data(); mv.vertex_nodes = mesh.vertex_nodes.data(); mv.vertex_R = mesh.vertex_R.data();
    odis_params p{};
    p.g = cfg.get_double("surface gravity"); p.h = cfg.get_double("ocean thickness"); p.alpha = cfg.get_double("friction coefficient");
    p.dt = dt; p.radius = cfg.get_double("radius"); p.omega = cfg.get_double("angular velocity");
    p.love_reduct = cfg.get_double("love reduction factor"); p.ecc = cfg.get_double("eccentricity"); p.obl = cfg.get_double("obliquity");
    p.shell_thickness = cfg.get_double("shell thickness"); p.semimajor_axis = cfg.get_double("semimajor axis");
    p.potential = cfg.tide_type; p.friction = cfg.fric_type; p.surface = cfg.surface_type;
    p.init_load = cfg.initial_condition == odis::INIT_LOAD; p.reorder = opt.reorder;
    // n_gpus > 1: the grid is cut into that many space-filling-curve parts, one solver per GPU (devices device .. device + n_gpus - 1) in
    // this one process, halos exchanged by the step kernels through peer memory (main.cpp:46-64 knows one process; so does this)
    const int world = opt.n_gpus > 1 ? opt.n_gpus : 1;
    std::vector<odis_solver*> ranks((size_t)world, nullptr);
    struct SolverGuard {                      // every return below releases the device memory and the streams
        std::vector<odis_solver*>& h;
        ~SolverGuard() { for (auto& x : h) if (x) { odis_destroy(x); x = nullptr; } }
    } solver_guard{ranks};
    int rc = ODIS_OK;
    for (int r = 0; r < world && rc == ODIS_OK; r++)
        rc = world == 1 ? odis_create(&mv, &p, opt.device, &ranks[(size_t)r])
                        : odis_create_partitioned(&mv, &p, opt.device + r, r, world, &ranks[(size_t)r]);
    if (rc != ODIS_OK) return terminate(rc, odis_last_error());
    if (world > 1) {
        const size_t bs = (size_t)odis_halo_blob_size();
        std::vector<unsigned char> blobs(bs * (size_t)world);
        for (int r = 0; r < world && rc == ODIS_OK; r++) rc = odis_halo_export(ranks[(size_t)r], blobs.data() + bs * (size_t)r);
        for (int r = 0; r < world && rc == ODIS_OK; r++) rc = odis_halo_connect(ranks[(size_t)r], blobs.data());
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
        log.out("grid partitioned over " + std::to_string(world) + " GPUs");
    }
    odis_solver* const s = ranks[0];          // unpartitioned calls below (advection, output snapshots) go to the one solver
    auto on_all = [&](auto&& call) {          // the same call on every rank, in rank order; first failure wins
        int r2 = ODIS_OK;
        for (int r = 0; r < world && r2 == ODIS_OK; r++) r2 = call(ranks[(size_t)r]);
        return r2;
    };
    // a field of the whole grid: every rank fills its own entries and leaves zeros elsewhere, so the sum is the field
    std::vector<double> part_buf;
    auto get_field_all = [&](int32_t field, std::vector<double>& out) {
        int r2 = odis_get_field(ranks[0], field, out.data());
        for (int r = 1; r < world && r2 == ODIS_OK; r++) {
            part_buf.resize(out.size());
            r2 = odis_get_field(ranks[(size_t)r], field, part_buf.data());
            if (r2 == ODIS_OK)
                for (size_t i = 0; i < out.size(); i++) out[i] += part_buf[i];
        }
        return r2;
    };
    auto dissipation_all = [&](double* out) {         // sum of the ranks' shares of sum(eps_e A_e) / (4 pi r^2)
        double tot = 0.0;
        int r2 = ODIS_OK;
        for (int r = 0; r < world && r2 == ODIS_OK; r++) {
            double x = 0.0;
            r2 = odis_get_dissipation_avg(ranks[(size_t)r], &x);
            tot += x;
        }
        *out = tot;
        return r2;
    };
    // tables of the nonlinear branch: the step needs them with `advection; true` (updateMomentum.cpp:37, updateEta.cpp:32), the
    // Cartesian velocity output needs operatorRBFinterp either way (timeIntegrator.cpp:173,290)
    odis::NonlinearTables nlt;
    if (cfg.get_bool("advection") || ds_ux >= 0) {
        if (odis::build_nonlinear_tables(mesh, mesh.radius, cfg.get_double("rbf epsilon"), nlt, err) != 0) return terminate(ODIS_ERR_GRID, err);
    }
    if (cfg.get_bool("advection") && world > 1)
        return terminate(ODIS_ERR_UNSUPPORTED, "advection; true runs on one GPU (the nonlinear branch is not partitioned)");
    if (cfg.get_bool("advection")) {
        auto csr = [](const odis::Csr& A) {
            odis_csr_view v;
            v.n_rows = A.n_rows; v.n_cols = A.n_cols; v.indptr = A.indptr.data(); v.indices = A.indices.data(); v.data = A.data.data();
            return v;
        };
        odis_nonlinear_view nv{};
        nv.curl = csr(nlt.curl); nv.rbf_interp = csr(nlt.rbf_interp); nv.directional_second_deriv = csr(nlt.directional_second_deriv);
        nv.vertex_sinlat = nlt.vertex_sinlat.data(); nv.vertex_area = nlt.vertex_area.data();
        rc = odis_enable_advection(s, &mv, &nv);
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
    }
    if (opt.self_gravity) {
        // pressureGradientSH (spatialOperators.cpp:387-462), dead code at reference HEAD: opt-in only
        const int l_max = cfg.get_int("sh degree");
        const std::vector<double>* factor = cfg.surface_type == odis::FREE_LOADING ? &cfg.loading_factor
                                            : (cfg.surface_type == odis::LID_LOVE || cfg.surface_type == odis::LID_MEMBR) ? &cfg.shell_factor_beta : nullptr;
        if (!factor || (int)factor->size() < l_max + 1)
            return terminate(ODIS_ERR_CONFIG, "self-gravity needs a FREE_LOADING or LID_* surface with its per-degree factors (boundaryConditions.cpp)");
        rc = on_all([&](odis_solver* h) { return odis_enable_self_gravity(h, &mv, l_max, factor->data(), opt.self_gravity == 2 ? 1 : 0); });
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
        log.out("self-gravity / shell pressure term: spherical harmonics to degree " + std::to_string(l_max));
    }

    std::vector<double> v((size_t)F, 0.0), eta((size_t)N, 0.0), dv((size_t)F * 3, 0.0), de((size_t)N * 3, 0.0);
    if (cfg.initial_condition == odis::INIT_LOAD) {                           // initialConditions.cpp:19-144
        const std::string fv = dir + "/InitialConditions/vel_init.txt", fp = dir + "/InitialConditions/pres_init.txt";
        // a missing file is the reference's warning (initialConditions.cpp:40-44: the run goes on from zeros); a file that does not
        // parse would leave half-loaded state, so that ends the run
        const int lv = load_restart(fv, (size_t)F, v, dv), lp = load_restart(fp, (size_t)N, eta, de);
        if (lv == -2 || lp == -2) return terminate(ODIS_ERR_CONFIG, "initial conditions file does not parse: " + (lv == -2 ? fv : fp));
        if (lv == 0) log.out("\nFound initial conditions file: " + fv);
        else log.err("WARNING: NO INITIAL CONDITION FILE FOUND " + fv);
        if (lp == 0) log.out("\nFound initial conditions file: " + fp);
        else log.err("WARNING: NO INITIAL CONDITION FILE FOUND " + fp);
        rc = on_all([&](odis_solver* h) { return odis_set_state(h, v.data(), eta.data(), dv.data(), de.data(), 0); });
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
    }
    if (cfg.initial_condition == odis::INIT_ANALYTICAL) {                     // initialConditions.cpp:311-313
        rc = odis_analytical_state(&mv, &p, v.data(), dv.data(), eta.data(), de.data());
        if (rc == ODIS_OK) rc = on_all([&](odis_solver* h) { return odis_set_state(h, v.data(), eta.data(), dv.data(), de.data(), 0); });
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
    }
    log.out("Defining arrays for Adams-Bashforth time integration...");       // timeIntegrator.cpp:140

    const bool stepping = cfg.surface_type == odis::FREE || cfg.surface_type == odis::FREE_LOADING ||
                          cfg.surface_type == odis::LID_LOVE || cfg.surface_type == odis::LID_MEMBR;      // timeIntegrator.cpp:207-210
    const double r = cfg.get_double("radius"), period = cfg.get_double("orbital period");
    std::vector<double> ven((size_t)F * 2), ediss((size_t)F), vcur((size_t)F);
    std::vector<float> fa((size_t)std::max(F, N)), fb((size_t)F);
    int out_count = 1;
    int64_t iter = 0;
    double e_diss = 0.0;
    auto dump = [&](double current_time) -> int {                             // timeIntegrator.cpp:190-196,296-302 + DumpData
        int rc2 = dissipation_all(&e_diss);
        if (rc2) return rc2;
        log.out(fmt("DUMPING DATA AT %f AVG DISS: %e GW%d", current_time / period, e_diss * 4 * odis::kPi * r * r / 1e9, out_count));
        const uint64_t row = (uint64_t)(out_count - 1);
        if (row < T) {                                                         // beyond the extent HDF5 refuses the selection
            if (ds_u >= 0) {
                if ((rc2 = get_field_all(ODIS_FIELD_VELOCITY_EN, ven))) return rc2;
                for (int i = 0; i < F; i++) { fa[i] = (float)ven[(size_t)i * 2]; fb[i] = (float)ven[(size_t)i * 2 + 1]; }   // outFiles.cpp:546-553
                wr(h5.write_rows(ds_u, row, 1, fa.data(), err));
                wr(h5.write_rows(ds_v, row, 1, fb.data(), err));
            }
            if (ds_ux >= 0) {                                                  // interpolateVelocityCartRBF, interpolation.cpp:116; outFiles.cpp:567-590
                if ((rc2 = get_field_all(ODIS_FIELD_VELOCITY, vcur))) return rc2;
                const odis::Csr& A = nlt.rbf_interp;
                for (int c = 0; c < 3; c++) {
                    for (int i = 0; i < N; i++) {
                        double tmp = 0;
                        for (int k = A.indptr[(size_t)3 * i + c]; k < A.indptr[(size_t)3 * i + c + 1]; k++) tmp += A.data[(size_t)k] * vcur[(size_t)A.indices[(size_t)k]];
                        fa[i] = (float)tmp;
                    }
                    wr(h5.write_rows(c == 0 ? ds_ux : c == 1 ? ds_uy : ds_uz, row, 1, fa.data(), err));
                }
            }
            if (ds_eta >= 0) {
                if ((rc2 = get_field_all(ODIS_FIELD_ETA, eta))) return rc2;
                for (int i = 0; i < N; i++) fa[i] = (float)eta[i];
                wr(h5.write_rows(ds_eta, row, 1, fa.data(), err));
            }
            if (ds_diss >= 0) {
                if ((rc2 = get_field_all(ODIS_FIELD_DISSIPATION, ediss))) return rc2;
                for (int i = 0; i < F; i++) fa[i] = (float)ediss[i];
                wr(h5.write_rows(ds_diss, row, 1, fa.data(), err));
            }
            if (ds_avg >= 0) { const float x = (float)e_diss; wr(h5.write_rows(ds_avg, row, 1, &x, err)); }
            if (ds_kin >= 0) { const float x = (float)current_time; wr(h5.write_rows(ds_kin, row, 1, &x, err)); }   // pp[] points at current_time, timeIntegrator.cpp:150
            if (ds_d1 >= 0) { std::fill(fa.begin(), fa.begin() + N, 0.0f); wr(h5.write_rows(ds_d1, row, 1, fa.data(), err)); }
        }
        out_count++;
        res->dumps++;
        return ODIS_OK;
    };

    // Overlapped output (options.overlap_output): a dump is split into begin (the copy-out is enqueued behind the steps taken so far,
    // odis_snapshot_begin) and finish (wait for the copy, log line, float conversion, data.h5 rows). The next interval's steps are
    // enqueued between the two, so the GPU computes them while the host writes. Same files as the synchronous path.
    // Partitioned runs: every rank's snapshot holds its OWN entries, compact; they are placed by the rank's partition map.
    const bool overlap = opt.overlap_output != 0;
    std::vector<std::vector<int32_t>> own_cells((size_t)world), own_edges((size_t)world);
    std::vector<double> g_eta, g_ven, g_diss, g_v;
    if (overlap && world > 1) {
        for (int k = 0; k < world && rc == ODIS_OK; k++) {
            int32_t nc = 0, ne = 0;
            rc = odis_get_partition(ranks[(size_t)k], nullptr, nullptr, &nc, &ne, nullptr, nullptr, nullptr);
            if (rc != ODIS_OK) break;
            own_cells[(size_t)k].resize((size_t)nc); own_edges[(size_t)k].resize((size_t)ne);
            rc = odis_get_partition_map(ranks[(size_t)k], own_cells[(size_t)k].data(), own_edges[(size_t)k].data());
        }
        if (rc != ODIS_OK) return terminate(rc, odis_last_error());
        g_eta.assign((size_t)N, 0.0); g_ven.assign((size_t)F * 2, 0.0); g_diss.assign((size_t)F, 0.0); g_v.assign((size_t)F, 0.0);
    }
    uint32_t snap_fields = 0;
    if (ds_u >= 0) snap_fields |= ODIS_SNAP_VELOCITY_EN;
    if (ds_ux >= 0) snap_fields |= ODIS_SNAP_VELOCITY;
    if (ds_eta >= 0) snap_fields |= ODIS_SNAP_ETA;
    if (ds_diss >= 0) snap_fields |= ODIS_SNAP_DISSIPATION;
    int snap_slot = 0, pending_slot = -1;
    double pending_time = 0.0;
    auto begin_dump = [&](double current_time) -> int {
        const int rc2 = on_all([&](odis_solver* h) { return odis_snapshot_begin(h, snap_slot, snap_fields); });
        if (rc2) return rc2;
        pending_slot = snap_slot;
        pending_time = current_time;
        snap_slot ^= 1;
        return ODIS_OK;
    };
    auto finish_dump = [&]() -> int {
        if (pending_slot < 0) return ODIS_OK;
        odis_snapshot_view view{};
        int rc2 = odis_snapshot_wait(s, pending_slot, &view);
        if (rc2 == ODIS_OK && world > 1) {
            // the ranks' own entries -> reference-ordered arrays of the whole grid; the dissipation sum is the sum of the ranks' shares
            double tot = 0.0;
            for (int k = 0; k < world && rc2 == ODIS_OK; k++) {
                odis_snapshot_view vk{};
                if (k > 0) rc2 = odis_snapshot_wait(ranks[(size_t)k], pending_slot, &vk);
                else vk = view;
                if (rc2 != ODIS_OK) break;
                tot += vk.dissipation_avg;
                const std::vector<int32_t>& cm = own_cells[(size_t)k];
                const std::vector<int32_t>& em = own_edges[(size_t)k];
                if (vk.eta) for (size_t i = 0; i < cm.size(); i++) g_eta[(size_t)cm[i]] = vk.eta[i];
                if (vk.velocity_en) for (size_t i = 0; i < em.size(); i++) { g_ven[(size_t)em[i] * 2] = vk.velocity_en[2 * i]; g_ven[(size_t)em[i] * 2 + 1] = vk.velocity_en[2 * i + 1]; }
                if (vk.dissipation) for (size_t i = 0; i < em.size(); i++) g_diss[(size_t)em[i]] = vk.dissipation[i];
                if (vk.velocity) for (size_t i = 0; i < em.size(); i++) g_v[(size_t)em[i]] = vk.velocity[i];
            }
            view.dissipation_avg = tot;
            if (view.eta) view.eta = g_eta.data();
            if (view.velocity_en) view.velocity_en = g_ven.data();
            if (view.dissipation) view.dissipation = g_diss.data();
            if (view.velocity) view.velocity = g_v.data();
        }
        pending_slot = -1;
        if (rc2) return rc2;
        e_diss = view.dissipation_avg;
        log.out(fmt("DUMPING DATA AT %f AVG DISS: %e GW%d", pending_time / period, e_diss * 4 * odis::kPi * r * r / 1e9, out_count));
        const uint64_t row = (uint64_t)(out_count - 1);
        if (row < T) {
            if (ds_u >= 0) {
                for (int i = 0; i < F; i++) { fa[i] = (float)view.velocity_en[(size_t)i * 2]; fb[i] = (float)view.velocity_en[(size_t)i * 2 + 1]; }
                wr(h5.write_rows(ds_u, row, 1, fa.data(), err));
                wr(h5.write_rows(ds_v, row, 1, fb.data(), err));
            }
            if (ds_ux >= 0) {
                const odis::Csr& A = nlt.rbf_interp;
                for (int c = 0; c < 3; c++) {
                    for (int i = 0; i < N; i++) {
                        double tmp = 0;
                        for (int k = A.indptr[(size_t)3 * i + c]; k < A.indptr[(size_t)3 * i + c + 1]; k++) tmp += A.data[(size_t)k] * view.velocity[(size_t)A.indices[(size_t)k]];
                        fa[i] = (float)tmp;
                    }
                    wr(h5.write_rows(c == 0 ? ds_ux : c == 1 ? ds_uy : ds_uz, row, 1, fa.data(), err));
                }
            }
            if (ds_eta >= 0) {
                for (int i = 0; i < N; i++) fa[i] = (float)view.eta[i];
                wr(h5.write_rows(ds_eta, row, 1, fa.data(), err));
            }
            if (ds_diss >= 0) {
                for (int i = 0; i < F; i++) fa[i] = (float)view.dissipation[i];
                wr(h5.write_rows(ds_diss, row, 1, fa.data(), err));
            }
            if (ds_avg >= 0) { const float x = (float)e_diss; wr(h5.write_rows(ds_avg, row, 1, &x, err)); }
            if (ds_kin >= 0) { const float x = (float)pending_time; wr(h5.write_rows(ds_kin, row, 1, &x, err)); }
            if (ds_d1 >= 0) { std::fill(fa.begin(), fa.begin() + N, 0.0f); wr(h5.write_rows(ds_d1, row, 1, fa.data(), err)); }
        }
        out_count++;
        res->dumps++;
        return ODIS_OK;
    };

    g_sigint = 0;
    struct sigaction sa_new {}, sa_old {};
    sa_new.sa_handler = on_sigint;
    sigaction(SIGINT, &sa_new, &sa_old);                                       // timeIntegrator.cpp:120
    rc = overlap ? begin_dump(dt * (double)iter) : dump(dt * (double)iter);
    const double bound = (double)total_iter * end_time;                        // timeIntegrator.cpp:205
    while (rc == ODIS_OK && (double)iter < bound) {
        // advance to the next output step, the loop bound or the caller's step budget, whichever is first
        int64_t n = out_freq - iter % out_freq;
        const int64_t left = (int64_t)std::ceil(bound - (double)iter);
        if (n > left) n = left;
        if (opt.max_steps > 0 && iter + n > opt.max_steps) n = opt.max_steps - iter;
        if (n <= 0) break;
        if (stepping) rc = on_all([&](odis_solver* h) { return odis_step(h, (int32_t)n); });   // asynchronous: enqueued behind a pending snapshot
        if (rc != ODIS_OK) break;
        if (overlap && (rc = finish_dump()) != ODIS_OK) break;                 // the previous dump is written while these steps run
        iter += n;
        if (iter % out_freq == 0) rc = overlap ? begin_dump(dt * (double)iter) : dump(dt * (double)iter);   // timeIntegrator.cpp:280-304
        else rc = on_all([&](odis_solver* h) { return odis_synchronize(h); });
        if (rc == ODIS_OK && io_rc != ODIS_OK) break;                          // stop at the first failed write
        // the run reads only the newest entry of the per-step dissipation series: forget the older ones, so that a 150-orbit run
        // (7 million steps) keeps a few KB of series instead of growing it by 56 B per step
        if (rc == ODIS_OK) rc = on_all([&](odis_solver* h) { return odis_trim_dissipation_series(h); });
        if (g_sigint) {                                                        // :307-312
            if (overlap && rc == ODIS_OK) rc = finish_dump();
            log.out("Terminate signal caught...");
            break;
        }
    }
    if (overlap && rc == ODIS_OK) rc = finish_dump();
    sigaction(SIGINT, &sa_old, nullptr);
    if (rc != ODIS_OK) return terminate(rc, odis_last_error());
    if (io_rc != ODIS_OK) return terminate(io_rc, "writing DATA/data.h5 failed: " + io_err);

    // restart files (writeInitialConditions, initialConditions.cpp:209-276)
    ::mkdir((dir + "/InitialConditions").c_str(), 0770);
    if (get_field_all(ODIS_FIELD_VELOCITY, v) == ODIS_OK && get_field_all(ODIS_FIELD_DVDT, dv) == ODIS_OK &&
        get_field_all(ODIS_FIELD_ETA, eta) == ODIS_OK && get_field_all(ODIS_FIELD_DETADT, de) == ODIS_OK) {
        write_restart(dir + "/InitialConditions/vel_init.txt", (size_t)F, v, dv);
        write_restart(dir + "/InitialConditions/pres_init.txt", (size_t)N, eta, de);
    }
    int64_t launches = 0;
    for (int r = 0; r < world; r++) {
        int64_t l = 0;
        odis_get_launch_count(ranks[(size_t)r], &l);
        launches += l;
    }
    for (auto& h : ranks) { odis_destroy(h); h = nullptr; }
    if (h5.close(err) != 0) return terminate(ODIS_ERR_IO, "closing DATA/data.h5 failed: " + err);
    res->steps = iter;
    res->n_cells = N; res->n_edges = F;
    res->dt = dt; res->steps_per_period = total_iter;
    res->last_dissipation_avg = e_diss;
    res->interrupted = g_sigint ? 1 : 0;
    res->kernel_launches = launches;
    if (g_sigint) log.out("SOLVER RETURNED WITH AN ERROR...\n");                // solver.cpp:52-55
    else log.out("Calculations appear to have finished!");                     // solver.cpp:57-58
    return ODIS_OK;
}

}  // extern "C"
