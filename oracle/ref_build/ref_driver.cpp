// TEST INFRASTRUCTURE — not product code.
//
// Entry point that stands in for the reference's main() (/root/reference/src/main.cpp:46-68)
// when the UNMODIFIED reference sources are linked into oracle/_ref/odis_ref_l<L>. It
// performs the same three calls (new Globals(0); new Mesh(...); solveODIS) from the
// current directory (input.in, input_files/grid_l<L>.txt, DATA/), and adds what a parity
// oracle needs and the reference cannot give: full-precision (FP64, binary) copies of
//   * every mesh table that feeds the hot loop (Mesh's public members),
//   * the arrays handed to OutFiles::DumpData at every dump,
//   * the complete solver state (v, eta and both AB3 histories) after the last step,
// plus wall-clock of the ab3Explicit while-loop. The last two use link-time interposition:
// the reference's outFiles.cpp / initialConditions.cpp are compiled with
// -DDumpData=DumpData_reference / -DwriteInitialConditions=writeInitialConditions_reference
// and the un-renamed symbols that timeIntegrator.cpp calls are defined here, forwarding to
// the renamed originals. No reference source is edited or copied.
//
// usage: odis_ref_l<L> [--no-run] [--quiet-restart]
//   outputs (under DATA/): ref_tables.bin, ref_dumps.bin, ref_final.bin, ref_timing.txt,
//                          h5shim_index.txt + h5shim_<k>.f32
#include "mesh.h"
#include "globals.h"
#include "outFiles.h"
#include "solver.h"
#include "initialConditions.h"
#include "gridConstants.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

FILE* g_tables = nullptr;
FILE* g_dumps = nullptr;
bool g_write_restart = true;
double g_t_loop_start = -1.0, g_t_loop_end = -1.0, g_t_in_dumps = 0.0;
int g_dump_calls = 0;

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// record: u32 name_len, name, u32 dtype (0=f64, 1=i32), u32 ndim, u64 dims[ndim], raw data
void put(FILE* f, const char* name, int dtype, std::vector<uint64_t> dims, const void* data) {
    uint32_t nl = (uint32_t)std::strlen(name), dt = (uint32_t)dtype, nd = (uint32_t)dims.size();
    size_t n = 1;
    for (auto d : dims) n *= (size_t)d;
    std::fwrite(&nl, 4, 1, f); std::fwrite(name, 1, nl, f);
    std::fwrite(&dt, 4, 1, f); std::fwrite(&nd, 4, 1, f);
    std::fwrite(dims.data(), 8, nd, f);
    std::fwrite(data, dtype == 0 ? 8 : 4, n, f);
}
void put_f64(FILE* f, const char* name, std::vector<uint64_t> dims, const double* p) { put(f, name, 0, dims, p); }
void put_i32(FILE* f, const char* name, std::vector<uint64_t> dims, const int* p) { put(f, name, 1, dims, p); }
void put_scalar(FILE* f, const char* name, double v) { put(f, name, 0, {1}, &v); }

void put_csr(FILE* f, const char* name, const SpMat& A) {
    A.flush();
    std::string n(name);
    int shape[2] = {A.rows(), A.cols()};
    put_i32(f, (n + ".shape").c_str(), {2}, shape);
    put_i32(f, (n + ".indptr").c_str(), {(uint64_t)A.ptr.size()}, A.ptr.data());
    put_i32(f, (n + ".indices").c_str(), {(uint64_t)A.idx.size()}, A.idx.data());
    put_f64(f, (n + ".data").c_str(), {(uint64_t)A.val.size()}, A.val.data());
}

void dump_tables(Globals* g, Mesh* m) {
    const uint64_t N = NODE_NUM, F = FACE_NUM, V = VERTEX_NUM;
    FILE* f = g_tables;
    // scalars after Globals/applySurfaceBCs/CalcMaxTimeStep have had their say
    put_scalar(f, "radius", g->radius.Value());
    put_scalar(f, "angVel", g->angVel.Value());
    put_scalar(f, "period", g->period.Value());
    put_scalar(f, "g", g->g.Value());
    put_scalar(f, "h", g->h.Value());
    put_scalar(f, "alpha", g->alpha.Value());
    put_scalar(f, "loveReduct", g->loveReduct.Value());
    put_scalar(f, "shell_thickness", g->shell_thickness.Value());
    put_scalar(f, "e", g->e.Value());
    put_scalar(f, "theta", g->theta.Value());
    put_scalar(f, "timeStep", g->timeStep.Value());
    put_scalar(f, "endTime", g->endTime.Value());
    put_scalar(f, "totalIter", (double)g->totalIter.Value());
    put_scalar(f, "outputTime", (double)g->outputTime.Value());
    put_scalar(f, "tide_type", (double)g->tide_type);
    put_scalar(f, "fric_type", (double)g->fric_type);
    put_scalar(f, "surface_type", (double)g->surface_type);
    put_scalar(f, "advection", (double)g->advection.Value());
    put_scalar(f, "rbf_eps", g->rbf_eps.Value());

    put_f64(f, "node_pos_sph", {N, 2}, &m->node_pos_sph(0, 0));
    put_i32(f, "node_friends", {N, 6}, &m->node_friends(0, 0));
    put_f64(f, "centroid_pos_sph", {N, 6, 2}, &m->centroid_pos_sph(0, 0, 0));
    put_f64(f, "node_pos_map", {N, 7, 2}, &m->node_pos_map(0, 0, 0));
    put_f64(f, "centroid_pos_map", {N, 6, 2}, &m->centroid_pos_map(0, 0, 0));
    put_f64(f, "control_volume_surf_area_map", {N}, &m->control_volume_surf_area_map(0));
    put_f64(f, "control_volume_mass", {N}, &m->control_volume_mass(0));
    put_f64(f, "trigLat", {N, 2}, &m->trigLat(0, 0));
    put_f64(f, "trigLon", {N, 2}, &m->trigLon(0, 0));
    put_f64(f, "trig2Lat", {N, 2}, &m->trig2Lat(0, 0));
    put_f64(f, "trig2Lon", {N, 2}, &m->trig2Lon(0, 0));
    put_f64(f, "trigSqLat", {N, 2}, &m->trigSqLat(0, 0));
    put_f64(f, "trigSqLon", {N, 2}, &m->trigSqLon(0, 0));

    put_i32(f, "faces", {N, 6}, &m->faces(0, 0));
    put_i32(f, "node_face_dir", {N, 6}, &m->node_face_dir(0, 0));
    put_i32(f, "face_nodes", {F, 2}, &m->face_nodes(0, 0));
    put_i32(f, "face_vertexes", {F, 2}, &m->face_vertexes(0, 0));
    put_i32(f, "face_interp_friends", {F, 10}, &m->face_interp_friends(0, 0));
    put_f64(f, "face_interp_weights", {F, 10}, &m->face_interp_weights(0, 0));
    put_f64(f, "face_len", {F}, &m->face_len(0));
    put_f64(f, "face_node_dist", {F}, &m->face_node_dist(0));
    put_f64(f, "face_centre_m", {F, 2}, &m->face_centre_m(0, 0));
    put_f64(f, "face_centre_pos_sph", {F, 2}, &m->face_centre_pos_sph(0, 0));
    put_f64(f, "face_intercept_pos_sph", {F, 2}, &m->face_intercept_pos_sph(0, 0));
    put_f64(f, "face_area", {F}, &m->face_area(0));
    put_f64(f, "face_normal_vec_map", {F, 2}, &m->face_normal_vec_map(0, 0));
    put_f64(f, "face_normal_vec_xyz", {F, 3}, &m->face_normal_vec_xyz(0, 0));

    put_i32(f, "vertexes", {N, 6}, &m->vertexes(0, 0));
    put_f64(f, "vertex_pos_sph", {V, 2}, &m->vertex_pos_sph(0, 0));
    put_i32(f, "vertex_nodes", {V, 3}, &m->vertex_nodes(0, 0));
    put_f64(f, "vertex_R", {V, 3}, &m->vertex_R(0, 0));
    put_i32(f, "vertex_faces", {V, 3}, &m->vertex_faces(0, 0));
    put_i32(f, "vertex_face_dir", {V, 3}, &m->vertex_face_dir(0, 0));
    put_f64(f, "vertex_area", {V}, &m->vertex_area(0));
    put_f64(f, "vertex_sinlat", {V}, &m->vertex_sinlat(0));

    put_csr(f, "operatorGradient", m->operatorGradient);
    put_csr(f, "operatorDivergence", m->operatorDivergence);
    put_csr(f, "operatorCoriolis", m->operatorCoriolis);
    put_csr(f, "operatorLinearDrag", m->operatorLinearDrag);
    put_csr(f, "operatorCurl", m->operatorCurl);
    put_csr(f, "operatorRBFinterp", m->operatorRBFinterp);
    put_csr(f, "operatorDirectionalSecondDeriv", m->operatorDirectionalSecondDeriv);
}

}  // namespace

// The reference's own implementations, renamed at compile time (see header comment).
int writeInitialConditions_reference(Globals*, Mesh*, Array1D<double>&, Array2D<double>&, Array1D<double>&,
                                     Array2D<double>&);
void odis_ref_DumpData_reference(OutFiles*, Globals*, int, double**) __asm__(
    "_ZN8OutFiles18DumpData_referenceEP7GlobalsiPPd");

// Called by ab3Explicit right after its while-loop (/root/reference/src/timeIntegrator.cpp:316).
int writeInitialConditions(Globals* globals, Mesh* mesh, Array1D<double>& v, Array2D<double>& dvdt,
                           Array1D<double>& p, Array2D<double>& dpdt) {
    g_t_loop_end = now_s();
    std::string fn = globals->path + SEP + "DATA" + SEP + "ref_final.bin";
    FILE* f = std::fopen(fn.c_str(), "wb");
    if (f) {
        put_f64(f, "v", {(uint64_t)FACE_NUM}, &v(0));
        put_f64(f, "dvdt", {(uint64_t)FACE_NUM, 3}, &dvdt(0, 0));
        put_f64(f, "eta", {(uint64_t)NODE_NUM}, &p(0));
        put_f64(f, "detadt", {(uint64_t)NODE_NUM, 3}, &dpdt(0, 0));
        std::fclose(f);
    }
    if (g_write_restart) return writeInitialConditions_reference(globals, mesh, v, dvdt, p, dpdt);
    return 1;
}

// Called by ab3Explicit before the loop (slice 1) and at every output step
// (/root/reference/src/timeIntegrator.cpp:195,301).
void OutFiles::DumpData(Globals* globals, int time_level, double** data) {
    const double t0 = now_s();
    if (g_dumps) {
        for (size_t j = 0; j < tags->size(); j++) {
            const std::string& tag = (*tags)[j];
            uint64_t n = 0;
            if (tag == "velocity output") n = 2ull * FACE_NUM;
            else if (tag == "velocity cartesian output") n = 3ull * NODE_NUM;
            else if (tag == "displacement output") n = NODE_NUM;
            else if (tag == "dissipation output") n = FACE_NUM;
            else if (tag == "dissipation avg output") n = 1;
            else if (tag == "kinetic avg output") n = 1;
            else if (tag == "dummy1 output") n = NODE_NUM;
            if (!n) continue;
            std::string name = std::to_string(time_level) + ":" + tag;
            put_f64(g_dumps, name.c_str(), {n}, data[j]);
        }
        std::fflush(g_dumps);
    }
    odis_ref_DumpData_reference(this, globals, time_level, data);
    const double t1 = now_s();
    if (g_dump_calls == 0) g_t_loop_start = t1;   // the first dump immediately precedes the loop
    else g_t_in_dumps += t1 - t0;
    g_dump_calls++;
}

int main(int argc, char** argv) {
    bool run = true;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--no-run")) run = false;
        if (!std::strcmp(argv[i], "--quiet-restart")) g_write_restart = false;
    }
    // identical to the reference's main(): /root/reference/src/main.cpp:52-58
    Globals* constants = new Globals(0);
    Mesh* grid = new Mesh(constants, constants->node_num, constants->face_num, constants->vertex_num,
                          (int)constants->dLat.Value(), constants->l_max.Value());
    constants->OutputConsts();

    const std::string data = constants->path + SEP + "DATA" + SEP;
    g_tables = std::fopen((data + "ref_tables.bin").c_str(), "wb");
    if (g_tables) { dump_tables(constants, grid); std::fclose(g_tables); }

    if (run) {
        g_dumps = std::fopen((data + "ref_dumps.bin").c_str(), "wb");
        solveODIS(constants, grid);                       // main.cpp:65
        if (g_dumps) std::fclose(g_dumps);
        const double steps = (double)constants->totalIter.Value() * constants->endTime.Value();
        FILE* ft = std::fopen((data + "ref_timing.txt").c_str(), "w");
        if (ft) {
            std::fprintf(ft, "loop_seconds %.9f\ndump_seconds_inside_loop %.9f\nloop_bound %.6f\ncells %d\n",
                         g_t_loop_end - g_t_loop_start, g_t_in_dumps, steps, NODE_NUM);
            std::fclose(ft);
        }
    }
    h5shim::dump_all((constants->path + SEP + "DATA").c_str());
    return 0;
}
