// Device engine behind the "solver" group of include/odis_b200.h: owns the device tables, the
// state and its renumbering, and sequences the two kernels of a time step.
//
// Replaces the state ownership and loop of ab3Explicit (/root/reference/src/timeIntegrator.cpp:70-102
// allocation, :205-313 loop). Memory kept per edge: {v,l_e} x2 (ping-pong), two AB3 history levels;
// per cell: {eta,U}, two history levels — the reference keeps 11 F-sized and 13 N-sized arrays and
// physically shifts its [.,3] histories every step.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/odis_b200.h"
#include "odis_error.h"
#include "odis_kernels.cuh"
#include "odis_reorder.h"
#include "odis_sphere.h"

using odis::fail;

#define ODIS_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess)                                                                    \
            return fail(ODIS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));       \
    } while (0)

struct odis_solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int N = 0, F = 0;
    odis_params prm{};
    odis::Physics phys{};
    double forcing_radius = 0.0;
    std::vector<int> cell_perm, cell_inv, edge_perm, edge_inv;   // perm[new] = old, inv[old] = new

    // device tables
    int2* d_cells = nullptr;
    double2* d_grad = nullptr;
    double *d_fcor = nullptr, *d_dist = nullptr, *d_sw = nullptr;
    int* d_sid = nullptr;
    double2* d_normal = nullptr;
    int* d_eid = nullptr;
    double *d_area = nullptr, *d_trig = nullptr, *d_trig_sq = nullptr;
    // device state
    double2* d_vl[2] = {nullptr, nullptr};
    int cur = 0;
    double2* d_eu = nullptr;
    double* d_hv[2] = {nullptr, nullptr};
    double* d_he[2] = {nullptr, nullptr};
    int hv1 = 0, he1 = 0;            // which of the two arrays holds history level 1
    double* d_block_partial = nullptr;
    unsigned int* d_ticket = nullptr;
    double* d_series = nullptr;
    size_t series_cap = 0;
    double2* d_vavg = nullptr;
    double* d_ediss = nullptr;
    int *d_edge_perm = nullptr, *d_cell_perm = nullptr;   // perm[new] = old, for renumbering on the device
    double* d_stage = nullptr;                             // 3F doubles: reference-ordered staging for H2D / D2H
    double *d_lvl0_v = nullptr, *d_lvl0_e = nullptr;       // AB3 history level 0 as loaded (device order)

    int64_t iter = 0, iter0 = 0;
    bool have_state = false, diag_current = false;
    int last_mode = -1;
    int64_t launches = 0;
    size_t device_bytes = 0;

    odis::EdgeTables edge_tables() const {
        odis::EdgeTables t;
        t.n_edges = F; t.cells = d_cells; t.grad = d_grad; t.fcor = d_fcor; t.dist = d_dist; t.sid = d_sid; t.sw = d_sw;
        return t;
    }
    odis::CellTables cell_tables() const {
        odis::CellTables t;
        t.n_cells = N; t.eid = d_eid; t.area = d_area; t.trig = d_trig; t.trig_sq = d_trig_sq;
        return t;
    }
};

namespace {

template <typename T>
int dev_alloc(odis_solver* s, T** p, size_t count) {
    ODIS_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    s->device_bytes += count * sizeof(T);
    return ODIS_OK;
}
template <typename T>
int upload(odis_solver* s, T** p, const std::vector<T>& h) {
    int rc = dev_alloc(s, p, h.size());
    if (rc) return rc;
    ODIS_CUDA(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    return ODIS_OK;
}

// Host-evaluated time factors for the potential at `time` (tidalPotentials.cpp:55-61).
odis::StepScalars step_scalars(double omega, double time) {
    odis::StepScalars m;
    m.cosM = std::cos(omega * time);
    m.sinM = std::sin(omega * time);
    m.cos2M = std::cos(2 * omega * time);
    m.sin2M = std::sin(2 * omega * time);
    m.cos3M = std::cos(3 * omega * time);
    m.cos4M = std::cos(4 * omega * time);
    return m;
}

int ab3_mode(const odis_solver* s, int64_t iter) {
    if (iter > 1 || s->prm.init_load) return odis::AB3_FULL;      // temporalOperators.cpp:36
    return iter == 0 ? odis::AB3_FIRST : odis::AB3_SECOND;
}

int ensure_series(odis_solver* s, size_t need) {
    if (need <= s->series_cap) return ODIS_OK;
    size_t cap = s->series_cap ? s->series_cap : 4096;
    while (cap < need) cap *= 2;
    double* nd = nullptr;
    ODIS_CUDA(cudaMalloc((void**)&nd, cap * sizeof(double)));
    ODIS_CUDA(cudaMemsetAsync(nd, 0, cap * sizeof(double), s->stream));
    if (s->d_series) {
        ODIS_CUDA(cudaMemcpyAsync(nd, s->d_series, s->series_cap * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
        cudaFree(s->d_series);
        s->device_bytes -= s->series_cap * sizeof(double);
    }
    s->device_bytes += cap * sizeof(double);
    s->d_series = nd;
    s->series_cap = cap;
    return ODIS_OK;
}

int run_diagnostics(odis_solver* s, bool want_fields) {
    if (s->diag_current && !want_fields) return ODIS_OK;
    int rc = ensure_series(s, (size_t)(s->iter - s->iter0) + 1);
    if (rc) return rc;
    odis::launch_edge_diagnostics(s->edge_tables(), s->phys, s->d_vl[s->cur], s->d_normal, want_fields ? s->d_vavg : nullptr,
                                  want_fields ? s->d_ediss : nullptr, s->d_block_partial, s->d_ticket,
                                  s->d_series + (s->iter - s->iter0), s->prm.block_threads, s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    s->diag_current = true;
    return ODIS_OK;
}

}  // namespace

extern "C" {

int odis_create(const odis_mesh_view* mv, const odis_params* prm, int32_t device, odis_solver** out) {
    if (!mv || !prm || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (mv->n_cells < 12 || mv->n_edges != 3 * mv->n_cells - 6) return fail(ODIS_ERR_ARG, "mesh sizes are inconsistent (F != 3N-6)");
    if (!(prm->dt > 0.0) || !(prm->radius > 0.0)) return fail(ODIS_ERR_ARG, "dt and radius must be positive");
    switch (prm->potential) {
        case odis::P_OBLIQ: case odis::P_OBLIQ_WEST: case odis::P_ECC: case odis::P_FULL: case odis::P_FULL2: case odis::P_NONE: break;
        default:
            return fail(ODIS_ERR_UNSUPPORTED, "potential type has no expression in the reference (tidalPotentials.cpp:80-285) or is outside the hot path");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(ODIS_ERR_CUDA, "no CUDA device available: the LTE solver has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(ODIS_ERR_ARG, "device ordinal out of range");
    ODIS_CUDA(cudaSetDevice(device));

    odis_solver* s = new odis_solver();
    s->device = device;
    s->N = mv->n_cells;
    s->F = mv->n_edges;
    s->prm = *prm;
    const int N = s->N, F = s->F;
    int rc = ODIS_OK;
    auto bail = [&](int code) { odis_destroy(s); return code; };
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "cudaStreamCreate failed"));
    cudaEventCreate(&s->ev0);
    cudaEventCreate(&s->ev1);

    // ---- physics scalars ----
    odis::Physics& ph = s->phys;
    ph.g = prm->g; ph.h = prm->h; ph.alpha = prm->alpha; ph.dt = prm->dt;
    ph.ecc = prm->ecc; ph.obl = prm->obl; ph.potential = prm->potential; ph.friction = prm->friction;
    double radius = prm->radius;
    ph.area_sphere_inv = 0.0;
    if (prm->surface == 2 /*LID_LOVE*/ || prm->surface == 3 /*LID_MEMBR*/) radius += prm->shell_thickness;   // tidalPotentials.cpp:50-53
    s->forcing_radius = radius;
    const double om2 = prm->omega * prm->omega, r2 = radius * radius;   // pow(x,2.0)
    ph.factor = 0.0; ph.factor2 = 0.0;
    switch (prm->potential) {
        case odis::P_ECC: ph.factor = 0.75 * prm->love_reduct * om2 * r2 * prm->ecc; break;                       // :84
        case odis::P_OBLIQ: ph.factor = -3. / 2. * prm->love_reduct * om2 * r2 * prm->obl; break;                 // :106
        case odis::P_OBLIQ_WEST: ph.factor = 0.5 * prm->love_reduct * om2 * r2 * prm->obl; break;                 // :120
        case odis::P_FULL2: ph.factor = 1 / 32. * prm->love_reduct * om2 * r2; break;                             // :135
        case odis::P_FULL:                                                                                         // :160-162
            ph.factor = 0.75 * prm->love_reduct * om2 * r2 * prm->ecc;
            ph.factor2 = -3. / 2. * prm->love_reduct * om2 * r2 * prm->obl;
            break;
        default: break;
    }

    // ---- renumbering ----
    s->cell_perm = odis::cell_locality_order(N, mv->node_pos_sph, prm->reorder == 0);
    s->cell_inv = odis::invert_permutation(s->cell_perm);
    s->edge_perm = odis::edge_locality_order(F, mv->face_nodes, s->cell_inv, prm->reorder == 0);
    s->edge_inv = odis::invert_permutation(s->edge_perm);

    // ---- edge tables ----
    {
        std::vector<int2> cells((size_t)F);
        std::vector<double2> grad((size_t)F), normal((size_t)F), vl((size_t)F);
        std::vector<double> fcor((size_t)F), dist((size_t)F), sw((size_t)F * odis::kStencil, 0.0);
        std::vector<int> sid((size_t)F * odis::kStencil, -1);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int en = 0; en < F; en++) {
            const int eo = s->edge_perm[en];
            const int c0 = mv->face_nodes[(size_t)eo * 2], c1 = mv->face_nodes[(size_t)eo * 2 + 1];
            cells[en] = make_int2(s->cell_inv[c0], s->cell_inv[c1]);
            const double d = mv->face_node_dist[eo];
            grad[en] = make_double2((-mv->face_centre_m[(size_t)eo * 2]) / d, (mv->face_centre_m[(size_t)eo * 2 + 1]) / d);   // mesh.cpp:3076-3080
            fcor[en] = -2.0 * prm->omega * std::sin(mv->face_centre_pos_sph[(size_t)eo * 2]);                                   // mesh.cpp:2881
            dist[en] = d;
            normal[en] = make_double2(mv->face_normal_vec_map[(size_t)eo * 2], mv->face_normal_vec_map[(size_t)eo * 2 + 1]);
            vl[en] = make_double2(0.0, mv->face_len[eo]);
            int cnt = 10;                                                                                                       // mesh.cpp:2866-2872
            if (mv->node_friends[(size_t)c0 * 6 + 5] < 0) cnt--;
            if (mv->node_friends[(size_t)c1 * 6 + 5] < 0) cnt--;
            int ids[10]; double ws[10];
            for (int j = 0; j < cnt; j++) { ids[j] = mv->face_interp_friends[(size_t)eo * 10 + j]; ws[j] = mv->face_interp_weights[(size_t)eo * 10 + j]; }
            for (int a = 1; a < cnt; a++) {                  // CSR column order: ascending reference edge id
                const int id = ids[a]; const double w = ws[a];
                int b = a - 1;
                while (b >= 0 && ids[b] > id) { ids[b + 1] = ids[b]; ws[b + 1] = ws[b]; b--; }
                ids[b + 1] = id; ws[b + 1] = w;
            }
            for (int j = 0; j < cnt; j++) {
                if (ids[j] < 0 || ids[j] >= F) { bad++; continue; }
                sid[(size_t)j * F + en] = s->edge_inv[ids[j]];
                sw[(size_t)j * F + en] = ws[j];
            }
        }
        if (bad) return bail(fail(ODIS_ERR_ARG, "face_interp_friends holds out-of-range edge ids"));
        if ((rc = upload(s, &s->d_cells, cells)) || (rc = upload(s, &s->d_grad, grad)) || (rc = upload(s, &s->d_fcor, fcor)) ||
            (rc = upload(s, &s->d_dist, dist)) || (rc = upload(s, &s->d_sid, sid)) || (rc = upload(s, &s->d_sw, sw)) ||
            (rc = upload(s, &s->d_normal, normal)) || (rc = upload(s, &s->d_vl[0], vl)) || (rc = upload(s, &s->d_vl[1], vl)))
            return bail(rc);
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "table upload failed"));
    }
    // ---- cell tables ----
    {
        std::vector<int> eid((size_t)N * odis::kCellEdges, -1);
        std::vector<double> area((size_t)N), trig((size_t)N * 8), trig_sq((size_t)N * 2);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int cn = 0; cn < N; cn++) {
            const int co = s->cell_perm[cn];
            const int n = (mv->node_friends[(size_t)co * 6 + 5] < 0) ? 5 : 6;
            int ids[6], dirs[6];
            for (int j = 0; j < n; j++) { ids[j] = mv->faces[(size_t)co * 6 + j]; dirs[j] = mv->node_face_dir[(size_t)co * 6 + j]; }
            for (int a = 1; a < n; a++) {                    // CSR column order of operatorDivergence
                const int id = ids[a], dr = dirs[a];
                int b = a - 1;
                while (b >= 0 && ids[b] > id) { ids[b + 1] = ids[b]; dirs[b + 1] = dirs[b]; b--; }
                ids[b + 1] = id; dirs[b + 1] = dr;
            }
            for (int j = 0; j < n; j++) {
                if (ids[j] < 0 || ids[j] >= F) { bad++; continue; }
                eid[(size_t)j * N + cn] = s->edge_inv[ids[j]] | (dirs[j] < 0 ? (int)0x80000000 : 0);
            }
            area[cn] = mv->control_volume_surf_area_map[co];
            const double lat = mv->node_pos_sph[(size_t)co * 2], lon = mv->node_pos_sph[(size_t)co * 2 + 1];
            trig[0 * (size_t)N + cn] = std::cos(lat);            // mesh.cpp:2132-2145
            trig[1 * (size_t)N + cn] = std::sin(lat);
            trig[2 * (size_t)N + cn] = std::cos(lon);
            trig[3 * (size_t)N + cn] = std::sin(lon);
            trig[4 * (size_t)N + cn] = std::cos(2.0 * lat);
            trig[5 * (size_t)N + cn] = std::sin(2.0 * lat);
            trig[6 * (size_t)N + cn] = std::cos(2.0 * lon);
            trig[7 * (size_t)N + cn] = std::sin(2.0 * lon);
            trig_sq[cn] = std::cos(lat) * std::cos(lat);
            trig_sq[(size_t)N + cn] = std::sin(lat) * std::sin(lat);
        }
        if (bad) return bail(fail(ODIS_ERR_ARG, "faces table holds out-of-range edge ids"));
        if ((rc = upload(s, &s->d_eid, eid)) || (rc = upload(s, &s->d_area, area)) || (rc = upload(s, &s->d_trig, trig)) ||
            (rc = upload(s, &s->d_trig_sq, trig_sq)))
            return bail(rc);
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "table upload failed"));
    }
    // ---- state ----
    const int blocks = (F + 31) / 32;      // one energy partial per warp of edges (>= blocks of edge_diagnostics)
    if ((rc = dev_alloc(s, &s->d_eu, (size_t)N)) || (rc = dev_alloc(s, &s->d_hv[0], (size_t)F)) || (rc = dev_alloc(s, &s->d_hv[1], (size_t)F)) ||
        (rc = dev_alloc(s, &s->d_he[0], (size_t)N)) || (rc = dev_alloc(s, &s->d_he[1], (size_t)N)) ||
        (rc = dev_alloc(s, &s->d_block_partial, (size_t)blocks)) || (rc = dev_alloc(s, &s->d_ticket, (size_t)1)) ||
        (rc = dev_alloc(s, &s->d_vavg, (size_t)F)) || (rc = dev_alloc(s, &s->d_ediss, (size_t)F)))
        return bail(rc);
    if ((rc = upload(s, &s->d_edge_perm, s->edge_perm)) || (rc = upload(s, &s->d_cell_perm, s->cell_perm)) ||
        (rc = dev_alloc(s, &s->d_stage, (size_t)F * 3)) || (rc = dev_alloc(s, &s->d_lvl0_v, (size_t)F)) ||
        (rc = dev_alloc(s, &s->d_lvl0_e, (size_t)N)))
        return bail(rc);
    cudaMemsetAsync(s->d_ticket, 0, sizeof(unsigned int), s->stream);
    *out = s;
    rc = odis_set_state(s, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc) { *out = nullptr; return bail(rc); }
    return ODIS_OK;
}

int odis_set_state(odis_solver* s, const double* v, const double* eta, const double* dvdt, const double* detadt, int64_t iter) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (iter < 0) return fail(ODIS_ERR_ARG, "iter must be >= 0");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int N = s->N, F = s->F;
    // Reference-ordered host arrays go through one device staging buffer and are renumbered by small
    // kernels; velocities keep the static edge length beside them (only .x is rewritten).
    auto stage = [&](const double* host, size_t n) -> int {
        if (!host) return ODIS_OK;
        ODIS_CUDA(cudaMemcpyAsync(s->d_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        return ODIS_OK;
    };
    int rc;
    if ((rc = stage(v, (size_t)F))) return rc;
    odis::launch_scatter_x(F, s->d_edge_perm, v ? s->d_stage : nullptr, s->d_vl[s->cur], 0, s->stream);
    if ((rc = stage(dvdt, (size_t)F * 3))) return rc;
    s->hv1 = 0;
    odis::launch_scatter_history(F, s->d_edge_perm, dvdt ? s->d_stage : nullptr, s->d_lvl0_v, s->d_hv[0], s->d_hv[1], s->stream);
    if ((rc = stage(eta, (size_t)N))) return rc;
    odis::launch_scatter_x(N, s->d_cell_perm, eta ? s->d_stage : nullptr, s->d_eu, 1, s->stream);
    if ((rc = stage(detadt, (size_t)N * 3))) return rc;
    s->he1 = 0;
    odis::launch_scatter_history(N, s->d_cell_perm, detadt ? s->d_stage : nullptr, s->d_lvl0_e, s->d_he[0], s->d_he[1], s->stream);
    s->launches += 4;
    s->iter = iter;
    s->iter0 = iter;
    s->last_mode = -1;
    s->diag_current = false;
    s->have_state = true;
    // potential for the first step: forcing(current_time + dt), timeIntegrator.cpp:187,218
    const double t = s->prm.dt * (double)iter + s->prm.dt;
    odis::CellState cs{s->d_vl[s->cur], s->d_eu, s->d_he[0], s->d_he[1], nullptr, 0, nullptr};
    odis::launch_cell_step(s->cell_tables(), s->phys, cs, odis::AB3_FULL, step_scalars(s->prm.omega, t), 0, s->prm.block_threads, s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

static int step_impl(odis_solver* s, int32_t nsteps, std::vector<cudaEvent_t>* marks);

int odis_step(odis_solver* s, int32_t nsteps) { return step_impl(s, nsteps, nullptr); }

int odis_step_profiled(odis_solver* s, int32_t nsteps, float* edge_ms_out, float* cell_ms_out) {
    if (!s || !edge_ms_out || !cell_ms_out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (nsteps < 0 || nsteps > 100000) return fail(ODIS_ERR_ARG, "nsteps out of range");
    ODIS_CUDA(cudaSetDevice(s->device));
    std::vector<cudaEvent_t> marks((size_t)nsteps * 3);
    for (auto& e : marks) ODIS_CUDA(cudaEventCreate(&e));
    int rc = step_impl(s, nsteps, &marks);
    if (!rc && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = fail(ODIS_ERR_CUDA, "synchronize failed");
    double edge = 0.0, cell = 0.0;
    if (!rc) {
        for (int k = 0; k < nsteps; k++) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, marks[(size_t)k * 3], marks[(size_t)k * 3 + 1]);
            cudaEventElapsedTime(&b, marks[(size_t)k * 3 + 1], marks[(size_t)k * 3 + 2]);
            edge += a; cell += b;
        }
    }
    for (auto& e : marks) cudaEventDestroy(e);
    *edge_ms_out = (float)edge; *cell_ms_out = (float)cell;
    return rc;
}

static int step_impl(odis_solver* s, int32_t nsteps, std::vector<cudaEvent_t>* marks) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "odis_set_state has not been called");
    if (nsteps < 0) return fail(ODIS_ERR_ARG, "nsteps must be >= 0");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ensure_series(s, (size_t)(s->iter - s->iter0) + (size_t)nsteps + 1);
    if (rc) return rc;
    const odis::EdgeTables et = s->edge_tables();
    const odis::CellTables ct = s->cell_tables();
    for (int k = 0; k < nsteps; k++) {
        const int mode = ab3_mode(s, s->iter);
        odis::EdgeState es;
        es.vl_in = s->d_vl[s->cur]; es.vl_out = s->d_vl[1 - s->cur]; es.eu = s->d_eu;
        es.h1 = s->d_hv[s->hv1]; es.h2 = s->d_hv[1 - s->hv1];
        es.block_partial = s->d_block_partial; es.ticket = s->d_ticket;
        es.energy_out = s->d_series + (s->iter - s->iter0);
        if (marks) cudaEventRecord((*marks)[(size_t)k * 3], s->stream);
        odis::launch_edge_step(et, s->phys, es, mode, s->prm.block_threads, s->stream);
        if (marks) cudaEventRecord((*marks)[(size_t)k * 3 + 1], s->stream);
        if (mode == odis::AB3_FULL) s->hv1 = 1 - s->hv1;
        odis::CellState cs{s->d_vl[1 - s->cur], s->d_eu, s->d_he[s->he1], s->d_he[1 - s->he1],
                           s->d_block_partial, (s->F + 31) / 32, es.energy_out};
        // the next step's forcing time: current_time = dt*(iter+1), evaluated at current_time + dt
        const double tnext = s->prm.dt * (double)(s->iter + 1) + s->prm.dt;
        odis::launch_cell_step(ct, s->phys, cs, mode, step_scalars(s->prm.omega, tnext), 1, s->prm.block_threads, s->stream);
        if (marks) cudaEventRecord((*marks)[(size_t)k * 3 + 2], s->stream);
        if (mode == odis::AB3_FULL) s->he1 = 1 - s->he1;
        s->cur = 1 - s->cur;
        s->iter++;
        s->last_mode = mode;
        s->launches += 2;
    }
    if (nsteps > 0) s->diag_current = false;
    ODIS_CUDA(cudaGetLastError());
    return ODIS_OK;
}

int odis_step_timed(odis_solver* s, int32_t nsteps, float* elapsed_ms_out) {
    if (!s || !elapsed_ms_out) return fail(ODIS_ERR_ARG, "NULL argument");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ensure_series(s, (size_t)(s->iter - s->iter0) + (size_t)(nsteps > 0 ? nsteps : 0) + 1);
    if (rc) return rc;
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    ODIS_CUDA(cudaEventRecord(s->ev0, s->stream));
    rc = odis_step(s, nsteps);
    if (rc) return rc;
    ODIS_CUDA(cudaEventRecord(s->ev1, s->stream));
    ODIS_CUDA(cudaEventSynchronize(s->ev1));
    ODIS_CUDA(cudaEventElapsedTime(elapsed_ms_out, s->ev0, s->ev1));
    return ODIS_OK;
}

int odis_get_field(odis_solver* s, int32_t field, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int N = s->N, F = s->F;
    size_t count = 0;
    switch (field) {
        case ODIS_FIELD_VELOCITY:
            odis::launch_gather_component(F, s->d_edge_perm, s->d_vl[s->cur], 0, s->d_stage, s->stream);
            count = (size_t)F;
            break;
        case ODIS_FIELD_ETA:
        case ODIS_FIELD_POTENTIAL:
            odis::launch_gather_component(N, s->d_cell_perm, s->d_eu, field == ODIS_FIELD_ETA ? 0 : 1, s->d_stage, s->stream);
            count = (size_t)N;
            break;
        case ODIS_FIELD_DVDT:
        case ODIS_FIELD_DETADT: {
            // level 0 = newest tendency: as loaded before any step, else where the last step stored it
            // (temporalOperators.cpp:47,56,65)
            const int which0 = s->last_mode < 0 ? 0 : (s->last_mode == odis::AB3_FIRST ? 2 : 1);
            if (field == ODIS_FIELD_DVDT) {
                odis::launch_gather_history(F, s->d_edge_perm, s->d_lvl0_v, s->d_hv[s->hv1], s->d_hv[1 - s->hv1], which0, s->d_stage, s->stream);
                count = (size_t)F * 3;
            } else {
                odis::launch_gather_history(N, s->d_cell_perm, s->d_lvl0_e, s->d_he[s->he1], s->d_he[1 - s->he1], which0, s->d_stage, s->stream);
                count = (size_t)N * 3;
            }
            break;
        }
        case ODIS_FIELD_VELOCITY_EN:
        case ODIS_FIELD_DISSIPATION: {
            int rc = run_diagnostics(s, true);
            if (rc) return rc;
            if (field == ODIS_FIELD_VELOCITY_EN) {
                odis::launch_gather_pair(F, s->d_edge_perm, s->d_vavg, s->d_stage, s->stream);
                count = (size_t)F * 2;
            } else {
                odis::launch_gather_scalar(F, s->d_edge_perm, s->d_ediss, s->d_stage, s->stream);
                count = (size_t)F;
            }
            break;
        }
        default:
            return fail(ODIS_ERR_ARG, "unknown field id");
    }
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaMemcpyAsync(out, s->d_stage, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

static double sphere_area(const odis_solver* s) { return 4 * odis::kPi * (s->prm.radius * s->prm.radius); }   // energy.cpp:60

int odis_get_dissipation_avg(odis_solver* s, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = run_diagnostics(s, false);
    if (rc) return rc;
    double v = 0.0;
    ODIS_CUDA(cudaMemcpyAsync(&v, s->d_series + (s->iter - s->iter0), sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    *out = v / sphere_area(s);
    return ODIS_OK;
}

int odis_get_dissipation_series(odis_solver* s, int64_t first, int64_t count, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    const int64_t have = s->iter - s->iter0 + 1;
    if (first < 0 || count < 0 || first + count > have) return fail(ODIS_ERR_ARG, "series range exceeds the steps taken since odis_set_state");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = run_diagnostics(s, false);
    if (rc) return rc;
    ODIS_CUDA(cudaMemcpyAsync(out, s->d_series + first, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    const double a = sphere_area(s);
    for (int64_t k = 0; k < count; k++) out[k] /= a;
    return ODIS_OK;
}

int odis_get_iter(odis_solver* s, int64_t* iter_out) {
    if (!s || !iter_out) return fail(ODIS_ERR_ARG, "NULL argument");
    *iter_out = s->iter;
    return ODIS_OK;
}

int odis_get_footprint(odis_solver* s, int64_t* device_bytes_out, int64_t* alg_bytes_out) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    if (device_bytes_out) *device_bytes_out = (int64_t)s->device_bytes;
    if (alg_bytes_out) *alg_bytes_out = 200LL * s->F + 128LL * s->N;
    return ODIS_OK;
}

int odis_get_launch_count(odis_solver* s, int64_t* launches_out) {
    if (!s || !launches_out) return fail(ODIS_ERR_ARG, "NULL argument");
    *launches_out = s->launches;
    return ODIS_OK;
}

int odis_synchronize(odis_solver* s) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    ODIS_CUDA(cudaSetDevice(s->device));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

void odis_destroy(odis_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    void* ptrs[] = {s->d_cells, s->d_grad, s->d_fcor, s->d_dist, s->d_sw, s->d_sid, s->d_normal, s->d_eid, s->d_area, s->d_trig,
                    s->d_trig_sq, s->d_vl[0], s->d_vl[1], s->d_eu, s->d_hv[0], s->d_hv[1], s->d_he[0], s->d_he[1],
                    s->d_block_partial, s->d_ticket, s->d_series, s->d_vavg, s->d_ediss, s->d_edge_perm, s->d_cell_perm, s->d_stage,
                    s->d_lvl0_v, s->d_lvl0_e};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // extern "C"
