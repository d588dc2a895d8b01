// Device engine behind the "solver" group of include/odis_b200.h: owns the device tables, the
// state and its renumbering, and sequences the two kernels of a time step.
//
// Replaces the state ownership and loop of ab3Explicit (/root/reference/src/timeIntegrator.cpp:70-102
// allocation, :205-313 loop). Memory kept per edge: {v,l_e} x2 (ping-pong), two AB3 history levels;
// per cell: {eta,U}, two history levels — the reference keeps 11 F-sized and 13 N-sized arrays and
// physically shifts its [.,3] histories every step.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/odis_b200.h"
#include "odis_error.h"
#include <unistd.h>

#include "odis_kernels.cuh"
#include "odis_kernels_nl.cuh"
#include "odis_partition.h"
#include "odis_reorder.h"
#include "odis_sh.cuh"
#include "odis_sh.h"
#include "odis_sphere.h"

using odis::fail;

#define ODIS_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess)                                                                    \
            return fail(ODIS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));       \
    } while (0)

constexpr int kMaxPeers = 8;

struct odis_solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int Ng = 0, Fg = 0;              // global sizes (what crosses the C ABI)
    int N = 0, F = 0;                // local sizes including the halo
    int No = 0, Fo = 0;              // owned cells / edges (the kernels' iteration spaces)
    int Np = 0, Fp = 0;              // SoA strides: N and Fo rounded up to the pipelined kernels' tile
    bool pipe_edge = true;           // params.reserved[0] bit 0 clear: the staged (bulk-async) kernels; set: the direct-load baseline kernels
    bool pipe_cell = true;           // staged cell update (with the staged edge kernel; ODIS_B200_DIRECT_CELL=1 keeps the direct one for A/B timing)
    int rank = 0, world = 1;
    odis_params prm{};
    odis::Physics phys{};
    double forcing_radius = 0.0;
    std::vector<int> cell_perm, edge_perm;   // local id -> reference id
    odis::Partition part;

    // device tables
    int2* d_cells = nullptr;
    double2* d_grad = nullptr;
    double *d_fcor = nullptr, *d_dist = nullptr, *d_sw = nullptr;
    int* d_sid = nullptr;
    short* d_sid16 = nullptr;                // params.reserved[0] bit 7: narrow stencil ids of the staged edge kernel ([tiles][10][128] offsets)
    unsigned char* d_tile_wide = nullptr;    //   + per-tile flag "an offset does not fit, read this tile from d_sid"
    bool edge_ids16 = false;
    int wide_tiles = 0;
    double2* d_normal = nullptr;
    int* d_eid = nullptr;
    double *d_area = nullptr, *d_trig = nullptr, *d_trig_sq = nullptr;
    // device state
    double2* d_vl[2] = {nullptr, nullptr};
    int cur = 0;
    double2* d_eu[2] = {nullptr, nullptr};   // {eta,U}: cell updates read one and write the other
    int ecur = 0;                    // which holds the newest values
    double* d_hv[2] = {nullptr, nullptr};
    double* d_he[3] = {nullptr, nullptr, nullptr};   // cell tendency history: levels 1, 2 and the slot the next update writes
    int hv1 = 0;                     // which of the two edge arrays holds history level 1
    int he1 = 0, he2 = 1, hefree = 2;
    double* d_block_partial = nullptr;
    unsigned int* d_ticket = nullptr;
    double* d_series = nullptr;
    size_t series_cap = 0;
    odis::StepCtl* d_ctl = nullptr;          // device-side step counter / halo epochs / current time factors
    odis::StepScalars* d_scal = nullptr;     // [series_cap] host-evaluated time factors of step k since set_state
    bool use_graph = true;                   // params.reserved[0] bit 3 clear: steady-state steps replay captured CUDA graphs
    std::map<unsigned, cudaGraphExec_t> graphs;   // one per phase of the buffer rotation (period 6 steps)
    int64_t graph_launches = 0;
    double2* d_vavg = nullptr;
    double* d_ediss = nullptr;
    int *d_edge_perm = nullptr, *d_cell_perm = nullptr;   // local id -> reference id, for renumbering on the device
    double* d_stage = nullptr;                             // 3*Fg doubles: reference-ordered staging for H2D / D2H
    double *d_lvl0_v = nullptr, *d_lvl0_e = nullptr;       // AB3 history level 0 as loaded (device order)

    // halo exchange (world > 1)
    int n_peers = 0;
    int peer_rank[kMaxPeers] = {0};
    bool connected = false;
    int n_send_e = 0, n_send_c = 0;
    int *d_send_e_local = nullptr, *d_send_e_remote = nullptr, *d_send_e_peer = nullptr;
    int *d_send_c_local = nullptr, *d_send_c_remote = nullptr, *d_send_c_peer = nullptr;
    // the edge send list as CSR over the boundary-first own numbering, for the exchange fused into the edge kernel
    int Fb = 0;                                            // boundary edges: local ids [0, Fb)
    int wait_from = 0;                                     // cells [wait_from, N) read ghost edges (own boundary cells, ghost cells)
    int *d_csr_e_first = nullptr, *d_csr_e_peer = nullptr, *d_csr_e_remote = nullptr;
    unsigned int* d_halo_done = nullptr;                   // boundary tiles finished
    unsigned long long* d_flags = nullptr;                 // [2][world] epochs written by the peers
    unsigned int* d_halo_ticket = nullptr;
    odis::HaloRemote remote_v[2], remote_c[2];             // peers' vl[0], vl[1], eu[0], eu[1] + their flag arrays
    void* ipc_opened[kMaxPeers][5] = {{nullptr}};

    // nonlinear branch (odis_enable_advection)
    bool nl_on = false;
    bool nl_folded = true;           // the 4-launch nonlinear step (vertex PV + cell Ekin in one grid, thickness flux / energy diagnostic inside the
                                     // edge update, next potential inside the cell update); the baseline selection (bit 0) keeps the six gather
                                     // launches + diagnostics + potential pass
    int nl_launches() const { return nl_folded ? odis::kNlLaunchesFolded : odis::kNlLaunches; }
    odis::NlTables nl{};
    double *d_nl_qv = nullptr, *d_nl_ekin = nullptr, *d_nl_flux = nullptr;
    double2* d_nl_fq = nullptr;
    std::vector<void*> nl_owned;
    // spherical-harmonic self-gravity / shell-pressure term (odis_enable_self_gravity)
    bool sh_on = false;
    int sh_lmax = 0, sh_rows = 0;
    bool sh_stored = false;          // basis rows kept in HBM (GEMV) instead of rebuilt per cell
    double* d_shRec = nullptr;
    double *d_shY = nullptr, *d_shGinv = nullptr, *d_shFactor = nullptr, *d_sh_partial = nullptr, *d_sh_b = nullptr, *d_sh_s = nullptr;
    std::vector<double> sh_ginv_host;
    // default for degrees 2..4 with the matrix-free basis: the harmonic analysis is folded into the staged cell update, the solve into the
    // synthesis launch (3 launches per step: edge, cell, synthesis)
    // ... and, on hardware, solve + synthesis behind a grid-wide barrier inside the same launch (sh_merged: 2 launches per step)
    bool sh_fused = false, sh_merged = false;
    double* d_sg_cta = nullptr;
    unsigned int* d_sg_ticket = nullptr;     // [0] last-CTA ticket, [8..9] grid barrier (arrival count, generation)
    int sg_cta_stride = 0;
    odis::CellSgAccum sg_accum() const {
        return odis::CellSgAccum{sh_lmax, No, d_sg_cta, sg_cta_stride, d_sh_b, d_sg_ticket, sh_merged ? 1 : 0, d_shGinv, d_shFactor, prm.g, d_sh_s,
                                 d_sg_ticket + 8};
    }
    int sg_cells() const { return world > 1 ? N : No; }        // cells the cell update covers (partitioned: ghost cells too)
    int sh_step_launches() const { return !sh_on ? 0 : (sh_fused ? (sh_merged ? 0 : 1) : sh_launches()); }     // per time step, after the cell update
    unsigned char* d_sh_xblock = nullptr;             // partitioned: this rank's exchange block (odis_sh.cuh), mapped by every other rank
    unsigned long long* d_sh_xctl = nullptr;
    unsigned char* sh_xremote[odis::kShMaxWorld] = {nullptr};
    void* sh_ipc_opened[odis::kShMaxWorld] = {nullptr};
    odis::ShExchange sh_exchange() const {
        odis::ShExchange x;
        x.world = world; x.rank = rank; x.ctl = d_sh_xctl;
        for (int r = 0; r < odis::kShMaxWorld; r++) x.block[r] = r < world ? sh_xremote[r] : nullptr;
        return x;
    }
    odis::ShTables sh_tables() const {
        odis::ShTables t;
        t.rows = sh_rows; t.stride = Np; t.Y = sh_stored ? d_shY : nullptr; t.Ginv = d_shGinv; t.factor = d_shFactor;
        t.l_max = sh_lmax; t.trig = d_trig; t.rec = d_shRec;
        return t;
    }
    odis::ShWork sh_work() const { return odis::ShWork{d_sh_partial, odis::kShMaxBlocks, 0, d_sh_b, d_sh_s}; }
    int sh_launches() const { return !sh_on ? 0 : odis::sh_analysis_launches(sh_rows, world > 1) + 1; }

    // output snapshots that overlap with stepping (odis_snapshot_begin / _wait): two slots, each a device staging buffer in
    // reference order [eta N | v_avg 2F | energy_diss F | v F | dissipation sum 1] and its page-locked host copy
    struct SnapshotSlot {
        double* d_buf = nullptr;
        double* h_buf = nullptr;
        cudaEvent_t ready = nullptr, done = nullptr;
        uint32_t fields = 0;
        bool pending = false;
        int64_t iter = 0;
    } snap[2];
    cudaStream_t copy_stream = nullptr;
    // next state staged while the current interval runs (odis_stage_state / odis_commit_state): reference-ordered device copy
    // [v F | dvdt 3F | eta N | detadt 3N], filled on the copy stream
    double* d_stage_next = nullptr;
    cudaEvent_t staged = nullptr, consumed = nullptr;
    bool stage_pending = false, stage_used = false;
    unsigned stage_mask = 0;                 // bit k: array k of the staged state was given (others are zero)
    double* h_stage_pack = nullptr;          // partitioned: page-locked copy of this rank's share of the staged state (pack_partitioned)

    // partitioned solvers move only the entries they hold across PCIe: page-locked host buffer for the packed state / fields
    double* h_pack = nullptr;
    size_t h_pack_cap = 0;
    cudaEvent_t pack_done = nullptr;         // the last H2D copy out of h_pack
    bool pack_pending = false;

    long long wait_cycles = 0;               // ODIS_B200_WAIT_TIMEOUT_S in clock64 ticks (0: the kernels' default)
    int64_t iter = 0, iter0 = 0;
    bool have_state = false, diag_current = false;
    int last_mode = -1;
    int64_t launches = 0;
    size_t device_bytes = 0;

    odis::EdgeTables edge_tables() const {
        odis::EdgeTables t;
        t.n_edges = Fo; t.stride = Fp; t.cells = d_cells; t.grad = d_grad; t.fcor = d_fcor; t.dist = d_dist; t.sid = d_sid; t.sw = d_sw;
        return t;
    }
    odis::CellTables cell_tables(int n_active) const {
        odis::CellTables t;
        t.n_cells = Np; t.n_active = n_active; t.eid = d_eid; t.area = d_area; t.trig = d_trig; t.trig_sq = d_trig_sq;
        return t;
    }
};

struct HaloBlob {                     // what one rank publishes to the others (odis_halo_export)
    int32_t rank, world;
    int64_t pid;
    void* raw[6];                     // vl[0], vl[1], eu[0], eu[1], flags, self-gravity exchange block — usable directly inside one process
    cudaIpcMemHandle_t ipc[6];
    int32_t device;
    int32_t pad;
};

namespace {

template <typename T>
int dev_alloc(odis_solver* s, T** p, size_t count) {
    if (count == 0) count = 1;
    ODIS_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
    s->device_bytes += count * sizeof(T);
    return ODIS_OK;
}
template <typename T>
int upload(odis_solver* s, T** p, const std::vector<T>& h) {
    int rc = dev_alloc(s, p, h.size());
    if (rc) return rc;
    if (!h.empty()) ODIS_CUDA(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    return ODIS_OK;
}

// Host-evaluated time factors for the potential at `time` (tidalPotentials.cpp:55-61).
odis::StepScalars step_scalars(const odis_solver* s, double time) {
    odis::StepScalars m;
    const double omega = s->prm.omega;
    if (s->prm.potential == odis::P_PLANET) {
        // PLANET (tidalPotentials.cpp:176-213): a companion on the inner 2:1 orbit with the reference's hard-wired mass and orbit (Io);
        // the kernel gets cosphi, sinphi, factor, p in the slots cosM, sinM, cos2M, sin2M
        const double m2 = 8.931938e+22, a1 = 421800000.0, a2 = s->prm.semimajor_axis;
        const double n2 = omega, n1 = n2 * 2.0, nij = (n1 - n2);
        const double cosnt = std::cos(nij * time), sinnt = std::sin(nij * time);
        const double pp = std::pow(a1, 2.0) + std::pow(a2, 2.0) - 2. * a1 * a2 * cosnt;
        m.cosM = (a1 - a2 * cosnt);
        m.sinM = a2 * sinnt;
        m.cos2M = 0.5 * 6.67408e-11 * m2 * std::pow(s->forcing_radius / pp, 2.0) / std::sqrt(pp);
        m.sin2M = pp;
        m.cos3M = m.cos4M = 0.0;
        return m;
    }
    m.cosM = std::cos(omega * time);
    m.sinM = std::sin(omega * time);
    m.cos2M = std::cos(2 * omega * time);
    m.sin2M = std::sin(2 * omega * time);
    m.cos3M = std::cos(3 * omega * time);
    m.cos4M = std::cos(4 * omega * time);
    return m;
}

// PLANET forcing: the step kernels leave U = 0 for this type; one more pass writes it (before any self-gravity term is added)
void enqueue_planet(odis_solver* s, double2* eu, int n_cells, const odis::StepScalars& host, const odis::StepScalars* dev) {
    if (s->prm.potential != odis::P_PLANET) return;
    odis::launch_planet_potential(s->cell_tables(n_cells), host, dev, eu, n_cells, s->stream);
}
int planet_launches(const odis_solver* s) { return s->prm.potential == odis::P_PLANET ? 1 : 0; }

int ab3_mode(const odis_solver* s, int64_t iter) {
    if (iter > 1 || s->prm.init_load) return odis::AB3_FULL;      // temporalOperators.cpp:36
    return iter == 0 ? odis::AB3_FIRST : odis::AB3_SECOND;
}

int ensure_series(odis_solver* s, size_t need) {
    if (need <= s->series_cap) return ODIS_OK;
    size_t cap = s->series_cap ? s->series_cap : 4096;
    while (cap < need) cap *= 2;
    double* nd = nullptr;
    odis::StepScalars* ns = nullptr;
    ODIS_CUDA(cudaMalloc((void**)&nd, cap * sizeof(double)));
    if (cudaMalloc((void**)&ns, cap * sizeof(odis::StepScalars)) != cudaSuccess) {
        cudaFree(nd);
        cudaGetLastError();
        return fail(ODIS_ERR_CUDA, "cudaMalloc of the time-factor table failed");
    }
    ODIS_CUDA(cudaMemsetAsync(nd, 0, cap * sizeof(double), s->stream));
    ODIS_CUDA(cudaMemsetAsync(ns, 0, cap * sizeof(odis::StepScalars), s->stream));
    if (s->d_series) {
        ODIS_CUDA(cudaMemcpyAsync(nd, s->d_series, s->series_cap * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        ODIS_CUDA(cudaMemcpyAsync(ns, s->d_scal, s->series_cap * sizeof(odis::StepScalars), cudaMemcpyDeviceToDevice, s->stream));
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
        cudaFree(s->d_series);
        cudaFree(s->d_scal);
        s->device_bytes -= s->series_cap * (sizeof(double) + sizeof(odis::StepScalars));
        // captured graphs hold the old table addresses
        for (auto& g : s->graphs) cudaGraphExecDestroy(g.second);
        s->graphs.clear();
    }
    s->device_bytes += cap * (sizeof(double) + sizeof(odis::StepScalars));
    s->d_series = nd;
    s->d_scal = ns;
    s->series_cap = cap;
    return ODIS_OK;
}

int halo_drain(odis_solver* s);

// U += g * sum_{l >= 2} factor_l * (least-squares harmonic coefficients of eta) * Y_lm on {eta,U} buffer `eu`
// (pressureGradientSH, spatialOperators.cpp:387-462)
int enqueue_self_gravity(odis_solver* s, double2* eu) {
    if (!s->sh_on) return ODIS_OK;
    const odis::ShTables t = s->sh_tables();
    const odis::ShWork w = s->sh_work();
    odis::ShExchange x;
    if (s->world > 1) x = s->sh_exchange();
    odis::launch_sh_analysis(t, w, eu, s->No, s->prm.g, s->world > 1 ? &x : nullptr, s->stream);
    odis::launch_sh_synthesis(t, w, eu, s->N, s->stream);
    return ODIS_OK;
}

int run_diagnostics(odis_solver* s, bool want_fields) {
    if (s->diag_current && !want_fields) return ODIS_OK;
    int rc = ensure_series(s, (size_t)(s->iter - s->iter0) + 1);
    if (rc) return rc;
    if ((rc = halo_drain(s))) return rc;
    odis::launch_edge_diagnostics(s->edge_tables(), s->phys, s->d_vl[s->cur], s->d_normal, want_fields ? s->d_vavg : nullptr,
                                  want_fields ? s->d_ediss : nullptr, s->d_block_partial, s->d_ticket,
                                  s->d_series + (s->iter - s->iter0), s->prm.block_threads, s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    s->diag_current = true;
    return ODIS_OK;
}

int create_impl(const odis_mesh_view* mv, const odis_params* prm, int32_t device, int32_t rank, int32_t world, odis_solver** out) {
    if (!mv || !prm || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (mv->n_cells < 12 || mv->n_edges != 3 * mv->n_cells - 6) return fail(ODIS_ERR_ARG, "mesh sizes are inconsistent (F != 3N-6)");
    if (!(prm->dt > 0.0) || !(prm->radius > 0.0)) return fail(ODIS_ERR_ARG, "dt and radius must be positive");
    if (world < 1 || rank < 0 || rank >= world) return fail(ODIS_ERR_ARG, "rank/world out of range");
    if (world > mv->n_cells / 16) return fail(ODIS_ERR_ARG, "too many ranks for this grid");
    switch (prm->potential) {
        case odis::P_OBLIQ: case odis::P_OBLIQ_WEST: case odis::P_ECC: case odis::P_FULL: case odis::P_FULL2: case odis::P_NONE: break;
        case odis::P_PLANET:
            if (!(prm->semimajor_axis > 0.0)) return fail(ODIS_ERR_ARG, "potential PLANET needs the semimajor axis (globals->a)");
            break;
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
    s->rank = rank; s->world = world;
    s->Ng = mv->n_cells;
    s->Fg = mv->n_edges;
    s->prm = *prm;
    const int Ng = s->Ng, Fg = s->Fg;
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

    // ---- global renumbering and, for world > 1, the partition with its halo ----
    odis::LocalNumbering num;
    odis::build_local_numbering(Ng, Fg, mv->node_pos_sph, mv->face_nodes, mv->faces, prm->reorder != 0, rank, world, num);
    if ((int)num.part.peers.size() > kMaxPeers) return bail(fail(ODIS_ERR_UNSUPPORTED, "more than 8 neighbouring ranks"));
    s->part = num.part;
    s->cell_perm = num.cell_perm;
    s->edge_perm = num.edge_perm;
    s->N = (int)s->cell_perm.size(); s->F = (int)s->edge_perm.size();
    s->No = s->part.n_own_cells; s->Fo = s->part.n_own_edges;
    const int tile = odis::pipe_tile();
    s->Np = (s->N + tile - 1) / tile * tile;
    s->Fp = (s->Fo + tile - 1) / tile * tile;
    // kernel selection, odis_params.reserved[0] (include/odis_b200.h): bit 0 = the direct-load baseline kernels (edge, cell, separate
    // self-gravity launches, 6-launch nonlinear step); bit 3 = no CUDA-graph replay; bit 7 = 32-bit stencil ids only (default: 16-bit
    // offsets where they fit); bit 8 = test hook of the narrow ids (+-1023 range)
    s->pipe_edge = (prm->reserved[0] & 1) == 0;
    s->pipe_cell = s->pipe_edge;
    s->use_graph = (prm->reserved[0] & 8) == 0;
    s->nl_folded = s->pipe_edge;
    s->edge_ids16 = (prm->reserved[0] & 128) == 0 && s->pipe_edge;
    const int N = s->N, F = s->F, No = s->No, Fo = s->Fo, Np = s->Np, Fp = s->Fp;
    const int Fvl = (F + tile - 1) / tile * tile;       // {v,l} arrays: every local edge, padded to whole tiles
    auto local_cell = [&](int old_id) { return num.local_cell_of_ref(old_id); };
    auto local_edge = [&](int old_id) { return num.local_edge_of_ref(old_id); };

    // ---- edge tables (owned edges; SoA stride Fo) and the {v,l} arrays (all local edges) ----
    {
        std::vector<int2> cells((size_t)Fp, make_int2(0, 0));
        std::vector<double2> grad((size_t)Fp, make_double2(0.0, 0.0)), normal((size_t)Fo), vl((size_t)Fvl, make_double2(0.0, 0.0));
        std::vector<double> fcor((size_t)Fp, 0.0), dist((size_t)Fp, 1.0), sw((size_t)Fp * odis::kStencil, 0.0);
        std::vector<int> sid((size_t)Fp * odis::kStencil, -1);
        int bad = 0;
#pragma omp parallel for schedule(static)
        for (int en = 0; en < F; en++) vl[en] = make_double2(0.0, mv->face_len[s->edge_perm[en]]);
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int en = 0; en < Fo; en++) {
            const int eo = s->edge_perm[en];
            const int c0 = mv->face_nodes[(size_t)eo * 2], c1 = mv->face_nodes[(size_t)eo * 2 + 1];
            cells[en] = make_int2(local_cell(c0), local_cell(c1));
            if (cells[en].x < 0 || cells[en].y < 0) bad++;
            const double d = mv->face_node_dist[eo];
            grad[en] = make_double2((-mv->face_centre_m[(size_t)eo * 2]) / d, (mv->face_centre_m[(size_t)eo * 2 + 1]) / d);   // mesh.cpp:3076-3080
            fcor[en] = -2.0 * prm->omega * std::sin(mv->face_centre_pos_sph[(size_t)eo * 2]);                                   // mesh.cpp:2881
            dist[en] = d;
            normal[en] = make_double2(mv->face_normal_vec_map[(size_t)eo * 2], mv->face_normal_vec_map[(size_t)eo * 2 + 1]);
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
                if (ids[j] < 0 || ids[j] >= Fg) { bad++; continue; }
                const int le = local_edge(ids[j]);
                if (le < 0) { bad++; continue; }
                sid[(size_t)j * Fp + en] = le;
                sw[(size_t)j * Fp + en] = ws[j];
            }
        }
        if (bad) return bail(fail(ODIS_ERR_ARG, "face_interp_friends / face_nodes hold out-of-range ids (or the halo is incomplete)"));
        if ((rc = upload(s, &s->d_cells, cells)) || (rc = upload(s, &s->d_grad, grad)) || (rc = upload(s, &s->d_fcor, fcor)) ||
            (rc = upload(s, &s->d_dist, dist)) || (rc = upload(s, &s->d_sid, sid)) || (rc = upload(s, &s->d_sw, sw)) ||
            (rc = upload(s, &s->d_normal, normal)) || (rc = upload(s, &s->d_vl[0], vl)) || (rc = upload(s, &s->d_vl[1], vl)))
            return bail(rc);
        if (s->edge_ids16) {
            // narrow ids: offsets from the edge's own id, tile-major; a tile with an offset beyond 16 bits stays on the int rows.
            // reserved[0] bit 8 (tests): pretend the range is +-1023 so that small grids have wide tiles too
            const int T = odis::pipe_tile(), n_tiles = Fp / T, lim = (prm->reserved[0] & 256) ? 1023 : 32767;
            std::vector<short> sid16((size_t)n_tiles * odis::kStencil * T, (short)0);
            std::vector<unsigned char> wide((size_t)(n_tiles + 15) / 16 * 16, (unsigned char)0);
            int n_wide = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_wide)
            for (int tile = 0; tile < n_tiles; tile++) {
                bool w = false;
                for (int j = 0; j < odis::kStencil; j++)
                    for (int k = 0; k < T; k++) {
                        const int en = tile * T + k, id = sid[(size_t)j * Fp + en];
                        const int off = id < 0 ? 0 : id - en;
                        if (off > lim || off < -lim) w = true;
                        sid16[((size_t)tile * odis::kStencil + j) * T + k] = (short)off;
                    }
                wide[(size_t)tile] = w ? 1 : 0;
                n_wide += w;
            }
            s->wide_tiles = n_wide;
            if (!odis::edge_ids16_fits(Fo)) s->edge_ids16 = false;      // more tiles per CTA than the flag buffer holds: stays on the int rows
            else if ((rc = upload(s, &s->d_sid16, sid16)) || (rc = upload(s, &s->d_tile_wide, wide))) return bail(rc);
        }
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "table upload failed"));
    }
    // ---- cell tables (SoA stride N = all local cells: ghosts need their potential at set_state) ----
    {
        std::vector<int> eid((size_t)Np * odis::kCellEdges, -1);
        std::vector<double> area((size_t)Np, 1.0), trig((size_t)Np * 8, 0.0), trig_sq((size_t)Np * 2, 0.0);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int cn = 0; cn < N; cn++) {
            const int co = s->cell_perm[cn];
            {
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
                    if (ids[j] < 0 || ids[j] >= Fg) { bad++; continue; }
                    const int le = local_edge(ids[j]);
                    if (le < 0) { bad++; continue; }
                    eid[(size_t)j * Np + cn] = le | (dirs[j] < 0 ? (int)0x80000000 : 0);
                }
            }
            area[cn] = mv->control_volume_surf_area_map[co];
            const double lat = mv->node_pos_sph[(size_t)co * 2], lon = mv->node_pos_sph[(size_t)co * 2 + 1];
            trig[0 * (size_t)Np + cn] = std::cos(lat);            // mesh.cpp:2132-2145
            trig[1 * (size_t)Np + cn] = std::sin(lat);
            trig[2 * (size_t)Np + cn] = std::cos(lon);
            trig[3 * (size_t)Np + cn] = std::sin(lon);
            trig[4 * (size_t)Np + cn] = std::cos(2.0 * lat);
            trig[5 * (size_t)Np + cn] = std::sin(2.0 * lat);
            trig[6 * (size_t)Np + cn] = std::cos(2.0 * lon);
            trig[7 * (size_t)Np + cn] = std::sin(2.0 * lon);
            trig_sq[cn] = std::cos(lat) * std::cos(lat);
            trig_sq[(size_t)Np + cn] = std::sin(lat) * std::sin(lat);
        }
        if (bad) return bail(fail(ODIS_ERR_ARG, "faces table holds out-of-range edge ids (or the halo is incomplete)"));
        if ((rc = upload(s, &s->d_eid, eid)) || (rc = upload(s, &s->d_area, area)) || (rc = upload(s, &s->d_trig, trig)) ||
            (rc = upload(s, &s->d_trig_sq, trig_sq)))
            return bail(rc);
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "table upload failed"));
    }
    // ---- state ----
    // one energy partial per warp of edges (>= blocks of edge_diagnostics); with block_threads 256 / 512 the direct edge kernel's last
    // block writes partials for warps beyond the 128-edge tile padding, hence the slack
    const int blocks = Fp / 32 + 16;
    if ((rc = dev_alloc(s, &s->d_eu[0], (size_t)Np)) || (rc = dev_alloc(s, &s->d_eu[1], (size_t)Np)) ||
        (rc = dev_alloc(s, &s->d_hv[0], (size_t)Fp)) || (rc = dev_alloc(s, &s->d_hv[1], (size_t)Fp)) ||
        (rc = dev_alloc(s, &s->d_he[0], (size_t)Np)) || (rc = dev_alloc(s, &s->d_he[1], (size_t)Np)) || (rc = dev_alloc(s, &s->d_he[2], (size_t)Np)) ||
        (rc = dev_alloc(s, &s->d_block_partial, (size_t)blocks)) || (rc = dev_alloc(s, &s->d_ticket, (size_t)1)) ||
        (rc = dev_alloc(s, &s->d_ctl, (size_t)1)) ||
        (rc = dev_alloc(s, &s->d_vavg, (size_t)Fo)) || (rc = dev_alloc(s, &s->d_ediss, (size_t)Fo)))
        return bail(rc);
    if ((rc = upload(s, &s->d_edge_perm, s->edge_perm)) || (rc = upload(s, &s->d_cell_perm, s->cell_perm)) ||
        (rc = dev_alloc(s, &s->d_stage, std::max((size_t)Fg * 3, (size_t)F + 3 * (size_t)Fo + 4 * (size_t)N))) || (rc = dev_alloc(s, &s->d_lvl0_v, (size_t)Fo)) ||
        (rc = dev_alloc(s, &s->d_lvl0_e, (size_t)Np)))
        return bail(rc);
    cudaMemsetAsync(s->d_ticket, 0, sizeof(unsigned int), s->stream);
    cudaMemsetAsync(s->d_ctl, 0, sizeof(odis::StepCtl), s->stream);
    {   // upper bound of every in-kernel wait for another rank (halo flags, harmonic all-reduce): ~10 s unless the environment says otherwise.
        // One host thread driving all ranks in turn must enqueue every rank's steps within this time of each other.
        const char* e = std::getenv("ODIS_B200_WAIT_TIMEOUT_S");
        const double sec = e ? std::atof(e) : 0.0;
        s->wait_cycles = sec > 0.0 ? (long long)(sec * 2.0e9) : 0ll;
        if (s->wait_cycles > 0)
            cudaMemcpyAsync(&s->d_ctl->spin_cycles, &s->wait_cycles, sizeof(long long), cudaMemcpyHostToDevice, s->stream);
    }
    if (cudaSuccess != odis::pipe_configure()) return bail(fail(ODIS_ERR_CUDA, "cudaFuncSetAttribute(shared memory size) failed"));
    cudaMemsetAsync(s->d_eu[0], 0, (size_t)Np * sizeof(double2), s->stream);
    cudaMemsetAsync(s->d_eu[1], 0, (size_t)Np * sizeof(double2), s->stream);
    cudaMemsetAsync(s->d_he[2], 0, (size_t)Np * sizeof(double), s->stream);
    cudaMemsetAsync(s->d_hv[0], 0, (size_t)Fp * sizeof(double), s->stream);
    cudaMemsetAsync(s->d_hv[1], 0, (size_t)Fp * sizeof(double), s->stream);
    cudaMemsetAsync(s->d_he[0], 0, (size_t)Np * sizeof(double), s->stream);
    cudaMemsetAsync(s->d_he[1], 0, (size_t)Np * sizeof(double), s->stream);
    // ---- halo send lists ----
    if (world > 1) {
        std::vector<int> el, er, ep, cl, cr, cp;
        s->n_peers = (int)s->part.peers.size();
        for (int k = 0; k < s->n_peers; k++) {
            const odis::HaloPeer& peer = s->part.peers[(size_t)k];
            s->peer_rank[k] = peer.rank;
            for (size_t i = 0; i < peer.send_edge_local.size(); i++) { el.push_back(peer.send_edge_local[i]); er.push_back(peer.send_edge_remote[i]); ep.push_back(k); }
            for (size_t i = 0; i < peer.send_cell_local.size(); i++) { cl.push_back(peer.send_cell_local[i]); cr.push_back(peer.send_cell_remote[i]); cp.push_back(k); }
        }
        s->n_send_e = (int)el.size(); s->n_send_c = (int)cl.size();
        if ((rc = upload(s, &s->d_send_e_local, el)) || (rc = upload(s, &s->d_send_e_remote, er)) || (rc = upload(s, &s->d_send_e_peer, ep)) ||
            (rc = upload(s, &s->d_send_c_local, cl)) || (rc = upload(s, &s->d_send_c_remote, cr)) || (rc = upload(s, &s->d_send_c_peer, cp)) ||
            (rc = dev_alloc(s, &s->d_flags, (size_t)2 * world)) || (rc = dev_alloc(s, &s->d_halo_ticket, (size_t)1)))
            return bail(rc);
        cudaMemsetAsync(s->d_flags, 0, (size_t)2 * world * sizeof(unsigned long long), s->stream);
        cudaMemsetAsync(s->d_halo_ticket, 0, sizeof(unsigned int), s->stream);
        // CSR form of the edge send list over the boundary edges (first in the own numbering, odis_partition.h). At
        // least one boundary tile, so that a rank with nothing to send still publishes its epoch.
        s->Fb = std::max(1, s->part.n_bnd_edges);
        s->wait_from = std::min(No - s->part.n_bnd_cells, N - 1);     // at least the last CTA waits
        std::vector<int> first((size_t)s->Fb + 1, 0), pk, rm;
        for (int k = 0; k < s->n_peers; k++)
            for (int l : s->part.peers[(size_t)k].send_edge_local) {
                if (l < 0 || l >= s->part.n_bnd_edges) return bail(fail(ODIS_ERR_STATE, "send list outside the boundary range"));
                first[(size_t)l + 1]++;
            }
        for (int i = 0; i < s->Fb; i++) first[(size_t)i + 1] += first[(size_t)i];
        pk.assign((size_t)first[(size_t)s->Fb], 0); rm.assign((size_t)first[(size_t)s->Fb], 0);
        {
            std::vector<int> fill(first.begin(), first.end() - 1);
            for (int k = 0; k < s->n_peers; k++) {
                const odis::HaloPeer& peer = s->part.peers[(size_t)k];
                for (size_t i = 0; i < peer.send_edge_local.size(); i++) {
                    const int at = fill[(size_t)peer.send_edge_local[i]]++;
                    pk[(size_t)at] = k; rm[(size_t)at] = peer.send_edge_remote[i];
                }
            }
        }
        if ((rc = upload(s, &s->d_csr_e_first, first)) || (rc = upload(s, &s->d_csr_e_peer, pk)) || (rc = upload(s, &s->d_csr_e_remote, rm)) ||
            (rc = dev_alloc(s, &s->d_halo_done, (size_t)1)) || (rc = dev_alloc(s, &s->d_sh_xblock, odis::kShXBytes)) ||
            (rc = dev_alloc(s, &s->d_sh_xctl, (size_t)4)))
            return bail(rc);
        if (world > odis::kShMaxWorld) return bail(fail(ODIS_ERR_ARG, "at most 8 ranks"));
        cudaMemsetAsync(s->d_halo_done, 0, sizeof(unsigned int), s->stream);
        cudaMemsetAsync(s->d_sh_xblock, 0, odis::kShXBytes, s->stream);
        cudaMemsetAsync(s->d_sh_xctl, 0, 4 * sizeof(unsigned long long), s->stream);
        if (s->wait_cycles > 0) cudaMemcpyAsync(s->d_sh_xctl + 3, &s->wait_cycles, sizeof(long long), cudaMemcpyHostToDevice, s->stream);
        s->sh_xremote[rank] = s->d_sh_xblock;
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(ODIS_ERR_CUDA, "halo table upload failed"));
    }
    *out = s;
    rc = odis_set_state(s, nullptr, nullptr, nullptr, nullptr, 0);
    if (rc) { *out = nullptr; return bail(rc); }
    return ODIS_OK;
}

// one halo exchange (one launch): push my boundary values into the peers' ghost slots, publish the epoch, wait for theirs
int halo_exchange(odis_solver* s, int kind /*0: edges {v,l}, 1: cells {eta,U}*/, const double2* src, int dst_buffer) {
    const odis::HaloRemote& rem = kind == 0 ? s->remote_v[dst_buffer] : s->remote_c[dst_buffer];
    odis::HaloWait w;
    w.n_peers = s->n_peers;
    for (int k = 0; k < s->n_peers; k++) w.flag[k] = s->d_flags + (size_t)kind * s->world + s->peer_rank[k];
    odis::launch_halo_exchange(kind == 0 ? s->n_send_e : s->n_send_c, kind == 0 ? s->d_send_e_local : s->d_send_c_local,
                               kind == 0 ? s->d_send_e_remote : s->d_send_c_remote, kind == 0 ? s->d_send_e_peer : s->d_send_c_peer, src, rem, w,
                               kind * s->world + s->rank, kind, s->d_ctl, s->d_halo_ticket, s->stream);
    s->launches += 1;
    ODIS_CUDA(cudaGetLastError());
    return ODIS_OK;
}

odis::HaloWait halo_wait_of(const odis_solver* s, int kind) {
    odis::HaloWait w;
    w.n_peers = s->n_peers;
    for (int k = 0; k < s->n_peers; k++) w.flag[k] = s->d_flags + (size_t)kind * s->world + s->peer_rank[k];
    return w;
}

// descriptors of the exchange fused into the step kernels (HaloInline): the edge kernel pushes, the cell kernel waits
odis::HaloInline halo_inline_edge(const odis_solver* s, int dst_buffer) {
    odis::HaloInline h;
    h.n_bnd = s->Fb;
    h.wait_from = 0x7fffffff;
    h.n_peers = s->n_peers;
    h.flag_slot = s->rank;
    h.send_first = s->d_csr_e_first; h.send_peer = s->d_csr_e_peer; h.send_remote = s->d_csr_e_remote;
    h.remote = s->remote_v[dst_buffer];
    h.wait_v = halo_wait_of(s, 0);
    h.done = s->d_halo_done;
    h.ctl = s->d_ctl;
    return h;
}
odis::HaloInline halo_inline_cell(const odis_solver* s) {
    odis::HaloInline h;
    h.n_bnd = 0;
    h.wait_from = s->wait_from;
    h.n_peers = s->n_peers;
    h.flag_slot = s->rank;
    h.send_first = h.send_peer = h.send_remote = nullptr;
    h.wait_v = halo_wait_of(s, 0);
    h.done = nullptr;
    h.ctl = s->d_ctl;
    return h;
}

// wait (bounded) until every push of the exchanges this rank took part in has landed: before the state is overwritten or
// read back, and in front of a cell kernel variant that does not wait itself
int halo_drain(odis_solver* s) {
    if (s->world <= 1 || !s->connected) return ODIS_OK;
    odis::launch_halo_drain(halo_wait_of(s, 0), s->d_ctl, s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    return ODIS_OK;
}

}  // namespace

extern "C" {

int odis_create(const odis_mesh_view* mv, const odis_params* prm, int32_t device, odis_solver** out) {
    return create_impl(mv, prm, device, 0, 1, out);
}

int odis_create_partitioned(const odis_mesh_view* mv, const odis_params* prm, int32_t device, int32_t rank, int32_t world,
                            odis_solver** out) {
    return create_impl(mv, prm, device, rank, world, out);
}

// reference ids of the entries a (partitioned) solver owns, in device order: the order of its compact snapshots
int odis_get_partition_map(odis_solver* s, int32_t* own_cell_ref_out, int32_t* own_edge_ref_out) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (own_cell_ref_out) std::memcpy(own_cell_ref_out, s->cell_perm.data(), (size_t)s->No * sizeof(int32_t));
    if (own_edge_ref_out) std::memcpy(own_edge_ref_out, s->edge_perm.data(), (size_t)s->Fo * sizeof(int32_t));
    return ODIS_OK;
}

int odis_halo_blob_size(void) { return (int)sizeof(HaloBlob); }

int odis_halo_export(odis_solver* s, void* blob_out) {
    if (!s || !blob_out) return fail(ODIS_ERR_ARG, "NULL argument");
    ODIS_CUDA(cudaSetDevice(s->device));
    HaloBlob b;
    std::memset(&b, 0, sizeof b);
    b.rank = s->rank; b.world = s->world; b.pid = (int64_t)getpid(); b.device = s->device;
    b.raw[0] = s->d_vl[0]; b.raw[1] = s->d_vl[1]; b.raw[2] = s->d_eu[0]; b.raw[3] = s->d_eu[1]; b.raw[4] = s->d_flags; b.raw[5] = s->d_sh_xblock;
    if (s->world > 1)
        for (int k = 0; k < 6; k++) ODIS_CUDA(cudaIpcGetMemHandle(&b.ipc[k], b.raw[k]));
    std::memcpy(blob_out, &b, sizeof b);
    return ODIS_OK;
}

int odis_halo_connect(odis_solver* s, const void* all_blobs) {
    if (!s || !all_blobs) return fail(ODIS_ERR_ARG, "NULL argument");
    if (s->world == 1) { s->connected = true; return ODIS_OK; }
    ODIS_CUDA(cudaSetDevice(s->device));
    const HaloBlob* blobs = (const HaloBlob*)all_blobs;
    for (int k = 0; k < s->n_peers; k++) {
        const HaloBlob& b = blobs[s->peer_rank[k]];
        if (b.rank != s->peer_rank[k] || b.world != s->world) return fail(ODIS_ERR_ARG, "halo blobs are not ordered by rank");
        void* p[5];
        if (b.pid == (int64_t)getpid()) {                       // same process: the raw pointers are usable after enabling peer access
            if (b.device != s->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ODIS_ERR_CUDA, "cudaDeviceEnablePeerAccess failed");
                cudaGetLastError();
            }
            for (int j = 0; j < 5; j++) p[j] = b.raw[j];
        } else {
            for (int j = 0; j < 5; j++) {
                ODIS_CUDA(cudaIpcOpenMemHandle(&p[j], b.ipc[j], cudaIpcMemLazyEnablePeerAccess));
                s->ipc_opened[k][j] = p[j];
            }
        }
        s->remote_v[0].data[k] = (double2*)p[0]; s->remote_v[1].data[k] = (double2*)p[1];
        s->remote_c[0].data[k] = (double2*)p[2]; s->remote_c[1].data[k] = (double2*)p[3];
        s->remote_v[0].flags[k] = s->remote_v[1].flags[k] = s->remote_c[0].flags[k] = s->remote_c[1].flags[k] = (unsigned long long*)p[4];
    }
    // the self-gravity exchange blocks of ALL ranks (the harmonic sums are global, not a neighbour exchange)
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) continue;
        const HaloBlob& b = blobs[r];
        if (b.rank != r || b.world != s->world) return fail(ODIS_ERR_ARG, "halo blobs are not ordered by rank");
        if (b.pid == (int64_t)getpid()) {
            if (b.device != s->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ODIS_ERR_CUDA, "cudaDeviceEnablePeerAccess failed");
                cudaGetLastError();
            }
            s->sh_xremote[r] = (unsigned char*)b.raw[5];
        } else {
            void* p = nullptr;
            ODIS_CUDA(cudaIpcOpenMemHandle(&p, b.ipc[5], cudaIpcMemLazyEnablePeerAccess));
            s->sh_ipc_opened[r] = p;
            s->sh_xremote[r] = (unsigned char*)p;
        }
    }
    s->connected = true;
    return ODIS_OK;
}

int odis_get_partition(odis_solver* s, int32_t* rank, int32_t* world, int32_t* own_cells, int32_t* own_edges, int32_t* ghost_cells,
                       int32_t* ghost_edges, int32_t* n_peers) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    if (rank) *rank = s->rank;
    if (world) *world = s->world;
    if (own_cells) *own_cells = s->No;
    if (own_edges) *own_edges = s->Fo;
    if (ghost_cells) *ghost_cells = s->N - s->No;
    if (ghost_edges) *ghost_edges = s->F - s->Fo;
    if (n_peers) *n_peers = s->n_peers;
    return ODIS_OK;
}

static int finish_set_state(odis_solver* s, int64_t iter, bool sync);

// host threads for packing / spreading a partitioned solver's entries: the launcher of multi-process runs (torchrun) sets
// OMP_NUM_THREADS=1, which would leave these memory-bound loops on one core; every rank takes its share of the cores instead
static int pack_threads(const odis_solver* s) {
    const unsigned hw = std::thread::hardware_concurrency();
    const int share = (int)(hw ? hw : 8u) / (s->world > 0 ? s->world : 1);
    return std::max(1, std::min(16, share));
}

// page-locked staging of a partitioned solver (at least `doubles` entries); waits for copies still reading the previous content
static int ensure_pack(odis_solver* s, size_t doubles) {
    if (s->pack_pending) {
        ODIS_CUDA(cudaEventSynchronize(s->pack_done));
        s->pack_pending = false;
    }
    if (doubles <= s->h_pack_cap) return ODIS_OK;
    if (s->h_pack) cudaFreeHost(s->h_pack);
    s->h_pack = nullptr; s->h_pack_cap = 0;
    ODIS_CUDA(cudaHostAlloc((void**)&s->h_pack, doubles * sizeof(double), cudaHostAllocDefault));
    s->h_pack_cap = doubles;
    if (!s->pack_done) ODIS_CUDA(cudaEventCreateWithFlags(&s->pack_done, cudaEventDisableTiming));
    return ODIS_OK;
}

// A partitioned solver's share of the caller's GLOBAL (reference-ordered) state arrays, packed in device order into page-locked
// memory: [v F | dvdt 3 Fo | eta N | detadt 3 N] (own + halo entries; the velocity history of the own edges only)
static void pack_partitioned(const odis_solver* s, double* dst, const double* v, const double* eta, const double* dvdt, const double* detadt) {
    const size_t N = (size_t)s->N, F = (size_t)s->F, Fo = (size_t)s->Fo;
    double* hv = dst; double* hdv = hv + F; double* he = hdv + 3 * Fo; double* hde = he + N;
    const int* ep = s->edge_perm.data(); const int* cp = s->cell_perm.data();
    const int nt = pack_threads(s);
    if (v) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)F; i++) hv[i] = v[ep[i]];
    }
    if (dvdt) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)Fo; i++) { const size_t o = (size_t)ep[i] * 3; hdv[3 * i] = dvdt[o]; hdv[3 * i + 1] = dvdt[o + 1]; hdv[3 * i + 2] = dvdt[o + 2]; }
    }
    if (eta) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)N; i++) he[i] = eta[cp[i]];
    }
    if (detadt) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)N; i++) { const size_t o = (size_t)cp[i] * 3; hde[3 * i] = detadt[o]; hde[3 * i + 1] = detadt[o + 1]; hde[3 * i + 2] = detadt[o + 2]; }
    }
}

// the four packed arrays, host -> device side by side at `d` (layout of pack_partitioned); mask bit k: array k was given
static int upload_partitioned(odis_solver* s, double* d, const double* h, unsigned mask, cudaStream_t stream) {
    const size_t N = (size_t)s->N, F = (size_t)s->F, Fo = (size_t)s->Fo;
    const size_t off[4] = {0, F, F + 3 * Fo, F + 3 * Fo + N}, cnt[4] = {F, 3 * Fo, N, 3 * N};
    for (int k = 0; k < 4; k++)
        if (mask & (1u << k)) ODIS_CUDA(cudaMemcpyAsync(d + off[k], h + off[k], cnt[k] * sizeof(double), cudaMemcpyHostToDevice, stream));
    return ODIS_OK;
}

// ... and from there into the state (already in device order: no permutation)
static void scatter_partitioned(odis_solver* s, const double* d, unsigned mask) {
    const size_t N = (size_t)s->N, F = (size_t)s->F, Fo = (size_t)s->Fo;
    const double* dv = d; const double* ddv = dv + F; const double* de = ddv + 3 * Fo; const double* dde = de + N;
    odis::launch_scatter_x(s->F, nullptr, (mask & 1u) ? dv : nullptr, s->d_vl[s->cur], 0, s->stream);
    s->hv1 = 0;
    odis::launch_scatter_history(s->Fo, nullptr, (mask & 2u) ? ddv : nullptr, s->d_lvl0_v, s->d_hv[0], s->d_hv[1], s->stream);
    odis::launch_scatter_x(s->N, nullptr, (mask & 4u) ? de : nullptr, s->d_eu[s->ecur], 1, s->stream);
    s->he1 = 0; s->he2 = 1; s->hefree = 2;
    odis::launch_scatter_history(s->N, nullptr, (mask & 8u) ? dde : nullptr, s->d_lvl0_e, s->d_he[0], s->d_he[1], s->stream);
}

// odis_set_state of a partitioned solver: the rank packs the entries it holds (own + halo) out of the caller's global arrays into
// page-locked memory, in device order, and uploads only those — 1/world of the state instead of all of it per rank
static int set_state_partitioned(odis_solver* s, const double* v, const double* eta, const double* dvdt, const double* detadt) {
    const size_t N = (size_t)s->N, F = (size_t)s->F, Fo = (size_t)s->Fo;
    int rc = ensure_pack(s, F + 3 * Fo + 4 * N);
    if (rc) return rc;
    pack_partitioned(s, s->h_pack, v, eta, dvdt, detadt);
    const unsigned mask = (v ? 1u : 0u) | (dvdt ? 2u : 0u) | (eta ? 4u : 0u) | (detadt ? 8u : 0u);
    // d_stage holds max(3 Fg, F + 3 Fo + 4 N) doubles (create_impl): the four packed arrays side by side
    if ((rc = upload_partitioned(s, s->d_stage, s->h_pack, mask, s->stream))) return rc;
    ODIS_CUDA(cudaEventRecord(s->pack_done, s->stream));
    s->pack_pending = true;
    scatter_partitioned(s, s->d_stage, mask);
    return ODIS_OK;
}

int odis_set_state(odis_solver* s, const double* v, const double* eta, const double* dvdt, const double* detadt, int64_t iter) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (iter < 0) return fail(ODIS_ERR_ARG, "iter must be >= 0");
    ODIS_CUDA(cudaSetDevice(s->device));
    if (s->world > 1) {
        int rcp = halo_drain(s);
        if (rcp || (rcp = set_state_partitioned(s, v, eta, dvdt, detadt))) return rcp;
        return finish_set_state(s, iter, true);
    }
    const int N = s->N, F = s->F, No = s->No, Fo = s->Fo;
    // Reference-ordered (global) host arrays go through one device staging buffer and are renumbered by
    // small kernels; velocities keep the static edge length beside them (only .x is rewritten). A
    // partitioned solver takes the same global arrays and keeps its own cells/edges plus its halo.
    auto stage = [&](const double* host, size_t n) -> int {
        if (!host) return ODIS_OK;
        ODIS_CUDA(cudaMemcpyAsync(s->d_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        return ODIS_OK;
    };
    int rc;
    if ((rc = halo_drain(s))) return rc;
    if ((rc = stage(v, (size_t)s->Fg))) return rc;
    odis::launch_scatter_x(F, s->d_edge_perm, v ? s->d_stage : nullptr, s->d_vl[s->cur], 0, s->stream);
    if ((rc = stage(dvdt, (size_t)s->Fg * 3))) return rc;
    s->hv1 = 0;
    odis::launch_scatter_history(Fo, s->d_edge_perm, dvdt ? s->d_stage : nullptr, s->d_lvl0_v, s->d_hv[0], s->d_hv[1], s->stream);
    if ((rc = stage(eta, (size_t)s->Ng))) return rc;
    odis::launch_scatter_x(N, s->d_cell_perm, eta ? s->d_stage : nullptr, s->d_eu[s->ecur], 1, s->stream);
    if ((rc = stage(detadt, (size_t)s->Ng * 3))) return rc;
    s->he1 = 0; s->he2 = 1; s->hefree = 2;
    // the history of every held cell is kept (halo cells are updated locally by the fused kernel)
    odis::launch_scatter_history(N, s->d_cell_perm, detadt ? s->d_stage : nullptr, s->d_lvl0_e, s->d_he[0], s->d_he[1], s->stream);
    return finish_set_state(s, iter, true);
}

// after the four scatter launches of odis_set_state / odis_commit_state: counters, the potential of the first step
static int finish_set_state(odis_solver* s, int64_t iter, bool sync) {
    const int N = s->N;
    int rc;
    s->launches += 4;
    ODIS_CUDA(cudaMemsetAsync(&s->d_ctl->count, 0, sizeof(unsigned long long), s->stream));
    s->iter = iter;
    s->iter0 = iter;
    s->last_mode = -1;
    s->diag_current = false;
    s->have_state = true;
    // potential for the first step: forcing(current_time + dt), timeIntegrator.cpp:187,218 — for every
    // held cell, halo included
    const double t = s->prm.dt * (double)iter + s->prm.dt;
    odis::CellState cs{s->d_vl[s->cur], s->d_eu[s->ecur], s->d_eu[s->ecur], s->d_he[0], s->d_he[1], s->d_he[2], nullptr, 0, nullptr, nullptr};
    odis::launch_cell_step(s->cell_tables(N), s->phys, cs, odis::AB3_FULL, step_scalars(s, t), odis::CELL_UPDATE_U, s->prm.block_threads,
                           nullptr, s->stream);
    enqueue_planet(s, s->d_eu[s->ecur], N, step_scalars(s, t), nullptr);
    s->launches += 1 + planet_launches(s);
    if ((rc = enqueue_self_gravity(s, s->d_eu[s->ecur]))) return rc;
    s->launches += s->sh_launches();
    ODIS_CUDA(cudaGetLastError());
    // partitioned + self-gravity: the harmonic sums wait for every rank's share, so the call must not block here (one host
    // thread may be driving all ranks in turn); the stream keeps the order
    if (sync && !(s->world > 1 && s->sh_on)) ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

// ---- the next interval's state staged while the current one runs ------------------------------------------------------
// odis_stage_state: host -> device copies of the (reference-ordered) arrays on the copy stream into a staging area of their own;
// returns at once, the caller keeps stepping. The host arrays must stay unchanged until odis_commit_state (page-locked memory makes
// the copies truly asynchronous). odis_commit_state: the solver's stream waits for the copies, renumbers the staged arrays into the
// state (the four scatter launches of odis_set_state) and evaluates the first potential; it does not wait on the host. A further
// odis_stage_state waits (on the device) until the commit has consumed the staging area.
int odis_stage_state(odis_solver* s, const double* v, const double* eta, const double* dvdt, const double* detadt) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (s->stage_pending) return fail(ODIS_ERR_STATE, "a staged state is waiting for odis_commit_state");
    ODIS_CUDA(cudaSetDevice(s->device));
    const bool part = s->world > 1;
    // unpartitioned: the reference-ordered arrays as they are; partitioned: this rank's share, packed in device order
    const size_t F = part ? (size_t)s->F : (size_t)s->Fg, N = part ? (size_t)s->N : (size_t)s->Ng, Fh = part ? (size_t)s->Fo : F;
    const size_t total = F + 3 * Fh + 4 * N;
    if (!s->copy_stream) ODIS_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    if (!s->d_stage_next) {
        int rc = dev_alloc(s, &s->d_stage_next, total);
        if (rc) return rc;
        ODIS_CUDA(cudaEventCreateWithFlags(&s->staged, cudaEventDisableTiming));
        ODIS_CUDA(cudaEventCreateWithFlags(&s->consumed, cudaEventDisableTiming));
    }
    if (part) {
        // the caller's global arrays are read HERE (packed on the host while the device steps), so they are free again on return;
        // the previous staged copy has left the page-locked buffer before it is overwritten
        if (!s->h_stage_pack) ODIS_CUDA(cudaHostAlloc((void**)&s->h_stage_pack, total * sizeof(double), cudaHostAllocDefault));
        if (s->stage_used) ODIS_CUDA(cudaEventSynchronize(s->staged));
        pack_partitioned(s, s->h_stage_pack, v, eta, dvdt, detadt);
    }
    if (s->stage_used) ODIS_CUDA(cudaStreamWaitEvent(s->copy_stream, s->consumed, 0));     // the previous commit's scatter launches have read it
    s->stage_mask = (v ? 1u : 0u) | (dvdt ? 2u : 0u) | (eta ? 4u : 0u) | (detadt ? 8u : 0u);
    if (part) {
        int rc = upload_partitioned(s, s->d_stage_next, s->h_stage_pack, s->stage_mask, s->copy_stream);
        if (rc) return rc;
    } else {
        double* d = s->d_stage_next;
        const double* host[4] = {v, dvdt, eta, detadt};
        const size_t count[4] = {F, 3 * F, N, 3 * N};
        for (int k = 0; k < 4; k++) {
            if (host[k]) ODIS_CUDA(cudaMemcpyAsync(d, host[k], count[k] * sizeof(double), cudaMemcpyHostToDevice, s->copy_stream));
            d += count[k];
        }
    }
    ODIS_CUDA(cudaEventRecord(s->staged, s->copy_stream));
    s->stage_pending = true;
    return ODIS_OK;
}

int odis_commit_state(odis_solver* s, int64_t iter) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (iter < 0) return fail(ODIS_ERR_ARG, "iter must be >= 0");
    if (!s->stage_pending) return fail(ODIS_ERR_STATE, "odis_stage_state has not been called");
    ODIS_CUDA(cudaSetDevice(s->device));
    const unsigned m = s->stage_mask;
    if (s->world > 1) {
        int rcp = halo_drain(s);            // (a launch, no host wait) the neighbours' last pushes have landed before the state is overwritten
        if (rcp) return rcp;
    }
    ODIS_CUDA(cudaStreamWaitEvent(s->stream, s->staged, 0));
    if (s->world > 1) scatter_partitioned(s, s->d_stage_next, m);
    else {
        const size_t F = (size_t)s->Fg, N = (size_t)s->Ng;
        const double* d_v = s->d_stage_next; const double* d_dv = d_v + F; const double* d_eta = d_dv + 3 * F; const double* d_de = d_eta + N;
        odis::launch_scatter_x(s->F, s->d_edge_perm, (m & 1u) ? d_v : nullptr, s->d_vl[s->cur], 0, s->stream);
        s->hv1 = 0;
        odis::launch_scatter_history(s->Fo, s->d_edge_perm, (m & 2u) ? d_dv : nullptr, s->d_lvl0_v, s->d_hv[0], s->d_hv[1], s->stream);
        odis::launch_scatter_x(s->N, s->d_cell_perm, (m & 4u) ? d_eta : nullptr, s->d_eu[s->ecur], 1, s->stream);
        s->he1 = 0; s->he2 = 1; s->hefree = 2;
        odis::launch_scatter_history(s->N, s->d_cell_perm, (m & 8u) ? d_de : nullptr, s->d_lvl0_e, s->d_he[0], s->d_he[1], s->stream);
    }
    ODIS_CUDA(cudaEventRecord(s->consumed, s->stream));
    s->stage_pending = false;
    s->stage_used = true;
    return finish_set_state(s, iter, false);
}

int odis_enable_self_gravity(odis_solver* s, const odis_mesh_view* mv, int32_t l_max, const double* factor, int32_t stored_basis) {
    if (!s || !mv || !factor) return fail(ODIS_ERR_ARG, "NULL argument");
    if (l_max < 2 || l_max > 31) return fail(ODIS_ERR_ARG, "sh degree must be in 2..31");
    if (mv->n_cells != s->Ng) return fail(ODIS_ERR_ARG, "mesh does not match the solver");
    if (s->world > 1 && !s->connected) return fail(ODIS_ERR_STATE, "call odis_halo_connect before odis_enable_self_gravity on a partitioned solver");
    if (s->world > 1 && !s->pipe_edge) return fail(ODIS_ERR_UNSUPPORTED, "partitioned self-gravity needs the default step kernels");
    if (s->sh_on) return fail(ODIS_ERR_STATE, "self-gravity is already enabled");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int rows = odis::sh_rows(l_max);
    if (rows >= s->Ng) return fail(ODIS_ERR_ARG, "sh degree too high for this grid");
    // normal-matrix inverse of the least-squares fit over ALL cells, in reference order (the same on every rank)
    {
        std::vector<double> Yg((size_t)rows * s->Ng);
        odis::sh_basis(s->Ng, mv->node_pos_sph, l_max, (size_t)s->Ng, Yg.data());
        if (odis::sh_normal_inverse(rows, s->Ng, (size_t)s->Ng, Yg.data(), 0, s->sh_ginv_host) != 0)
            return fail(ODIS_ERR_ARG, "spherical-harmonic normal matrix is not positive definite");
    }
    std::vector<double> fac((size_t)rows, 0.0), rec((size_t)odis::kShRecDoubles);
    for (int k = odis::kShSkipRows; k < rows; k++) fac[(size_t)k] = factor[odis::sh_row_degree(k)];
    odis::sh_recurrence_table(rec.data());
    s->sh_lmax = l_max; s->sh_rows = rows; s->sh_stored = stored_basis != 0;
    int rc;
    if (s->sh_stored) {
        // basis rows of the held cells in device order; padded cells keep 0
        std::vector<double> pos((size_t)s->N * 2), Y((size_t)rows * s->Np, 0.0);
        for (int i = 0; i < s->N; i++) {
            pos[2 * (size_t)i] = mv->node_pos_sph[2 * (size_t)s->cell_perm[(size_t)i]];
            pos[2 * (size_t)i + 1] = mv->node_pos_sph[2 * (size_t)s->cell_perm[(size_t)i] + 1];
        }
        odis::sh_basis(s->N, pos.data(), l_max, (size_t)s->Np, Y.data());
        if ((rc = upload(s, &s->d_shY, Y))) return rc;
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
    }
    ODIS_CUDA(odis::sh_configure());
    if ((rc = upload(s, &s->d_shRec, rec)) || (rc = upload(s, &s->d_shGinv, s->sh_ginv_host)) || (rc = upload(s, &s->d_shFactor, fac)) ||
        (rc = dev_alloc(s, &s->d_sh_partial, (size_t)odis::kShMaxBlocks * rows)) || (rc = dev_alloc(s, &s->d_sh_b, (size_t)rows)) ||
        (rc = dev_alloc(s, &s->d_sh_s, (size_t)rows)))
        return rc;
    ODIS_CUDA(cudaMemsetAsync(s->d_sh_b, 0, (size_t)rows * sizeof(double), s->stream));
    ODIS_CUDA(cudaMemsetAsync(s->d_sh_s, 0, (size_t)rows * sizeof(double), s->stream));
    if (s->pipe_cell && !s->nl_on && !s->sh_stored && odis::cell_pipe_supports_sg(l_max)) {
        // degrees 2..4, matrix-free: the staged cell update accumulates the harmonic sums on the way (3 launches per step)
        s->sg_cta_stride = (odis::cell_pipe_grid(s->sg_cells()) + 31) / 32 * 32;
        if ((rc = dev_alloc(s, &s->d_sg_cta, (size_t)rows * s->sg_cta_stride)) || (rc = dev_alloc(s, &s->d_sg_ticket, (size_t)32))) return rc;
        ODIS_CUDA(cudaMemsetAsync(s->d_sg_cta, 0, (size_t)rows * s->sg_cta_stride * sizeof(double), s->stream));
        ODIS_CUDA(cudaMemsetAsync(s->d_sg_ticket, 0, 32 * sizeof(unsigned int), s->stream));
        s->sh_fused = true;
        // merged solve + synthesis (grid-wide barrier, cooperative launch) on unpartitioned solvers; partitioned ones keep the separate
        // launch: measured faster there (655,362 cells on 8 GPUs: 48.9 vs 52.0 us per step), and no kernel of theirs then needs all of
        // its CTAs resident at once, so a collective of the caller's running beside the steps cannot close a wait cycle
        // (measured again after the all-reduce became LL lines, 2 and 4 GPUs: merged 59.3 / 39.8 us per step, separate 59.5 / 39.7)
        s->sh_merged = odis::cell_pipe_merged() && s->world == 1;
    }
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    for (auto& g : s->graphs) cudaGraphExecDestroy(g.second);     // captured without the extra launches
    s->graphs.clear();
    s->sh_on = true;
    // the potential of the pending step gets the term of the current eta (as odis_set_state does from now on)
    if (s->have_state) {
        if ((rc = enqueue_self_gravity(s, s->d_eu[s->ecur]))) return rc;
        s->launches += s->sh_launches();
        if (s->world == 1) ODIS_CUDA(cudaStreamSynchronize(s->stream));
    }
    return ODIS_OK;
}

// ELL form (device numbering) of a reference-numbered CSR operator: rows through row_map, columns through col_map, slots in the
// CSR's own (ascending reference column) order. row_of(r) gives the reference row of sub-row r; n_rows device rows of stride `stride`.
static int build_ell(odis_solver* s, const odis_csr_view& A, int n_rows, int stride, int row_mul, int row_off, const std::vector<int>& row_ref,
                     const std::vector<int>& col_map, odis::Ell* out) {
    int width = 0;
    for (int r = 0; r < n_rows; r++) {
        const int rr = row_ref[(size_t)r] * row_mul + row_off;
        width = std::max(width, A.indptr[rr + 1] - A.indptr[rr]);
    }
    if (width > 16) return fail(ODIS_ERR_ARG, "nonlinear operator has more than 16 entries in a row");
    std::vector<int> id((size_t)width * stride, -1);
    std::vector<double> w((size_t)width * stride, 0.0);
    for (int r = 0; r < n_rows; r++) {
        const int rr = row_ref[(size_t)r] * row_mul + row_off;
        int k = 0;
        for (int q = A.indptr[rr]; q < A.indptr[rr + 1]; q++, k++) {
            const int col = A.indices[q];
            if (col < 0 || col >= (int)col_map.size()) return fail(ODIS_ERR_ARG, "nonlinear operator column out of range");
            id[(size_t)k * stride + r] = col_map[(size_t)col];
            w[(size_t)k * stride + r] = A.data[q];
        }
    }
    int* did = nullptr; double* dw = nullptr;
    int rc;
    if ((rc = upload(s, &did, id)) || (rc = upload(s, &dw, w))) return rc;
    s->nl_owned.push_back(did); s->nl_owned.push_back(dw);
    out->width = width; out->stride = stride; out->id = did; out->w = dw;
    return ODIS_OK;
}

int odis_enable_advection(odis_solver* s, const odis_mesh_view* mv, const odis_nonlinear_view* nv) {
    if (!s || !mv || !nv) return fail(ODIS_ERR_ARG, "NULL argument");
    if (s->nl_on) return fail(ODIS_ERR_STATE, "the nonlinear branch is already enabled");
    if (s->world > 1) return fail(ODIS_ERR_UNSUPPORTED, "the nonlinear branch runs on a single, unpartitioned solver");
    if (s->sh_fused) return fail(ODIS_ERR_STATE, "call odis_enable_advection before odis_enable_self_gravity");
    if (mv->n_cells != s->Ng || mv->n_edges != s->Fg) return fail(ODIS_ERR_ARG, "mesh does not match the solver");
    const int N = s->N, F = s->F, V = mv->n_vertices;
    if (nv->curl.n_rows != V || nv->curl.n_cols != F || nv->rbf_interp.n_rows != 3 * N || nv->rbf_interp.n_cols != F ||
        nv->directional_second_deriv.n_rows != 2 * F || nv->directional_second_deriv.n_cols != N)
        return fail(ODIS_ERR_ARG, "nonlinear operators have the wrong shapes (curl VxF, rbf 3NxF, second derivative 2FxN)");
    if (!mv->vertex_nodes || !mv->vertex_R || !mv->face_vertexes || !nv->vertex_sinlat || !nv->vertex_area)
        return fail(ODIS_ERR_ARG, "vertex tables missing");
    ODIS_CUDA(cudaSetDevice(s->device));
    // reference id -> device id
    std::vector<int> cmap((size_t)N), emap((size_t)F), vmap((size_t)V), vref((size_t)V);
    for (int i = 0; i < N; i++) cmap[(size_t)s->cell_perm[(size_t)i]] = i;
    for (int e = 0; e < F; e++) emap[(size_t)s->edge_perm[(size_t)e]] = e;
    // vertices follow their lowest-numbered cell (locality of the gathers)
    {
        std::vector<std::pair<int, int>> key((size_t)V);
        for (int v = 0; v < V; v++) {
            int lo = N;
            for (int j = 0; j < 3; j++) lo = std::min(lo, cmap[(size_t)mv->vertex_nodes[(size_t)v * 3 + j]]);
            key[(size_t)v] = {lo, v};
        }
        std::sort(key.begin(), key.end());
        for (int k = 0; k < V; k++) { vref[(size_t)k] = key[(size_t)k].second; vmap[(size_t)key[(size_t)k].second] = k; }
    }
    const int Vs = (V + 31) / 32 * 32, Fp = s->Fp, Np = s->Np;
    odis::NlTables& t = s->nl;
    t.n_vertices = V; t.n_edges = F; t.n_cells = N; t.vstride = Vs; t.estride = Fp; t.cstride = Np; t.omega = s->prm.omega;
    int rc;
    if ((rc = build_ell(s, nv->curl, V, Vs, 1, 0, vref, emap, &t.curl))) return rc;
    for (int c = 0; c < 3; c++)
        if ((rc = build_ell(s, nv->rbf_interp, N, Np, 3, c, s->cell_perm, emap, &t.rbf[c]))) return rc;
    for (int c = 0; c < 2; c++)
        if ((rc = build_ell(s, nv->directional_second_deriv, F, Fp, 2, c, s->edge_perm, cmap, &t.d2[c]))) return rc;
    std::vector<int> vnode((size_t)3 * Vs, 0), nid((size_t)odis::kStencil * Fp, -1);
    std::vector<double> vR((size_t)3 * Vs, 0.0), vsin((size_t)Vs, 0.0), varea_r((size_t)Vs, 1.0), ncoef((size_t)odis::kStencil * Fp, 0.0);
    std::vector<int2> fvert((size_t)Fp, make_int2(0, 0));
    for (int k = 0; k < V; k++) {
        const int v = vref[(size_t)k];
        for (int j = 0; j < 3; j++) {
            vnode[(size_t)j * Vs + k] = cmap[(size_t)mv->vertex_nodes[(size_t)v * 3 + j]];
            vR[(size_t)j * Vs + k] = mv->vertex_R[(size_t)v * 3 + j];
        }
        vsin[(size_t)k] = nv->vertex_sinlat[v];
        varea_r[(size_t)k] = 1.0 / nv->vertex_area[v];                              // mesh.cpp:1147
    }
    for (int e = 0; e < F; e++) {
        const int er = s->edge_perm[(size_t)e];
        fvert[(size_t)e] = make_int2(vmap[(size_t)mv->face_vertexes[(size_t)er * 2]], vmap[(size_t)mv->face_vertexes[(size_t)er * 2 + 1]]);
        int friend_num = 10;                                                        // momAdvection.cpp:104-111
        if (mv->node_friends[(size_t)mv->face_nodes[(size_t)er * 2] * 6 + 5] < 0) friend_num--;
        if (mv->node_friends[(size_t)mv->face_nodes[(size_t)er * 2 + 1] * 6 + 5] < 0) friend_num--;
        const double dist_r = 1.0 / mv->face_node_dist[er];                         // mesh.cpp:624
        for (int j = 0; j < friend_num; j++) {
            const int f = mv->face_interp_friends[(size_t)er * 10 + j];
            nid[(size_t)j * Fp + e] = emap[(size_t)f];
            ncoef[(size_t)j * Fp + e] = mv->face_interp_weights[(size_t)er * 10 + j] * mv->face_len[f] * dist_r;   // momAdvection.cpp:129
        }
    }
    int* d_vnode = nullptr; int* d_nid = nullptr; int2* d_fvert = nullptr;
    double *d_vR = nullptr, *d_vsin = nullptr, *d_varea_r = nullptr, *d_ncoef = nullptr;
    if ((rc = upload(s, &d_vnode, vnode)) || (rc = upload(s, &d_vR, vR)) || (rc = upload(s, &d_vsin, vsin)) || (rc = upload(s, &d_varea_r, varea_r)) ||
        (rc = upload(s, &d_fvert, fvert)) || (rc = upload(s, &d_nid, nid)) || (rc = upload(s, &d_ncoef, ncoef)) ||
        (rc = dev_alloc(s, &s->d_nl_qv, (size_t)Vs)) || (rc = dev_alloc(s, &s->d_nl_fq, (size_t)Fp)) || (rc = dev_alloc(s, &s->d_nl_ekin, (size_t)Np)) ||
        (rc = dev_alloc(s, &s->d_nl_flux, (size_t)Fp)))
        return rc;
    for (void* p : {(void*)d_vnode, (void*)d_vR, (void*)d_vsin, (void*)d_varea_r, (void*)d_fvert, (void*)d_nid, (void*)d_ncoef, (void*)s->d_nl_qv,
                    (void*)s->d_nl_fq, (void*)s->d_nl_ekin, (void*)s->d_nl_flux})
        s->nl_owned.push_back(p);
    t.vnode = d_vnode; t.vR = d_vR; t.vsin = d_vsin; t.varea_r = d_varea_r; t.carea = s->d_area; t.fvert = d_fvert; t.nid = d_nid; t.ncoef = d_ncoef;
    t.cells = s->d_cells; t.grad = s->d_grad; t.dist = s->d_dist; t.eid = s->d_eid; t.area = s->d_area;
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    for (auto& g : s->graphs) cudaGraphExecDestroy(g.second);
    s->graphs.clear();
    s->nl_on = true;
    return ODIS_OK;
}

int odis_get_sh_coefficients(odis_solver* s, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->sh_on) return fail(ODIS_ERR_STATE, "self-gravity is not enabled");
    ODIS_CUDA(cudaSetDevice(s->device));
    const size_t R = (size_t)s->sh_rows;
    std::vector<double> b(R);
    ODIS_CUDA(cudaMemcpyAsync(b.data(), s->d_sh_b, R * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t j = 0; j < R; j++) {
        double acc = 0.0;
        for (size_t k = 0; k < R; k++) acc += s->sh_ginv_host[j * R + k] * b[k];
        out[j] = acc;
    }
    return ODIS_OK;
}

static int step_impl(odis_solver* s, int32_t nsteps, std::vector<cudaEvent_t>* marks);
static int check_halo_timeout(odis_solver* s);

int odis_step(odis_solver* s, int32_t nsteps) { return step_impl(s, nsteps, nullptr); }

static int step_profiled_impl(odis_solver* s, int32_t nsteps, float out[3]) {
    if (nsteps < 0 || nsteps > 100000) return fail(ODIS_ERR_ARG, "nsteps out of range");
    ODIS_CUDA(cudaSetDevice(s->device));
    std::vector<cudaEvent_t> marks((size_t)nsteps * 4);
    for (auto& e : marks) ODIS_CUDA(cudaEventCreate(&e));
    int rc = step_impl(s, nsteps, &marks);
    if (!rc && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = fail(ODIS_ERR_CUDA, "synchronize failed");
    double sum[3] = {0.0, 0.0, 0.0};
    if (!rc) {
        for (int k = 0; k < nsteps; k++)
            for (int q = 0; q < 3; q++) {
                float a = 0.f;
                cudaEventElapsedTime(&a, marks[(size_t)k * 4 + q], marks[(size_t)k * 4 + q + 1]);
                sum[q] += a;
            }
    }
    for (auto& e : marks) cudaEventDestroy(e);
    for (int q = 0; q < 3; q++) out[q] = (float)sum[q];
    return rc;
}

int odis_step_profiled(odis_solver* s, int32_t nsteps, float* edge_ms_out, float* cell_ms_out) {
    if (!s || !edge_ms_out || !cell_ms_out) return fail(ODIS_ERR_ARG, "NULL argument");
    float out[3];
    int rc = step_profiled_impl(s, nsteps, out);
    *edge_ms_out = out[0]; *cell_ms_out = out[1];
    return rc;
}

int odis_step_profiled_sh(odis_solver* s, int32_t nsteps, float* edge_ms_out, float* cell_ms_out, float* sh_ms_out) {
    if (!s || !edge_ms_out || !cell_ms_out || !sh_ms_out) return fail(ODIS_ERR_ARG, "NULL argument");
    float out[3];
    int rc = step_profiled_impl(s, nsteps, out);
    *edge_ms_out = out[0]; *cell_ms_out = out[1]; *sh_ms_out = out[2];
    return rc;
}

// the cell tendency history is three arrays whose roles rotate (temporalOperators.cpp:44-65 shifts in place)
static void rotate_cell_history(odis_solver* s, int mode) {
    const int l1 = s->he1, l2 = s->he2, fr = s->hefree;
    if (mode == odis::AB3_FIRST) { s->he2 = fr; s->hefree = l2; }                       // level 2 := f0
    else if (mode == odis::AB3_SECOND) { s->he1 = fr; s->hefree = l1; }                 // level 1 := f0
    else { s->he1 = fr; s->he2 = l1; s->hefree = l2; }                                  // shift
}

constexpr int kGraphSteps = 12;      // steps per captured graph: a multiple of the rotation period (2 x 2 x 3 -> 6)

// Launches of one time step on the solver's stream (also under stream capture) and the rotation of the buffers.
// One step of the nonlinear branch. Default (folded): the 4 launches of launch_step_nonlinear_folded — energy diagnostic of v^n inside the
// edge update, next potential inside the cell update. Otherwise: diagnostics of v^n (the linear edge kernel produces them on the fly), the launches of
// odis_kernels_nl.cu, the potential pass for the next step.
static int enqueue_step_nonlinear(odis_solver* s, int mode, std::vector<cudaEvent_t>* marks, int k) {
    const bool folded = s->nl_folded;
    // forcing for the next step (current_time = dt*(iter+1), evaluated at current_time + dt)
    const odis::StepScalars next = step_scalars(s, s->prm.dt * (double)(s->iter + 1) + s->prm.dt);
    if (!folded)
        odis::launch_edge_diagnostics(s->edge_tables(), s->phys, s->d_vl[s->cur], s->d_normal, nullptr, nullptr, s->d_block_partial, s->d_ticket,
                                      s->d_series + (s->iter - s->iter0), s->prm.block_threads, s->stream);
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4], s->stream);
    odis::NlState ns;
    ns.vl_in = s->d_vl[s->cur]; ns.vl_out = s->d_vl[1 - s->cur];
    ns.eu_in = s->d_eu[s->ecur]; ns.eu_out = s->d_eu[1 - s->ecur];
    ns.h1 = s->d_hv[s->hv1]; ns.h2 = s->d_hv[1 - s->hv1];
    ns.ch1 = s->d_he[s->he1]; ns.ch2 = s->d_he[s->he2]; ns.chw = s->d_he[s->hefree];
    ns.qv = s->d_nl_qv; ns.fq = s->d_nl_fq; ns.ekin = s->d_nl_ekin; ns.flux = s->d_nl_flux;
    ns.block_partial = s->d_block_partial; ns.ticket = s->d_ticket; ns.energy_out = s->d_series + (s->iter - s->iter0);
    if (folded) odis::launch_step_nonlinear_folded(s->nl, s->phys, ns, mode, s->cell_tables(s->N), next, s->stream);
    else odis::launch_step_nonlinear(s->nl, s->phys, ns, mode, s->stream);
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 1], s->stream);
    if (mode == odis::AB3_FULL) s->hv1 = 1 - s->hv1;
    rotate_cell_history(s, mode);
    s->ecur = 1 - s->ecur;
    if (!folded) {      // the potential in place on the new {eta,U}, in a pass of its own
        odis::CellState cs{s->d_vl[1 - s->cur], s->d_eu[s->ecur], s->d_eu[s->ecur], s->d_he[0], s->d_he[1], s->d_he[2], nullptr, 0, nullptr, nullptr};
        odis::launch_cell_step(s->cell_tables(s->N), s->phys, cs, odis::AB3_FULL, next, odis::CELL_UPDATE_U, s->prm.block_threads, nullptr, s->stream);
    }
    enqueue_planet(s, s->d_eu[s->ecur], s->N, next, nullptr);
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 2], s->stream);
    { int rc2 = enqueue_self_gravity(s, s->d_eu[s->ecur]); if (rc2) return rc2; }
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 3], s->stream);
    s->cur = 1 - s->cur;
    s->iter++;
    s->last_mode = mode;
    s->launches += (folded ? 0 : 2) + s->nl_launches() + planet_launches(s) + s->sh_launches();
    return ODIS_OK;
}

static int enqueue_step(odis_solver* s, int mode, bool dev_ctl, std::vector<cudaEvent_t>* marks, int k) {
    if (s->nl_on) return enqueue_step_nonlinear(s, mode, marks, k);
    const odis::EdgeTables et = s->edge_tables();
    const odis::CellTables ct = s->cell_tables(s->world > 1 ? s->N : s->No);     // partitioned: ghost cells are updated locally
    odis::EdgeState es;
    es.vl_in = s->d_vl[s->cur]; es.vl_out = s->d_vl[1 - s->cur]; es.eu = s->d_eu[s->ecur];
    es.h1 = s->d_hv[s->hv1]; es.h2 = s->d_hv[1 - s->hv1];
    es.block_partial = s->d_block_partial; es.ticket = s->d_ticket;
    es.energy_out = dev_ctl ? nullptr : s->d_series + (s->iter - s->iter0);
    es.ctl = dev_ctl ? s->d_ctl : nullptr; es.series = s->d_series; es.scal = s->d_scal;
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4], s->stream);
    // partitioned: ONE exchange per step, of v^{n+1} on the boundary edges. The staged edge kernel pushes it itself
    // (HaloInline) and the direct cell kernel waits for the neighbours' in its last CTAs; the other variants get a separate
    // exchange launch (which also waits). Ghost cells are updated locally, so {eta,U} is never exchanged.
    const bool part = s->world > 1, inline_e = part && s->pipe_edge;
    if (s->pipe_edge) {
        odis::HaloInline hv;
        if (inline_e) hv = halo_inline_edge(s, 1 - s->cur);
        if (s->edge_ids16) ODIS_CUDA(odis::launch_edge_step_pipe16(et, s->phys, es, mode, inline_e ? &hv : nullptr, s->d_sid16, s->d_tile_wide, s->stream));
        else ODIS_CUDA(odis::launch_edge_step_pipe(et, s->phys, es, mode, inline_e ? &hv : nullptr, s->stream));
    } else odis::launch_edge_step(et, s->phys, es, mode, s->prm.block_threads, s->stream);
    if (part && !inline_e) {
        int rc2 = halo_exchange(s, 0, s->d_vl[1 - s->cur], 1 - s->cur);
        if (rc2) return rc2;
    }
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 1], s->stream);
    if (mode == odis::AB3_FULL) s->hv1 = 1 - s->hv1;
    // the staged edge kernel finishes its own energy sum; the direct one leaves per-warp partials to the cell kernel
    odis::CellState cs{s->d_vl[1 - s->cur], s->d_eu[s->ecur], s->d_eu[1 - s->ecur], s->d_he[s->he1], s->d_he[s->he2], s->d_he[s->hefree],
                       s->d_block_partial, (s->Fo + 31) / 32, s->pipe_edge ? nullptr : es.energy_out, dev_ctl ? &s->d_ctl->cur : nullptr};
    // the next step's forcing time: current_time = dt*(iter+1), evaluated at current_time + dt
    const odis::StepScalars next = dev_ctl ? odis::StepScalars{} : step_scalars(s, s->prm.dt * (double)(s->iter + 1) + s->prm.dt);
    if (s->pipe_cell) {
        // staged cell update; with the self-gravity term to degree <= 4 it also leaves the harmonic sums of eta^{n+1} (this rank's cells)
        odis::HaloInline hc;
        if (inline_e) hc = halo_inline_cell(s);
        const odis::CellSgAccum sg = s->sg_accum();
        const odis::ShExchange x = s->sh_exchange();
        const bool with_sg = s->sh_on && s->sh_fused;
        ODIS_CUDA(odis::launch_cell_step_pipe(ct, s->phys, cs, mode, next, inline_e ? &hc : nullptr, with_sg ? &sg : nullptr,
                                              with_sg && part ? &x : nullptr, s->stream));
    } else {
        odis::HaloInline hc;
        if (inline_e) hc = halo_inline_cell(s);
        odis::launch_cell_step(ct, s->phys, cs, mode, next, odis::CELL_UPDATE_ETA | odis::CELL_UPDATE_U,
                               s->prm.block_threads, inline_e ? &hc : nullptr, s->stream);
    }
    rotate_cell_history(s, mode);
    s->ecur = 1 - s->ecur;
    enqueue_planet(s, s->d_eu[s->ecur], ct.n_active, next, dev_ctl ? &s->d_ctl->cur : nullptr);
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 2], s->stream);
    if (s->sh_on && s->sh_fused && s->sh_merged) {
        // solve + synthesis happened inside the cell update's launch
    } else if (s->sh_on && s->sh_fused) {
        const odis::ShExchange x = s->sh_exchange();
        odis::launch_sh_bsolve_synthesis(s->sh_tables(), s->sh_work(), part ? &x : nullptr, s->prm.g, s->d_eu[s->ecur], s->N, s->stream);
    } else { int rc2 = enqueue_self_gravity(s, s->d_eu[s->ecur]); if (rc2) return rc2; }
    if (marks) cudaEventRecord((*marks)[(size_t)k * 4 + 3], s->stream);
    s->cur = 1 - s->cur;
    s->iter++;
    s->last_mode = mode;
    s->launches += 2 + planet_launches(s) + s->sh_step_launches();
    return ODIS_OK;
}

static int step_impl(odis_solver* s, int32_t nsteps, std::vector<cudaEvent_t>* marks) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "odis_set_state has not been called");
    if (nsteps < 0) return fail(ODIS_ERR_ARG, "nsteps must be >= 0");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ensure_series(s, (size_t)(s->iter - s->iter0) + (size_t)nsteps + 1);
    if (rc) return rc;
    if (s->world > 1 && !s->connected) return fail(ODIS_ERR_STATE, "odis_halo_connect has not been called on this partitioned solver");
    // ---- two launches per step (+ one halo exchange launch after each when partitioned) ----
    const bool dev_ctl = s->pipe_edge && !s->nl_on;   // the staged edge kernel keeps the step counter / time factors on the device
    if (dev_ctl && nsteps > 0) {             // time factors of every step of this call, uploaded ahead (tidalPotentials.cpp:55-61)
        std::vector<odis::StepScalars> sc((size_t)nsteps);
        // forcing for the NEXT step: current_time = dt*(iter+1), evaluated at current_time + dt
        for (int k = 0; k < nsteps; k++) sc[(size_t)k] = step_scalars(s, s->prm.dt * (double)(s->iter + k + 1) + s->prm.dt);
        ODIS_CUDA(cudaMemcpyAsync(s->d_scal + (s->iter - s->iter0), sc.data(), sc.size() * sizeof(odis::StepScalars), cudaMemcpyHostToDevice, s->stream));
    }
    int done = 0;
    while (done < nsteps) {
        const int mode = ab3_mode(s, s->iter);
        // steady state: replay a captured graph of kGraphSteps steps (the buffer rotation has period 6)
        if (dev_ctl && s->use_graph && !marks && mode == odis::AB3_FULL && nsteps - done >= kGraphSteps) {
            const unsigned key = (unsigned)s->cur | (unsigned)s->ecur << 1 | (unsigned)s->hv1 << 2 | (unsigned)s->he1 << 3 | (unsigned)s->he2 << 5 |
                                 (unsigned)s->hefree << 7;
            auto it = s->graphs.find(key);
            if (it == s->graphs.end()) {
                const int64_t iter_save = s->iter, launches_save = s->launches;
                cudaGraph_t graph = nullptr;
                ODIS_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
                int rcc = ODIS_OK;
                for (int k = 0; k < kGraphSteps && !rcc; k++) rcc = enqueue_step(s, odis::AB3_FULL, true, nullptr, 0);
                cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
                s->iter = iter_save; s->launches = launches_save;      // nothing ran; the rotation is back where it started
                if (rcc) { if (graph) cudaGraphDestroy(graph); return rcc; }
                if (ce != cudaSuccess) return fail(ODIS_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
                cudaGraphExec_t exec = nullptr;
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) return fail(ODIS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
                it = s->graphs.emplace(key, exec).first;
            }
            const int reps = (nsteps - done) / kGraphSteps;
            for (int r = 0; r < reps; r++) ODIS_CUDA(cudaGraphLaunch(it->second, s->stream));
            const int adv = reps * kGraphSteps;
            s->graph_launches += reps;
            s->launches += (int64_t)adv * (2 + planet_launches(s) + s->sh_step_launches());
            s->iter += adv;
            s->last_mode = odis::AB3_FULL;
            done += adv;
            continue;
        }
        int rc2 = enqueue_step(s, mode, dev_ctl, marks, done);
        if (rc2) return rc2;
        done++;
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
    return check_halo_timeout(s);
}

int odis_get_field(odis_solver* s, int32_t field, double* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    ODIS_CUDA(cudaSetDevice(s->device));
    const int No = s->No, Fo = s->Fo;
    size_t count = 0;
    // a partitioned solver fills its own cells/edges of the global array and leaves zeros elsewhere (the caller sums the ranks'
    // arrays): the own entries come back compact, in device order (perm == nullptr below), and are spread over `out` on the host
    const bool part = s->world > 1;
    const int* eperm = part ? nullptr : s->d_edge_perm;
    const int* cperm = part ? nullptr : s->d_cell_perm;
    size_t own = 0, comps = 1;
    const int* hperm = nullptr;
    switch (field) {
        case ODIS_FIELD_VELOCITY:
            count = (size_t)s->Fg; own = (size_t)Fo; hperm = s->edge_perm.data();
            odis::launch_gather_component(Fo, eperm, s->d_vl[s->cur], 0, s->d_stage, s->stream);
            break;
        case ODIS_FIELD_ETA:
        case ODIS_FIELD_POTENTIAL:
            count = (size_t)s->Ng; own = (size_t)No; hperm = s->cell_perm.data();
            odis::launch_gather_component(No, cperm, s->d_eu[s->ecur], field == ODIS_FIELD_ETA ? 0 : 1, s->d_stage, s->stream);
            break;
        case ODIS_FIELD_DVDT:
        case ODIS_FIELD_DETADT: {
            // level 0 = newest tendency: as loaded before any step, else where the last step stored it
            // (temporalOperators.cpp:47,56,65)
            const int which0 = s->last_mode < 0 ? 0 : (s->last_mode == odis::AB3_FIRST ? 2 : 1);
            comps = 3;
            if (field == ODIS_FIELD_DVDT) {
                count = (size_t)s->Fg * 3; own = (size_t)Fo; hperm = s->edge_perm.data();
                odis::launch_gather_history(Fo, eperm, s->d_lvl0_v, s->d_hv[s->hv1], s->d_hv[1 - s->hv1], which0, s->d_stage, s->stream);
            } else {
                count = (size_t)s->Ng * 3; own = (size_t)No; hperm = s->cell_perm.data();
                odis::launch_gather_history(No, cperm, s->d_lvl0_e, s->d_he[s->he1], s->d_he[s->he2], which0, s->d_stage, s->stream);
            }
            break;
        }
        case ODIS_FIELD_VELOCITY_EN:
        case ODIS_FIELD_DISSIPATION: {
            int rc = run_diagnostics(s, true);
            if (rc) return rc;
            own = (size_t)Fo; hperm = s->edge_perm.data();
            if (field == ODIS_FIELD_VELOCITY_EN) {
                count = (size_t)s->Fg * 2; comps = 2;
                odis::launch_gather_pair(Fo, eperm, s->d_vavg, s->d_stage, s->stream);
            } else {
                count = (size_t)s->Fg;
                odis::launch_gather_scalar(Fo, eperm, s->d_ediss, s->d_stage, s->stream);
            }
            break;
        }
        default:
            return fail(ODIS_ERR_ARG, "unknown field id");
    }
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    if (!part) {
        ODIS_CUDA(cudaMemcpyAsync(out, s->d_stage, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        ODIS_CUDA(cudaStreamSynchronize(s->stream));
        return check_halo_timeout(s);
    }
    // partitioned: only the own entries cross PCIe (compact, device order), the host spreads them over the zeroed global array
    int rcp = ensure_pack(s, own * comps);
    if (rcp) return rcp;
    ODIS_CUDA(cudaMemcpyAsync(s->h_pack, s->d_stage, own * comps * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    std::memset(out, 0, count * sizeof(double));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    const double* h = s->h_pack;
    const int nt = pack_threads(s);
    if (comps == 1) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)own; i++) out[hperm[i]] = h[i];
    } else {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (long long i = 0; i < (long long)own; i++)
            for (size_t c = 0; c < comps; c++) out[(size_t)hperm[i] * comps + c] = h[(size_t)i * comps + c];
    }
    return check_halo_timeout(s);
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

// Forget the per-step dissipation series before the current step (entry 0 becomes the current state's): whole runs read only the
// newest entry, and the series / time-factor tables would otherwise grow by 56 B per step for as long as the run lasts.
int odis_trim_dissipation_series(odis_solver* s) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    if (s->iter == s->iter0) return ODIS_OK;
    ODIS_CUDA(cudaSetDevice(s->device));
    // the newest entry (if the diagnostics are current) moves to the front; the device-side step counter restarts with it
    const size_t last = (size_t)(s->iter - s->iter0);
    ODIS_CUDA(cudaMemcpyAsync(s->d_series, s->d_series + last, sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    ODIS_CUDA(cudaMemsetAsync(&s->d_ctl->count, 0, sizeof(unsigned long long), s->stream));
    s->iter0 = s->iter;
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

// ---- operator surface: the free functions the reference's loop calls (timeIntegrator.cpp:205-313), one call each, on host
// arrays in reference numbering. Built from the step kernels themselves (an AB3_FIRST launch leaves exactly the operator's
// result in the tendency slot), so each result is the same arithmetic as inside odis_step. The solver's device state is
// the scratch space: afterwards odis_step refuses to run until odis_set_state is called again. ----
static int op_begin(odis_solver* s) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (s->world > 1) return fail(ODIS_ERR_UNSUPPORTED, "operator calls need an unpartitioned solver");
    if (s->nl_on) return fail(ODIS_ERR_UNSUPPORTED, "operator calls cover the linear branch (advection; false) only");
    ODIS_CUDA(cudaSetDevice(s->device));
    int rc = ensure_series(s, 1);
    if (rc) return rc;
    s->have_state = false;          // the state arrays are about to be overwritten
    s->diag_current = false;
    return ODIS_OK;
}
static int op_load_velocity(odis_solver* s, const double* v) {
    ODIS_CUDA(cudaMemcpyAsync(s->d_stage, v, (size_t)s->Fg * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    odis::launch_scatter_x(s->F, s->d_edge_perm, s->d_stage, s->d_vl[s->cur], 0, s->stream);
    s->launches++;
    return ODIS_OK;
}
static int op_finish(odis_solver* s, double* out, size_t count) {
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaMemcpyAsync(out, s->d_stage, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return ODIS_OK;
}

int odis_op_update_momentum(odis_solver* s, const double* v, const double* eta, double* dvdt_out) {
    if (!v || !eta || !dvdt_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int rc = op_begin(s);
    if (rc || (rc = op_load_velocity(s, v))) return rc;
    ODIS_CUDA(cudaMemcpyAsync(s->d_stage, eta, (size_t)s->Ng * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    odis::launch_scatter_x(s->N, s->d_cell_perm, s->d_stage, s->d_eu[s->ecur], 1, s->stream);
    // one edge update in start-up mode: the tendency -g G eta + C v (updateMomentum.cpp:42) lands in history level 2
    odis::EdgeState es;
    es.vl_in = s->d_vl[s->cur]; es.vl_out = s->d_vl[1 - s->cur]; es.eu = s->d_eu[s->ecur];
    es.h1 = s->d_hv[s->hv1]; es.h2 = s->d_hv[1 - s->hv1];
    es.block_partial = s->d_block_partial; es.ticket = s->d_ticket;
    es.energy_out = s->d_series; es.ctl = nullptr; es.series = s->d_series; es.scal = s->d_scal;
    odis::launch_edge_step(s->edge_tables(), s->phys, es, odis::AB3_FIRST, s->prm.block_threads, s->stream);
    odis::launch_gather_scalar(s->Fo, s->d_edge_perm, s->d_hv[1 - s->hv1], s->d_stage, s->stream);
    s->launches += 3;
    return op_finish(s, dvdt_out, (size_t)s->Fg);
}

int odis_op_update_eta(odis_solver* s, const double* v, double* detadt_out) {
    if (!v || !detadt_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int rc = op_begin(s);
    if (rc || (rc = op_load_velocity(s, v))) return rc;
    // one cell update in start-up mode: h Div v (updateEta.cpp:39) lands in the free tendency slot
    odis::CellState cs{s->d_vl[s->cur], s->d_eu[s->ecur], s->d_eu[1 - s->ecur], s->d_he[s->he1], s->d_he[s->he2], s->d_he[s->hefree],
                       nullptr, 0, nullptr, nullptr};
    odis::launch_cell_step(s->cell_tables(s->No), s->phys, cs, odis::AB3_FIRST, odis::StepScalars{}, odis::CELL_UPDATE_ETA, s->prm.block_threads,
                           nullptr, s->stream);
    odis::launch_gather_scalar(s->No, s->d_cell_perm, s->d_he[s->hefree], s->d_stage, s->stream);
    s->launches += 2;
    return op_finish(s, detadt_out, (size_t)s->Ng);
}

int odis_op_forcing(odis_solver* s, double time, double* potential_out) {
    if (!potential_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int rc = op_begin(s);
    if (rc) return rc;
    double2* eu = s->d_eu[1 - s->ecur];
    ODIS_CUDA(cudaMemsetAsync(eu, 0, (size_t)s->Np * sizeof(double2), s->stream));     // potential NONE leaves the array as it was: zeros
    odis::CellState cs{s->d_vl[s->cur], eu, eu, s->d_he[0], s->d_he[1], s->d_he[2], nullptr, 0, nullptr, nullptr};
    odis::launch_cell_step(s->cell_tables(s->N), s->phys, cs, odis::AB3_FULL, step_scalars(s, time), odis::CELL_UPDATE_U,
                           s->prm.block_threads, nullptr, s->stream);
    enqueue_planet(s, eu, s->N, step_scalars(s, time), nullptr);
    odis::launch_gather_component(s->No, s->d_cell_perm, eu, 1, s->d_stage, s->stream);
    s->launches += 2 + planet_launches(s);
    return op_finish(s, potential_out, (size_t)s->Ng);
}

int odis_op_integrate_ab3_scalar(odis_solver* s, double* solution, double* dsolution_dt, int64_t iter, int32_t n) {
    if (!solution || !dsolution_dt) return fail(ODIS_ERR_ARG, "NULL argument");
    if (iter < 0) return fail(ODIS_ERR_ARG, "iter must be >= 0");
    int rc = op_begin(s);
    if (rc) return rc;
    if (n < 0 || n > s->Fg) return fail(ODIS_ERR_ARG, "n must be between 0 and the number of edges");
    if (n == 0) return ODIS_OK;
    double* d_sol = s->d_ediss;        // [F]
    double* d_hist = s->d_stage;       // [3F]
    ODIS_CUDA(cudaMemcpyAsync(d_sol, solution, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    ODIS_CUDA(cudaMemcpyAsync(d_hist, dsolution_dt, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    odis::launch_ab3_scalar(n, d_sol, d_hist, s->prm.dt, ab3_mode(s, iter), s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaMemcpyAsync(solution, d_sol, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    return op_finish(s, dsolution_dt, (size_t)n * 3);
}

int odis_op_interpolate_velocity(odis_solver* s, const double* v, double* v_avg_out) {
    if (!v || !v_avg_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int rc = op_begin(s);
    if (rc || (rc = op_load_velocity(s, v))) return rc;
    odis::launch_edge_diagnostics(s->edge_tables(), s->phys, s->d_vl[s->cur], s->d_normal, s->d_vavg, s->d_ediss, s->d_block_partial, s->d_ticket,
                                  s->d_series, s->prm.block_threads, s->stream);
    odis::launch_gather_pair(s->Fo, s->d_edge_perm, s->d_vavg, s->d_stage, s->stream);
    s->launches += 2;
    return op_finish(s, v_avg_out, (size_t)s->Fg * 2);
}

int odis_op_update_energy(odis_solver* s, const double* v_avg, const double* areas, double* e_flux_out, double* avg_flux_out) {
    if (!v_avg || !areas || !e_flux_out || !avg_flux_out) return fail(ODIS_ERR_ARG, "NULL argument");
    int rc = op_begin(s);
    if (rc) return rc;
    const size_t F = (size_t)s->Fg;
    ODIS_CUDA(cudaMemcpyAsync(s->d_stage, v_avg, 2 * F * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    ODIS_CUDA(cudaMemcpyAsync(s->d_stage + 2 * F, areas, F * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    odis::launch_energy_from_components(s->Fg, s->phys, s->d_stage, s->d_stage + 2 * F, s->d_ediss, s->d_block_partial, s->d_ticket, s->d_series,
                                        s->stream);
    s->launches++;
    ODIS_CUDA(cudaGetLastError());
    double sum = 0.0;
    ODIS_CUDA(cudaMemcpyAsync(e_flux_out, s->d_ediss, F * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaMemcpyAsync(&sum, s->d_series, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    *avg_flux_out = sum / sphere_area(s);        // energy.cpp:60
    return ODIS_OK;
}

// ---- output snapshots overlapped with stepping (the dump pipeline) ----
// begin: on the solver's stream, the diagnostics + renumbering of the requested fields into the slot's device buffer; on a
// second stream, behind an event, the device -> host copy into page-locked memory. The call returns at once and the caller
// can enqueue the next interval's steps: they run while the copy and the caller's file output proceed.
int odis_snapshot_begin(odis_solver* s, int32_t slot, uint32_t fields) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL solver");
    if (slot < 0 || slot > 1) return fail(ODIS_ERR_ARG, "snapshot slot must be 0 or 1");
    if (!s->have_state) return fail(ODIS_ERR_STATE, "no state");
    ODIS_CUDA(cudaSetDevice(s->device));
    // partitioned: the rank's OWN entries, compact, in device order (odis_get_partition_map gives their reference ids)
    const bool part = s->world > 1;
    const size_t N = part ? (size_t)s->No : (size_t)s->Ng, F = part ? (size_t)s->Fo : (size_t)s->Fg, total = N + 4 * F + 1;
    const int* cperm = part ? nullptr : s->d_cell_perm;
    const int* eperm = part ? nullptr : s->d_edge_perm;
    odis_solver::SnapshotSlot& sl = s->snap[slot];
    if (!s->copy_stream) ODIS_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    if (!sl.d_buf) {
        int rc = dev_alloc(s, &sl.d_buf, total);
        if (rc) return rc;
        ODIS_CUDA(cudaHostAlloc((void**)&sl.h_buf, total * sizeof(double), cudaHostAllocDefault));
        ODIS_CUDA(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
        ODIS_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    if (sl.pending) ODIS_CUDA(cudaStreamWaitEvent(s->stream, sl.done, 0));     // the slot's previous copy must have left the buffer
    int rc;
    const bool want_diag_fields = (fields & (ODIS_SNAP_VELOCITY_EN | ODIS_SNAP_DISSIPATION)) != 0;
    if ((rc = run_diagnostics(s, want_diag_fields))) return rc;
    double* d_eta = sl.d_buf; double* d_ven = d_eta + N; double* d_diss = d_ven + 2 * F; double* d_v = d_diss + F; double* d_sum = d_v + F;
    if (fields & ODIS_SNAP_ETA) { odis::launch_gather_component(s->No, cperm, s->d_eu[s->ecur], 0, d_eta, s->stream); s->launches++; }
    if (fields & ODIS_SNAP_VELOCITY_EN) { odis::launch_gather_pair(s->Fo, eperm, s->d_vavg, d_ven, s->stream); s->launches++; }
    if (fields & ODIS_SNAP_DISSIPATION) { odis::launch_gather_scalar(s->Fo, eperm, s->d_ediss, d_diss, s->stream); s->launches++; }
    if (fields & ODIS_SNAP_VELOCITY) { odis::launch_gather_component(s->Fo, eperm, s->d_vl[s->cur], 0, d_v, s->stream); s->launches++; }
    ODIS_CUDA(cudaGetLastError());
    ODIS_CUDA(cudaMemcpyAsync(d_sum, s->d_series + (s->iter - s->iter0), sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    ODIS_CUDA(cudaEventRecord(sl.ready, s->stream));
    ODIS_CUDA(cudaStreamWaitEvent(s->copy_stream, sl.ready, 0));
    auto d2h = [&](const double* d, size_t n) { return cudaMemcpyAsync(sl.h_buf + (d - sl.d_buf), d, n * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream); };
    if (fields & ODIS_SNAP_ETA) ODIS_CUDA(d2h(d_eta, N));
    if (fields & ODIS_SNAP_VELOCITY_EN) ODIS_CUDA(d2h(d_ven, 2 * F));
    if (fields & ODIS_SNAP_DISSIPATION) ODIS_CUDA(d2h(d_diss, F));
    if (fields & ODIS_SNAP_VELOCITY) ODIS_CUDA(d2h(d_v, F));
    ODIS_CUDA(d2h(d_sum, 1));
    ODIS_CUDA(cudaEventRecord(sl.done, s->copy_stream));
    sl.fields = fields;
    sl.pending = true;
    sl.iter = s->iter;
    return ODIS_OK;
}

int odis_snapshot_wait(odis_solver* s, int32_t slot, odis_snapshot_view* out) {
    if (!s || !out) return fail(ODIS_ERR_ARG, "NULL argument");
    if (slot < 0 || slot > 1) return fail(ODIS_ERR_ARG, "snapshot slot must be 0 or 1");
    odis_solver::SnapshotSlot& sl = s->snap[slot];
    if (!sl.pending) return fail(ODIS_ERR_STATE, "no snapshot was begun on this slot");
    ODIS_CUDA(cudaSetDevice(s->device));
    ODIS_CUDA(cudaEventSynchronize(sl.done));
    const bool part = s->world > 1;
    const size_t N = part ? (size_t)s->No : (size_t)s->Ng, F = part ? (size_t)s->Fo : (size_t)s->Fg;
    const double* h = sl.h_buf;
    out->eta = (sl.fields & ODIS_SNAP_ETA) ? h : nullptr;
    out->velocity_en = (sl.fields & ODIS_SNAP_VELOCITY_EN) ? h + N : nullptr;
    out->dissipation = (sl.fields & ODIS_SNAP_DISSIPATION) ? h + N + 2 * F : nullptr;
    out->velocity = (sl.fields & ODIS_SNAP_VELOCITY) ? h + N + 3 * F : nullptr;
    out->dissipation_avg = h[N + 4 * F] / sphere_area(s);
    out->iter = sl.iter;
    return ODIS_OK;
}

#ifdef ODIS_TRACE
// variant library only (scripts/halo_trace.py): device buffer of [2 kernels][slots][ctas][8 events][2] 64-bit time stamps on the current device
int odis_debug_trace_enable(void* device_buffer, int32_t slots, int32_t ctas) {
    ODIS_CUDA(odis::trace_enable(static_cast<unsigned long long*>(device_buffer), (unsigned int)slots, (unsigned int)ctas));
    return ODIS_OK;
}
#endif

int odis_get_iter(odis_solver* s, int64_t* iter_out) {
    if (!s || !iter_out) return fail(ODIS_ERR_ARG, "NULL argument");
    *iter_out = s->iter;
    return ODIS_OK;
}

int odis_get_footprint(odis_solver* s, int64_t* device_bytes_out, int64_t* alg_bytes_out) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    if (device_bytes_out) *device_bytes_out = (int64_t)s->device_bytes;
    if (alg_bytes_out) {
        *alg_bytes_out = 200LL * s->Fo + 128LL * s->No;    // SURVEY.md §8(d), this rank's share
        // self-gravity: Y streamed by the analysis (all rows) and the synthesis (degrees >= 2), {eta,U} read twice, U written
        if (s->sh_on && s->sh_stored) *alg_bytes_out += 8LL * s->sh_rows * s->No + 8LL * (s->sh_rows - odis::kShSkipRows) * s->N + 16LL * s->No + 24LL * s->N;
        // matrix-free: 4 trig values + {eta,U} per cell per pass, U written
        if (s->sh_on && !s->sh_stored) *alg_bytes_out += 48LL * s->No + 56LL * s->N;
    }
    return ODIS_OK;
}

int odis_get_launch_count(odis_solver* s, int64_t* launches_out) {
    if (!s || !launches_out) return fail(ODIS_ERR_ARG, "NULL argument");
    *launches_out = s->launches;
    return ODIS_OK;
}

// partitioned runs: did any in-kernel wait for a neighbour give up? (the stream must be idle)
static int check_halo_timeout(odis_solver* s) {
    if (s->world <= 1) return ODIS_OK;
    unsigned long long flag = 0;
    ODIS_CUDA(cudaMemcpy(&flag, &s->d_ctl->pad, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) return fail(ODIS_ERR_STATE, "halo exchange timed out: a neighbouring rank did not take the same steps (results are invalid)");
    if (s->sh_on) {
        ODIS_CUDA(cudaMemcpy(&flag, s->d_sh_xctl + 2, sizeof flag, cudaMemcpyDeviceToHost));
        if (flag) return fail(ODIS_ERR_STATE, "self-gravity all-reduce timed out: a rank did not take the same steps (results are invalid)");
    }
    return ODIS_OK;
}

int odis_synchronize(odis_solver* s) {
    if (!s) return fail(ODIS_ERR_ARG, "NULL argument");
    ODIS_CUDA(cudaSetDevice(s->device));
    ODIS_CUDA(cudaStreamSynchronize(s->stream));
    return check_halo_timeout(s);
}

void odis_destroy(odis_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream && s->have_state) halo_drain(s);        // neighbours may still be pushing into this rank's halo
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& g : s->graphs) cudaGraphExecDestroy(g.second);
    void* ptrs[] = {s->d_csr_e_first, s->d_csr_e_peer, s->d_csr_e_remote, s->d_halo_done,
                    s->d_ctl, s->d_scal, s->d_cells, s->d_grad, s->d_fcor, s->d_dist, s->d_sw, s->d_sid, s->d_sid16, s->d_tile_wide, s->d_normal, s->d_eid, s->d_area, s->d_trig,
                    s->d_trig_sq, s->d_vl[0], s->d_vl[1], s->d_eu[0], s->d_eu[1], s->d_hv[0], s->d_hv[1], s->d_he[0], s->d_he[1], s->d_he[2],
                    s->d_block_partial, s->d_ticket, s->d_series, s->d_vavg, s->d_ediss, s->d_edge_perm, s->d_cell_perm, s->d_stage,
                    s->d_lvl0_v, s->d_lvl0_e, s->d_send_e_local, s->d_send_e_remote, s->d_send_e_peer, s->d_send_c_local, s->d_send_c_remote,
                    s->d_send_c_peer, s->d_flags, s->d_halo_ticket, s->d_shY, s->d_shRec, s->d_shGinv, s->d_shFactor, s->d_sh_partial, s->d_sh_b, s->d_sh_s, s->d_sh_xblock, s->d_sh_xctl,
                    s->d_sg_cta, s->d_sg_ticket};
    for (int k = 0; k < kMaxPeers; k++)
        for (int j = 0; j < 5; j++)
            if (s->ipc_opened[k][j]) cudaIpcCloseMemHandle(s->ipc_opened[k][j]);
    for (int r = 0; r < odis::kShMaxWorld; r++)
        if (s->sh_ipc_opened[r]) cudaIpcCloseMemHandle(s->sh_ipc_opened[r]);
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (void* p : s->nl_owned)
        if (p) cudaFree(p);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    for (auto& sl : s->snap) {
        if (sl.d_buf) cudaFree(sl.d_buf);
        if (sl.h_buf) cudaFreeHost(sl.h_buf);
        if (sl.ready) cudaEventDestroy(sl.ready);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    if (s->h_pack) cudaFreeHost(s->h_pack);
    if (s->h_stage_pack) cudaFreeHost(s->h_stage_pack);
    if (s->pack_done) cudaEventDestroy(s->pack_done);
    if (s->d_stage_next) cudaFree(s->d_stage_next);
    if (s->staged) cudaEventDestroy(s->staged);
    if (s->consumed) cudaEventDestroy(s->consumed);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // extern "C"
