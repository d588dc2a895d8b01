// Device kernels of the LTE time step (sm_100a, FP64, built with -fmad=false).
//
// One reference time step (/root/reference/src/timeIntegrator.cpp:205-313) is two launches:
//
//   edge_step   for every edge e: tangential-velocity reconstruction of v^n over the 10-point
//               TRiSK stencil — used both for the Coriolis term of dv/dt and for the dissipated
//               energy of v^n —, pressure gradient from eta^n, linear drag, gradient of the tidal
//               potential U(t_n + dt), Adams-Bashforth update, forward-Euler drag/forcing add.
//               Replaces updateMomentum (updateMomentum.cpp:42), the drag/forcing SpMV
//               (timeIntegrator.cpp:219), integrateAB3scalar (temporalOperators.cpp:36-65),
//               timeIntegrator.cpp:242 and, for the previous step, interpolateVelocity
//               (interpolation.cpp:31-59) + updateEnergy (energy.cpp:32-60).
//   cell_step   for every cell i: divergence of v^{n+1} over its 5/6 edges, Adams-Bashforth update
//               of eta, and the tidal potential for the next step. Replaces updateEta
//               (updateEta.cpp:39), integrateAB3scalar and forcing (tidalPotentials.cpp:80-172).
//
// Floating point follows the reference's operation order so that v and eta can agree with the
// CPU solver to the last bit: sums run in the reference's CSR column order (ascending reference
// id — baked into the slot order of the stencil tables on the host), scalar*coefficient products
// are formed before the multiply with the field, and FMA contraction is off.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace odis {

constexpr int kStencil = 10;   // TRiSK tangential stencil width (9 beside pentagons)
constexpr int kCellEdges = 6;  // edges per cell (5 for the 12 pentagons)

// potential types handled on device; values follow enum Potential (include/globals.h:60-76)
enum DevPotential : int { P_OBLIQ = 0, P_OBLIQ_WEST = 1, P_ECC = 5, P_FULL = 8, P_FULL2 = 9, P_PLANET = 13, P_NONE = 16 };

// AB3 start-up modes (temporalOperators.cpp:36-65)
enum Ab3Mode : int { AB3_FIRST = 0, AB3_SECOND = 1, AB3_FULL = 2 };

struct StepScalars {          // per-step host-evaluated trigonometry (tidalPotentials.cpp:55-61)
    double cosM, sinM, cos2M, sin2M, cos3M, cos4M;
};

// Device-resident step bookkeeping, so that a captured CUDA graph can be replayed without per-step host arguments:
// the edge kernel's last CTA files the energy sum under series[count], copies the host-evaluated time factors of
// this step (scal[count], uploaded ahead for the whole odis_step call) to `cur` for the cell kernel that follows,
// and advances `count`; the halo exchange kernels keep their epochs here.
struct StepCtl {
    unsigned long long count;       // steps taken since odis_set_state
    unsigned long long epoch[2];    // [0] halo exchanges ({v,l} of the boundary edges) published so far; [1] unused
    unsigned long long pad;         // set to 1 when a halo wait gave up (kHaloSpinCycles)
    StepScalars cur;
};

struct HaloInline;   // halo exchange fused into the step kernels, defined below
struct HaloWait;

struct EdgeTables {
    int n_edges;              // edges this rank updates
    int stride;               // SoA row stride of sid / sw (n_edges rounded up to the 128-edge tile)
    const int2* cells;        // [F] inner, outer cell (device numbering)
    const double2* grad;      // [F] G_e,inner = -m_e0/d_e ; G_e,outer = +m_e1/d_e   (mesh.cpp:3076-3080)
    const double* fcor;       // [F] (-2.0*Omega)*sin(lat_e)                          (mesh.cpp:2881)
    const double* dist;       // [F] d_e = face_node_dist
    const int* sid;           // [10][stride] stencil edge ids, slot order = ascending reference id, -1 pad
    const double* sw;         // [10][stride] TRiSK weights w_ee' in the same slot order
};

struct CellTables {
    int n_cells;              // SoA stride = cells held (own + halo)
    int n_active;             // cells this launch updates (own cells; all held cells for the potential-only pass)
    const int* eid;           // [6][N] edge ids, bit 31 set when the cell is the edge's outer cell; -1 pad
    const double* area;       // [N] control_volume_surf_area_map
    const double* trig;       // [8][N] cosLat sinLat cosLon sinLon cos2Lat sin2Lat cos2Lon sin2Lon   (mesh.cpp:2132-2142)
    const double* trig_sq;    // [2][N] cos^2 lat, sin^2 lat                                           (mesh.cpp:2144-2145)
};

struct Physics {
    double g, h, alpha, dt;
    double area_sphere_inv;   // unused by kernels; 1/(4 pi r^2) applied on the host like energy.cpp:60
    double factor, factor2;   // potential prefactors (tidalPotentials.cpp:84,106,120,135,160-162)
    double ecc, obl;
    int potential;
    int friction;             // 0 linear, 1 quadratic (energy.cpp:30-56)
};

struct EdgeState {
    const double2* vl_in;     // [F] {v^n, l_e}
    double2* vl_out;          // [F] {v^{n+1}, l_e}
    const double2* eu;        // [N] {eta^n, U(t_n+dt)}
    double* h1;               // [F] dv/dt of step n-1
    double* h2;               // [F] dv/dt of step n-2
    double* block_partial;    // direct kernel: [ceil(F/32)] per-warp sums of eps_e * A_e for v^n, finished by the next cell_step;
                              // staged kernel: [grid] per-CTA sums, finished by its own last CTA
    unsigned int* ticket;     // last-CTA-done counter (staged kernel, edge_diagnostics)
    double* energy_out;       // where the finished sum goes (series[iter]) when ctl == nullptr
    StepCtl* ctl;             // staged kernel: device-side step counter (see StepCtl); series / scal are indexed by it
    double* series;
    const StepScalars* scal;
};

struct CellState {
    const double2* vl;        // [F] {v^{n+1}, l_e}
    const double2* eu_in;     // [N] {eta^n, U}
    double2* eu_out;          // [N] {eta^{n+1}, U(t_{n+1}+dt)}; may alias eu_in (each cell touches only its own entry)
    const double* h1;         // [N] d eta/dt of step n-1
    const double* h2;         // [N] d eta/dt of step n-2
    double* hw;               // [N] where this step's tendency goes (the host rotates three arrays)
    const double* energy_partial;   // per-warp partials left by edge_step (block 0 sums them), or unused
    int n_energy_partials;
    double* energy_out;             // nullptr: no energy sum to finish
    const StepScalars* next_dev;    // non-null: the time factors are read from device memory (StepCtl::cur) instead of the argument
};

// cell update flags
enum CellFlags : int { CELL_UPDATE_ETA = 1, CELL_UPDATE_U = 2 };

// ---- fused step (odis_kernels_fused.cu): cell update of the previous step + edge update in one kernel ----
struct FusedTables {
    EdgeTables e;
    const unsigned long long* cmap;   // [stride] per edge: for its two cells, which of {own, stencil slot 1..10} is the cell's
                                      // m-th edge in ascending reference id (4 bits, 15 = none) + outer-cell sign bit,
                                      // 6 entries x 5 bits per cell; bits 60/61: this edge stores cell 0 / cell 1
    const double* area;               // [cell_stride]
    const double* trig;               // [8][cell_stride]
    const double* trig_sq;            // [2][cell_stride]
    int cell_stride;
};
struct FusedState {
    const double2* vl_in;
    double2* vl_out;
    const double2* eu_in;             // {eta^{n-1}, U(t_n+dt)} (eta^n already when update_eta == 0)
    double2* eu_out;                  // {eta^n, U(t_{n+1}+dt)}
    double* h1;                       // edge tendency history, updated in place as in edge_step
    double* h2;
    const double* ch1;                // cell tendency history levels 1, 2
    const double* ch2;
    double* chw;                      // new cell tendency
    double* block_partial;            // [grid] one energy partial per CTA (summed in index order by the last CTA to finish)
    unsigned int* ticket;
    double* energy_out;
};
cudaError_t launch_step_fused(const FusedTables& t, const Physics& p, const FusedState& s, int mode_edge, int mode_cell, int update_eta,
                              const StepScalars& next, cudaStream_t stream);

void launch_edge_step(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, int block_threads,
                      cudaStream_t stream);
// flags: CellFlags (update eta and/or the potential). block_threads: 128 (default), 256, 512, or kCellOccupancyVariant (128 threads with
// the register count capped at 64 for 50 % occupancy; opt-in, odis_params.reserved[0] bit 6)
constexpr int kCellOccupancyVariant = -128;
// 128 threads + the streamed rows of the tile one GPU-full of CTAs ahead prefetched into L2 (default registers / capped at 64)
constexpr int kCellPrefetchVariant = -129;
constexpr int kCellPrefetchOccupancyVariant = -130;
void launch_cell_step(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                      int flags, int block_threads, const HaloInline* halo, cudaStream_t stream);
// Opt-in variant for runs with the self-gravity term (odis_params.reserved[0] bit 4): the cell update also accumulates the
// harmonic analysis b = Y eta^{n+1} of the cells [0, n_fit) (matrix-free basis, degrees 2..kCellSgMaxDegree), leaving sums over
// groups of kCellSgGroup consecutive CTAs; launch_sh_solve_synthesis (odis_sh.cuh) finishes the sum, solves and adds the term
// to the potential: 3 launches per step instead of 5. Unpartitioned solvers.
constexpr int kCellSgThreads = 128;
constexpr int kCellSgGroup = 32;
constexpr int kCellSgMaxDegree = 4;
struct CellSgWork {
    int l_max;
    int n_fit;                      // cells [0, n_fit) enter the least-squares fit
    double* cta_partial;            // [rows][cta_stride]
    int cta_stride;                 // >= cell_sg_ctas(n_active)
    double* group_partial;          // [rows][group_stride]
    int group_stride;               // >= ceil(cell_sg_ctas / kCellSgGroup)
    unsigned int* group_ticket;     // [group_stride], zero before the first launch (each launch leaves it zero again)
    int prefetch_ahead;             // > 0: prefetch the rows of the tile this many CTAs ahead into L2 (CTAs resident on the GPU); 0: off
};
int resident_cell_ctas(bool capped);   // 128-thread cell-update CTAs resident on the current GPU (SMs x 6, or x 8 with the register cap)
cudaError_t cell_sg_configure();    // recurrence coefficients -> constant memory (once per device, before any capture)
int cell_sg_ctas(int n_active);
bool cell_sg_supports(int l_max);
void launch_cell_step_sg(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next, const CellSgWork& sg,
                         bool cap_registers, cudaStream_t stream);
// The same on a partitioned solver (t.n_active = own + ghost cells, sg.n_fit = own cells): boundary CTAs wait for the neighbours' halo push,
// and the last group of CTAs to finish publishes this rank's harmonic sums for the in-kernel all-reduce that
// launch_sh_allsolve_synthesis (odis_sh.cuh) completes. sg.group_ticket needs one more counter at [group_stride].
struct HaloInline;
struct ShExchange;
void launch_cell_step_sgx(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next, const CellSgWork& sg,
                          const HaloInline& halo, const ShExchange& x, cudaStream_t stream);
// spins (bounded by kHaloSpinCycles) until every neighbour's flag has reached ctl->epoch[0]: all pushes of the
// exchanges this rank took part in have landed
void launch_halo_drain(const HaloWait& wait_v, StepCtl* ctl, cudaStream_t stream);
// Diagnostics / output fields of the current velocity (interpolation.cpp:31-59, energy.cpp:32-56):
// v_avg [F][2] and energy_diss [F]; either pointer may be null. Also leaves sum(eps_e*A_e) in energy_out.
void launch_edge_diagnostics(const EdgeTables& t, const Physics& p, const double2* vl, const double2* normal,
                             double2* v_avg, double* energy_diss, double* block_partial, unsigned int* ticket,
                             double* energy_out, int block_threads, cudaStream_t stream);
int edge_grid_blocks(int n_edges, int block_threads);

// Pipelined variants (odis_kernels_pipe.cu): persistent CTAs, tables staged through shared memory by
// cp.async.bulk + mbarrier, 128-entity tiles. Same results bit for bit. All arrays a tile touches must
// be allocated up to the next multiple of pipe_tile().
int pipe_tile();
// opts the staged kernels into their dynamic shared-memory size on the current device (before any stream capture)
cudaError_t pipe_configure();
cudaError_t launch_edge_step_pipe(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                  cudaStream_t stream);
cudaError_t launch_cell_step_pipe(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                                  cudaStream_t stream);
// Opt-in: the staged edge kernel with narrow stencil ids. sid16 = [tiles][10][128] 16-bit offsets from the edge's own id (0 for an
// empty slot), tile_wide[tile] != 0 where an offset does not fit (that tile is read from EdgeTables.sid as usual). Bit-identical results.
// edge_ids16_fits: false when a CTA would hold more tiles than its flag buffer (the launcher then returns cudaErrorInvalidValue).
bool edge_ids16_fits(int n_edges);
cudaError_t launch_edge_step_pipe16(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                    const short* sid16, const unsigned char* tile_wide, cudaStream_t stream);

// ---- halo exchange between ranks (one GPU each): peers' arrays are mapped into this process ----
constexpr int kHaloMaxPeers = 8;
constexpr long long kHaloSpinCycles = 20000000000ll;   // ~10 s: upper bound of any in-kernel wait for a neighbour
struct HaloRemote {
    double2* data[kHaloMaxPeers];                 // the peer's {v,l} (or {eta,U}) array, its local numbering
    unsigned long long* flags[kHaloMaxPeers];     // the peer's epoch flags [2][world]
};
struct HaloWait {
    int n_peers;
    const unsigned long long* flag[kHaloMaxPeers];   // my flags that the peers raise
};
// Halo exchange fused into the step kernels (partitioned runs): one exchange per step, of the new edge velocities.
//   edge kernel: the own edges [0, n_bnd) are the boundary and are updated first; each stores its new {v,l} straight
//                into the ghost slots of the neighbours that hold it (NVLink peer stores) and, when the last boundary
//                tile is done, the epoch ctl->epoch[0]+1 is published to every neighbour (system-scope release). The
//                interior is updated while stores and flag are in flight. It never waits: the ghost values it reads
//                were awaited by the previous cell kernel.
//   cell kernel: the cells [wait_from, n) - own cells with a non-own edge, then the ghost cells, whose eta every rank
//                updates itself instead of receiving it - come last; their CTAs wait until every neighbour's flag has
//                reached ctl->epoch[0] and gather with ld.global.cg (ghost slots are written by other GPUs while the
//                kernel runs). Nothing is sent.
struct HaloInline {
    int n_bnd;                        // edge kernel: boundary edges; 0: not partitioned (nothing below is read)
    int wait_from;                    // cell kernel: first cell that reads ghost edges; INT_MAX: not partitioned
    int n_peers;
    int flag_slot;                    // my slot in the neighbours' flag arrays
    const int* send_first;            // [n_bnd + 1] CSR over the boundary edges
    const int* send_peer;             // [n_send] neighbour index
    const int* send_remote;           // [n_send] ghost slot in that neighbour's numbering
    HaloRemote remote;                // the neighbours' {v,l} arrays (this step's output buffer) and flag arrays
    HaloWait wait_v;                  // my flags the neighbours raise
    unsigned int* done;               // boundary tiles finished (reset by the publisher)
    StepCtl* ctl;
};

// One launch per exchange: remote.data[peer[k]][remote_idx[k]] = src[local_idx[k]] for every k (direct stores into
// the neighbours' memory over NVLink); the last block to finish then publishes epoch E = ctl->epoch[kind] + 1 to
// every peer (system-scope release store to remote.flags[p][flag_slot]), spins (one thread per peer, system-scope
// acquire loads, bounded) until all of its own flags in `w` reach E, and records ctl->epoch[kind] = E. Used by the
// kernel variants that do not carry the exchange themselves (direct-load edge kernel, fused one-launch step). The epoch lives on
// the device so that the launch has no per-step argument (CUDA graph replay).
void launch_halo_exchange(int n, const int* local_idx, const int* remote_idx, const int* peer, const double2* src, const HaloRemote& remote,
                          const HaloWait& w, int flag_slot, int kind, StepCtl* ctl, unsigned int* ticket, cudaStream_t stream);

// ---- renumbering on the device: fields cross the C ABI in reference numbering, perm[new] = old ----
// x component of a double2 array from a reference-ordered source (src == nullptr: zeros); y untouched/kept.
void launch_scatter_x(int n, const int* perm, const double* src_ref, double2* dst_new, int zero_y, cudaStream_t stream);
// AB3 history [n][3] in reference order -> level-0 copy, level 1, level 2 in device order (src == nullptr: zeros)
void launch_scatter_history(int n, const int* perm, const double* src_ref3, double* lvl0_new, double* h1_new, double* h2_new,
                            cudaStream_t stream);
// component (0 = x, 1 = y) of a device-ordered double2 array -> reference order
void launch_gather_component(int n, const int* perm, const double2* src_new, int component, double* dst_ref, cudaStream_t stream);
// device-ordered double2 array -> reference-ordered [n][2]
void launch_gather_pair(int n, const int* perm, const double2* src_new, double* dst_ref2, cudaStream_t stream);
// device-ordered scalar array -> reference order
void launch_gather_scalar(int n, const int* perm, const double* src_new, double* dst_ref, cudaStream_t stream);
// history levels in device order -> reference-ordered [n][3]; level 0 is lvl0_new, h1_new or h2_new depending on `which0` (0,1,2)
void launch_gather_history(int n, const int* perm, const double* lvl0_new, const double* h1_new, const double* h2_new, int which0,
                           double* dst_ref3, cudaStream_t stream);

// PLANET forcing (tidalPotentials.cpp:176-225) in a pass of its own over the cells [0, n) after the cell update (the step kernels
// leave U = 0 for this type): U_i = factor (3 (cos lat_i (cos lon_i cosphi + sin lon_i sinphi))^2 - p), with the four time factors
// evaluated on the host and handed over in StepScalars {cosM: cosphi, sinM: sinphi, cos2M: factor, sin2M: p} — by value, or read
// from `dev` (StepCtl::cur) under graph replay.
void launch_planet_potential(const CellTables& t, const StepScalars& host, const StepScalars* dev, double2* eu, int n, cudaStream_t stream);

// ---- operator surface (odis_op_*): reference-ordered arrays staged on the device ----
// integrateAB3scalar (temporalOperators.cpp:17-68) in place on sol[n] and hist3[n][3]
void launch_ab3_scalar(int n, double* sol, double* hist3, double dt, int mode, cudaStream_t stream);
// updateEnergy (energy.cpp:13-62): e_flux[n] from vel2[n][2]; energy_out = sum e_flux*areas (block tree; block_partial needs
// ceil(n/128) entries)
void launch_energy_from_components(int n, const Physics& p, const double* vel2, const double* areas, double* e_flux, double* block_partial,
                                   unsigned int* ticket, double* energy_out, cudaStream_t stream);

}  // namespace odis
