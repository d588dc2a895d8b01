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
    unsigned long long pad;         // set to 1 when a halo wait gave up
    StepScalars cur;
    long long spin_cycles;          // upper bound of an in-kernel wait for another rank / CTA in clock64 ticks (0: kHaloSpinCycles);
                                    // set at odis_create from ODIS_B200_WAIT_TIMEOUT_S
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

void launch_edge_step(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, int block_threads,
                      cudaStream_t stream);
// flags: CellFlags (update eta and/or the potential). block_threads: 128 (default), 256, 512
void launch_cell_step(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                      int flags, int block_threads, const HaloInline* halo, cudaStream_t stream);
// spins (bounded by kHaloSpinCycles) until every neighbour's flag has reached ctl->epoch[0]: all pushes of the
// exchanges this rank took part in have landed
void launch_halo_drain(const HaloWait& wait_v, StepCtl* ctl, cudaStream_t stream);
// Diagnostics / output fields of the current velocity (interpolation.cpp:31-59, energy.cpp:32-56):
// v_avg [F][2] and energy_diss [F]; either pointer may be null. Also leaves sum(eps_e*A_e) in energy_out.
void launch_edge_diagnostics(const EdgeTables& t, const Physics& p, const double2* vl, const double2* normal,
                             double2* v_avg, double* energy_diss, double* block_partial, unsigned int* ticket,
                             double* energy_out, int block_threads, cudaStream_t stream);
int edge_grid_blocks(int n_edges, int block_threads);

#ifdef ODIS_TRACE
cudaError_t trace_enable(unsigned long long* buf, unsigned int slots, unsigned int ctas);   // tuning aid, odis_kernels_pipe.cu
#endif
// Pipelined variants (odis_kernels_pipe.cu): persistent CTAs, tables staged through shared memory by
// cp.async.bulk + mbarrier, 128-entity tiles. Same results bit for bit. All arrays a tile touches must
// be allocated up to the next multiple of pipe_tile().
int pipe_tile();
// opts the staged kernels into their dynamic shared-memory size on the current device (before any stream capture)
cudaError_t pipe_configure();
cudaError_t launch_edge_step_pipe(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                  cudaStream_t stream);
// Staged cell update (the default with the staged edge kernel): persistent CTAs, bit-identical to launch_cell_step with
// CELL_UPDATE_ETA | CELL_UPDATE_U. halo != nullptr: partitioned solver (the last tiles wait for the neighbours' push, ghost values are
// read past L1). sg != nullptr: the harmonic analysis b = Y eta^{n+1} of the cells [0, sg->n_fit) is accumulated on the way (degrees
// 2..4, matrix-free basis) and left in sg->b_out — or, partitioned (x != nullptr), published for the all-reduce through peer memory
// that launch_sh_bsolve_synthesis (odis_sh.cuh) completes.
constexpr int kCellMaxRows = 10;
struct CellRows {             // which trig rows a launch stages, and where the potential / the harmonic basis find their inputs (bit fields:
    int n;                    // the kernel indexes them with compile-time shifts)        rows staged
    int n_pot;                // inputs of the potential
    unsigned long long src;   // 4 bits per stage row k: its table row, 0..7 CellTables.trig, 8..9 CellTables.trig_sq
    unsigned long long pot;   // 5 bits per input k of the potential: stage row; bit 4 set: the square of that row's value
    unsigned basis;           // 4 bits each: stage rows of cos lat, sin lat, cos lon, sin lon
};
CellRows cell_rows_for(int potential, bool with_basis);
struct CellSgAccum {
    int l_max;                // 2..4 (0: off)
    int n_fit;                // cells [0, n_fit) enter the least-squares fit
    double* cta_partial;      // [rows][cta_stride] per-CTA sums
    int cta_stride;           // >= cell_pipe_grid(n_active)
    double* b_out;            // [rows] this rank's sums (unpartitioned: the final b; merged + partitioned: the all-reduced b, for read-back)
    unsigned int* ticket;     // zero before the first launch (every launch leaves it zero again)
    // merged kernel (cell_pipe_merged()): behind a grid-wide barrier every CTA solves s = g factor (Ginv b) and adds the term to U of
    // the tiles it has just updated, so that the separate solve + synthesis launch is not needed (2 launches per step)
    int merged;
    const double* ginv;       // [rows][rows]
    const double* factor;     // [rows]
    double g;
    double* s_out;            // [rows] for read-back
    unsigned int* bar;        // [2] arrival count, generation (zero before the first launch)
};
struct ShExchange;
bool cell_pipe_supports_sg(int l_max);
bool cell_pipe_merged();              // the merged cell + solve + synthesis kernel is in use (cooperative launch; see odis_kernels_pipe.cu)
int cell_pipe_grid(int n_cells);      // upper bound of the CTAs a launch over n_cells cells uses
cudaError_t launch_cell_step_pipe(const CellTables& t, const Physics& p, const CellState& s, int mode, const StepScalars& next,
                                  const HaloInline* halo, const CellSgAccum* sg, const ShExchange* x, cudaStream_t stream);
// Opt-in: the staged edge kernel with narrow stencil ids. sid16 = [tiles][10][128] 16-bit offsets from the edge's own id (0 for an
// empty slot), tile_wide[tile] != 0 where an offset does not fit (that tile is read from EdgeTables.sid as usual). Bit-identical results.
// edge_ids16_fits: false when a CTA would hold more tiles than its flag buffer (the launcher then returns cudaErrorInvalidValue).
bool edge_ids16_fits(int n_edges);
cudaError_t launch_edge_step_pipe16(const EdgeTables& t, const Physics& p, const EdgeState& s, int mode, const HaloInline* halo,
                                    const short* sid16, const unsigned char* tile_wide, cudaStream_t stream);

// ---- halo exchange between ranks (one GPU each): peers' arrays are mapped into this process ----
constexpr int kHaloMaxPeers = 8;
constexpr long long kHaloSpinCycles = 20000000000ll;   // ~10 s: default upper bound of any in-kernel wait for a neighbour
// (configurable per solver: StepCtl::spin_cycles, environment variable ODIS_B200_WAIT_TIMEOUT_S read by odis_create)
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
