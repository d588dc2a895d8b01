/* odis_b200.h — C ABI of the B200-native LTE time-step engine (drop-in for the GeodesicODIS hot path).
 *
 * The reference (hamishHay/GeodesicODIS) has no FFI or plugin interface; its boundary is the
 * C++ call surface of one process: main() -> Globals -> Mesh -> solveODIS -> ab3Explicit and the
 * free functions that loop calls. Every entry point below names the reference interface it
 * replaces (paths are relative to the reference tree). All arguments are plain pointers, sizes and
 * scalars; arrays are in the reference's own row-major layouts and numbering. The library owns
 * all device memory and any internal renumbering. Functions return ODIS_OK (0) or a negative
 * odis_status; odis_last_error() gives the message for the calling thread.
 *
 * There is no CPU fallback: entry points in the "solver" group fail with ODIS_ERR_CUDA when no
 * CUDA device is usable. The "config", "grid" and "mesh" groups are host-only.
 */
#ifndef ODIS_B200_H
#define ODIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum odis_status {
    ODIS_OK = 0,
    ODIS_ERR_ARG = -1,        /* bad argument / unknown key / size mismatch */
    ODIS_ERR_IO = -2,         /* file missing or malformed */
    ODIS_ERR_GRID = -3,       /* grid is not a usable closed icosahedral Voronoi grid */
    ODIS_ERR_CONFIG = -4,     /* input.in value rejected (the reference would TerminateODIS) */
    ODIS_ERR_CUDA = -5,       /* no device, launch or allocation failure */
    ODIS_ERR_UNSUPPORTED = -6,/* configuration outside the LTE hot path */
    ODIS_ERR_STATE = -7       /* call order violated (e.g. step before set_state) */
} odis_status;

const char* odis_last_error(void);
/* library / build identification, e.g. "odis_b200 0.1 sm_100a" */
const char* odis_version(void);

/* ------------------------------------------------------------------------------------------
 * config — replaces `new Globals(0)`: src/globals.cpp:35-323 (ctor), :325-477 (ReadGlobals),
 * :479-571 (SetDefault) and applySurfaceBCs, src/boundaryConditions.cpp:7-399.
 * Keys are the input.in key strings (src/globals.cpp:58-203), lower case.
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_config odis_config;

/* Titan defaults only (Globals(1)). */
int odis_config_create(odis_config** out);
/* Defaults, then <run_dir>/input.in, then the derived values (period rounding, enums, surface BCs). */
int odis_config_load(const char* run_dir, odis_config** out);
/* Set one key from its input.in text form; call odis_config_finalize afterwards. */
int odis_config_set(odis_config* cfg, const char* key, const char* value_text);
int odis_config_finalize(odis_config* cfg);
int odis_config_get_double(const odis_config* cfg, const char* key, double* out);
int odis_config_get_int(const odis_config* cfg, const char* key, int32_t* out);
int odis_config_get_bool(const odis_config* cfg, const char* key, int32_t* out);
/* Copies at most buflen-1 bytes + NUL. */
int odis_config_get_string(const odis_config* cfg, const char* key, char* buf, int32_t buflen);
/* Enumerations after finalize, numbered as include/globals.h:45-80:
 * which = 0 friction, 1 surface, 2 solver, 3 potential, 4 initial condition. */
int odis_config_get_enum(const odis_config* cfg, int32_t which, int32_t* out);
void odis_config_free(odis_config* cfg);

/* Time-step quantisation of Mesh::CalcMaxTimeStep, src/mesh.cpp:1601-1618. */
int odis_quantise_time_step(double period, double target_dt, double* dt_out, int32_t* steps_per_period_out);

/* ------------------------------------------------------------------------------------------
 * grid + mesh — replaces `new Mesh(...)`: src/mesh.cpp:32-197 (ctor), :4016-4101 (ReadMeshFile)
 * and the table builders it calls (:436-506, :508-1154, :1384-1484, :1828-1867, :2122-2152).
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_mesh odis_mesh;

/* Borrowed views of the mesh tables, reference names and layouts (include/mesh.h:77-187).
 * Valid until odis_mesh_free. Integer tables are int32; pads are -1 (0 in face_interp_friends). */
typedef struct odis_mesh_view {
    int32_t n_cells, n_edges, n_vertices;
    double radius;
    const double* node_pos_sph;                 /* [N][2] lat, lon (rad) */
    const int32_t* node_friends;                /* [N][6] */
    const double* centroid_pos_sph;             /* [N][6][2] */
    const double* control_volume_surf_area_map; /* [N] */
    const int32_t* faces;                       /* [N][6] */
    const int32_t* node_face_dir;               /* [N][6] */
    const int32_t* vertexes;                    /* [N][6] */
    const int32_t* face_nodes;                  /* [F][2] */
    const int32_t* face_vertexes;               /* [F][2] */
    const int32_t* face_interp_friends;         /* [F][10] */
    const double* face_interp_weights;          /* [F][10] */
    const double* face_len;                     /* [F] */
    const double* face_node_dist;               /* [F] */
    const double* face_centre_m;                /* [F][2] */
    const double* face_centre_pos_sph;          /* [F][2] */
    const double* face_intercept_pos_sph;       /* [F][2] */
    const double* face_area;                    /* [F] */
    const double* face_normal_vec_map;          /* [F][2] */
    const double* vertex_pos_sph;               /* [V][2] */
    const int32_t* vertex_nodes;                /* [V][3] */
    const double* vertex_R;                     /* [V][3] */
} odis_mesh_view;

/* Read input_files/grid_l<L>.txt style file and build all tables for a sphere of `radius`.
 * threads <= 0: OpenMP default. */
int odis_mesh_from_file(const char* grid_path, double radius, int32_t threads, odis_mesh** out);
/* Same, from already-parsed arrays (radians). */
int odis_mesh_from_arrays(int32_t n_cells, const double* node_pos_sph, const int32_t* node_friends,
                          const double* centroid_pos_sph, double radius, int32_t threads, odis_mesh** out);
int odis_mesh_get_view(const odis_mesh* mesh, odis_mesh_view* view);
void odis_mesh_free(odis_mesh* mesh);

/* Synthetic icosahedral-bisection grid in the grid_l<L>.txt conventions (the reference ships only
 * levels 3-6; input_files/grid_l7.txt, grid_l8.txt are listed in .MISSING_LARGE_BLOBS). `level`
 * follows the reference's file-name convention: 10*4^(level-1)+2 cells
 * (constants/gridConstants.h:19-32). Outputs are malloc'ed; release with odis_free. */
int odis_grid_generate(int32_t level, int32_t* n_cells_out, double** node_pos_sph_out,
                       int32_t** node_friends_out, double** centroid_pos_sph_out);
int odis_grid_write_file(const char* path, int32_t n_cells, const double* node_pos_sph,
                         const int32_t* node_friends, const double* centroid_pos_sph);
void odis_free(void* p);

/* ------------------------------------------------------------------------------------------
 * solver — replaces the body of ab3Explicit, src/timeIntegrator.cpp:57-322, whose loop (:205-313)
 * calls updateMomentum (src/updateMomentum.cpp:16-47), forcing (src/tidalPotentials.cpp:29-328),
 * the drag/forcing-gradient SpMV (src/timeIntegrator.cpp:219), integrateAB3scalar
 * (src/temporalOperators.cpp:17-68), updateEta (src/updateEta.cpp:7-44), interpolateVelocity
 * (src/interpolation.cpp:26-62) and updateEnergy (src/energy.cpp:13-62).
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_solver odis_solver;

typedef struct odis_params {
    double g;               /* surface gravity                       globals->g            */
    double h;               /* ocean thickness                       globals->h            */
    double alpha;           /* linear drag coefficient               globals->alpha        */
    double dt;              /* quantised time step                   globals->timeStep     */
    double radius;          /* radius after applySurfaceBCs          globals->radius       */
    double omega;           /* spin rate after period rounding       globals->angVel       */
    double love_reduct;     /* tidal potential prefactor             globals->loveReduct   */
    double ecc;             /* eccentricity                          globals->e            */
    double obl;             /* obliquity (rad)                       globals->theta        */
    double shell_thickness; /* added back to r in forcing for LID_*  tidalPotentials.cpp:50-53 */
    double semimajor_axis;  /* PLANET forcing only (required > 0 there) globals->a         */
    int32_t potential;      /* enum Potential, include/globals.h:60-76 */
    int32_t friction;       /* enum Friction (only the diagnostic differs, energy.cpp:46-55) */
    int32_t surface;        /* enum Surface */
    int32_t init_load;      /* 1: AB3 uses the 3-level formula from step 0 (temporalOperators.cpp:36) */
    int32_t reorder;        /* 1: locality (space-filling-curve) renumbering on device; 0: reference order */
    int32_t block_threads;  /* 0 = default */
    int32_t reserved[4];    /* [0] kernel selection. 0 = default: per step one bulk-async staged edge kernel (stencil ids as 16-bit
                             *     offsets from the edge's own id where they fit, 32-bit rows elsewhere) and one staged cell kernel; with
                             *     odis_enable_self_gravity to degree <= 4 the cell kernel also accumulates the harmonic analysis and,
                             *     behind a grid-wide barrier (cooperative launch), solves and adds the term: 2 launches per step, also on
                             *     partitioned solvers. With odis_enable_advection: the nonlinear step in 4 gather launches.
                             *     bit 0: the direct-load baseline kernels (edge, cell, separate analysis / reduce-solve / synthesis
                             *     launches, 6-launch nonlinear step): same fields bit for bit (self-gravity term: within 1e-11).
                             *     bit 3: no CUDA-graph replay. bit 7: 32-bit stencil ids only. bit 8 (tests): 16-bit offset range
                             *     +-1023 instead of +-32767, so that small grids exercise the fallback rows. Rest must be 0. */
} odis_params;

typedef enum odis_field {
    ODIS_FIELD_VELOCITY = 0,      /* v_t0          [F]    normal velocity on edges               */
    ODIS_FIELD_ETA = 1,           /* p_t0          [N]    surface displacement                   */
    ODIS_FIELD_DVDT = 2,          /* dv_dt         [F][3] AB3 history, level 0 newest            */
    ODIS_FIELD_DETADT = 3,        /* dp_dt         [N][3]                                        */
    ODIS_FIELD_VELOCITY_EN = 4,   /* v_avg         [F][2] east, north components at edges        */
    ODIS_FIELD_DISSIPATION = 5,   /* energy_diss   [F]    per-edge dissipated energy flux        */
    ODIS_FIELD_POTENTIAL = 6      /* forcing_potential [N] at the time used by the last step     */
} odis_field;

/* Copies the mesh tables to device `device` (cudaSetDevice ordinal), builds the stencil tables and
 * zero state. */
int odis_create(const odis_mesh_view* mesh, const odis_params* params, int32_t device, odis_solver** out);
/* Multi-GPU: one solver per rank (process or device), each holding a contiguous part of the
 * space-filling-curve cell order plus a one-ring halo. All ranks pass the same global mesh and params.
 * After creation every rank publishes odis_halo_blob_size() bytes with odis_halo_export; the world*size
 * bytes of all ranks, ordered by rank (e.g. an all-gather), go to odis_halo_connect, which maps the
 * neighbours' halo buffers (CUDA IPC across processes, peer access inside one). From then on odis_step
 * exchanges boundary values by direct stores into the neighbours' memory. odis_set_state takes the GLOBAL
 * arrays on every rank; odis_get_field fills this rank's own entries of the global array and zeros
 * elsewhere; the dissipation getters return this rank's partial sum. All calls are collective. */
int odis_create_partitioned(const odis_mesh_view* mesh, const odis_params* params, int32_t device, int32_t rank,
                            int32_t world, odis_solver** out);
int odis_halo_blob_size(void);
int odis_halo_export(odis_solver* s, void* blob_out);
int odis_halo_connect(odis_solver* s, const void* all_blobs);
int odis_get_partition(odis_solver* s, int32_t* rank, int32_t* world, int32_t* own_cells, int32_t* own_edges,
                       int32_t* ghost_cells, int32_t* ghost_edges, int32_t* n_peers);
/* Reference ids of the cells [own_cells] / edges [own_edges] this solver owns, in the order of its device arrays — the order in
 * which a partitioned solver's snapshots (odis_snapshot_wait) hand out their compact arrays. Either pointer may be NULL. */
int odis_get_partition_map(odis_solver* s, int32_t* own_cell_ref_out, int32_t* own_edge_ref_out);
/* The decomposition odis_create_partitioned would use, computed on the host only (no GPU needed): which
 * cells/edges (reference ids) rank `rank` holds — own first, then halo —, its neighbours, and for every
 * neighbour what it sends (reference id + the slot in the neighbour's local numbering). Arrays are
 * malloc'ed; release with odis_partition_plan_free. */
typedef struct odis_partition_plan_t {
    int32_t rank, world;
    int32_t own_cells, own_edges, local_cells, local_edges;
    int32_t n_peers, reserved;
    int32_t* local_cell_ref;    /* [local_cells] */
    int32_t* local_edge_ref;    /* [local_edges] */
    int32_t* peer_rank;         /* [n_peers] ascending */
    int32_t* peer_counts;       /* [n_peers][4]: edges sent, cells sent, edges received, cells received */
    int32_t* send_edge_ref;     /* concatenated over peers in peer order */
    int32_t* send_edge_slot;
    int32_t* send_cell_ref;
    int32_t* send_cell_slot;
} odis_partition_plan_t;
int odis_partition_plan(const odis_mesh_view* mesh, int32_t reorder, int32_t rank, int32_t world, odis_partition_plan_t* plan_out);
void odis_partition_plan_free(odis_partition_plan_t* plan);

/* `initial conditions; ANALYTICAL` (host only): the state getInitialConditions hands to the loop from analyticalInitialConditions
 * (src/initialConditions.cpp:146-208) / analyticalLTE (src/analyticalLTE.cpp:47-181) — the two-mode analytical response to the westward
 * obliquity tide at t = 0 with tendencies at 0, -dt, -2dt: v[F], dvdt[F][3], eta[N], detadt[N][3], ready for odis_set_state(..., 0)
 * (params.init_load stays 0: the reference starts such runs with its Euler / two-level steps, src/temporalOperators.cpp:36).
 * params.potential must be OBLIQ_WEST, the only type the reference has a solution for. Bit-identical to the reference's arrays. */
int odis_analytical_state(const odis_mesh_view* mesh, const odis_params* params, double* v, double* dvdt, double* eta, double* detadt);
/* Host -> device state in reference numbering. NULL pointers mean zeros. `iter` is the number of
 * steps already taken (current_time = dt*iter, src/timeIntegrator.cpp:187,277). */
int odis_set_state(odis_solver* s, const double* v, const double* eta, const double* dvdt /*[F][3]*/,
                   const double* detadt /*[N][3]*/, int64_t iter);
/* Advance nsteps time steps (asynchronous on the solver's stream; odis_get_* synchronise).
 * Partitioned solvers: the steps in flight wait INSIDE the kernels for the neighbours' flags (bounded: ~10 s, ODIS_B200_WAIT_TIMEOUT_S),
 * so every rank must take the same steps, and the stream must be drained (odis_synchronize) before the caller runs anything ELSE
 * that waits on another GPU on the same device — e.g. an NCCL collective: its kernel, resident beside a step that waits for a peer
 * whose own collective is queued behind ITS steps, closes a cycle that only the time limit ends (reported as ODIS_ERR_STATE). */
int odis_step(odis_solver* s, int32_t nsteps);
/* As odis_step, bracketed by CUDA events on the solver's stream; returns elapsed device ms. */
int odis_step_timed(odis_solver* s, int32_t nsteps, float* elapsed_ms_out);
/* As odis_step, with every kernel launch bracketed by its own CUDA event pair; returns the summed device
 * time (ms) of the edge-update and of the cell-update launches separately (roofline instrumentation). */
int odis_step_profiled(odis_solver* s, int32_t nsteps, float* edge_ms_out, float* cell_ms_out);
/* Nonlinear branch of the step (`advection; true`): calculateMomentumAdvection (src/momAdvection.cpp:11-268: potential-vorticity
 * flux over the TRiSK stencil + gradient of the kinetic energy from operatorRBFinterp) replaces the Coriolis product in
 * updateMomentum (src/updateMomentum.cpp:37-38), and updateEta takes the divergence of the third-order edge flux of
 * interpolateLSQFlux (src/updateEta.cpp:32-33, src/interpolation.cpp:311-364). The three operators only this branch uses are
 * handed over as the reference builds them — row-major CSR, columns ascending (Eigen's compressed storage: outerIndexPtr /
 * innerIndexPtr / valuePtr of Mesh::operatorCurl, ::operatorRBFinterp, ::operatorDirectionalSecondDeriv, src/mesh.cpp:3122-3175,
 * 2263-2361, 2364-2719) — with Mesh::vertex_sinlat and ::vertex_area; `mesh` must carry vertex_nodes, vertex_R and face_vertexes.
 * Fields then match the reference's nonlinear solver bit for bit. Single, unpartitioned solvers only. */
typedef struct odis_csr_view {
    int32_t n_rows, n_cols;
    const int32_t* indptr;   /* [n_rows + 1] */
    const int32_t* indices;  /* [nnz] */
    const double* data;      /* [nnz] */
} odis_csr_view;
typedef struct odis_nonlinear_view {
    odis_csr_view curl;                      /* V x F */
    odis_csr_view rbf_interp;                /* 3N x F, rows 3i, 3i+1, 3i+2 = x, y, z at cell i */
    odis_csr_view directional_second_deriv;  /* 2F x N, rows 2e, 2e+1 = inner, outer cell side of edge e */
    const double* vertex_sinlat;             /* [V] */
    const double* vertex_area;               /* [V] */
} odis_nonlinear_view;
int odis_enable_advection(odis_solver* s, const odis_mesh_view* mesh, const odis_nonlinear_view* nl);
/* The same tables built by the library from its own mesh (host only): the vertex part of Mesh::AssignFaces
 * (src/mesh.cpp:910-1147), CalcControlVolumeInterpMatrix (:199-434), CalcRBFInterpMatrix (:2263-2361), CalcRBFInterpMatrix2
 * (:2364-2719), CalcCurlOperatorCoeffs (:3122-3175). rbf_eps = input.in's "rbf epsilon". The view stays valid until
 * odis_nonlinear_free. */
typedef struct odis_nonlinear odis_nonlinear;
int odis_nonlinear_create(const odis_mesh* mesh, double rbf_eps, odis_nonlinear** out);
int odis_nonlinear_get_view(const odis_nonlinear* nl, odis_nonlinear_view* view);
void odis_nonlinear_free(odis_nonlinear* nl);
/* Self-gravity / shell-pressure term by spherical harmonics — the reference's pressureGradientSH
 * (src/spatialOperators.cpp:387-462, commented out at HEAD) with the basis of Mesh::CalcLegendreFuncs
 * (src/mesh.cpp:2154-2260) and the least-squares coefficients of getSHCoeffsGG (src/sphericalHarmonics.cpp:16-72 ->
 * src/extractSHCoeffGG.f95). From this call on, every step adds
 *     g * sum_{l=2..l_max} factor[l] * sum_m (C_lm cos(m lon) + S_lm sin(m lon)) Pbar_lm(cos colat)
 * to forcing_potential, C_lm/S_lm being the least-squares (degrees 0..l_max) coefficients of eta at the start of
 * the step. factor[l] = globals->shell_factor_beta[l] (= 1 - beta_l, boundaryConditions.cpp:373) for the LID_*
 * surfaces, globals->loading_factor[l] for FREE_LOADING; 4-pi normalised harmonics with the Condon-Shortley phase
 * (src/legendre.f95). On the device: a dense matrix-vector product per direction (analysis, synthesis) —
 * stored_basis = 1: the basis matrix lives in HBM and is streamed twice per step (the reference's dgemv on
 * sh_matrix_fort); stored_basis = 0 (recommended): matrix-free, the basis values of a cell are rebuilt from its
 * latitude / longitude by recurrence inside both kernels, which moves 32 B per cell instead of 8 (l_max+1)^2.
 * Call it after odis_create, before or after odis_set_state; `mesh` is the mesh the solver was created from. */
int odis_enable_self_gravity(odis_solver* s, const odis_mesh_view* mesh, int32_t l_max, const double* factor /*[l_max+1]*/,
                             int32_t stored_basis);
/* Least-squares coefficients of the eta the last potential was built from: [(l_max+1)^2], degree-major, per degree
 * m = 0, then (cos, sin) for m = 1..l. */
int odis_get_sh_coefficients(odis_solver* s, double* out);
/* As odis_step_profiled, with the self-gravity launches timed separately. */
int odis_step_profiled_sh(odis_solver* s, int32_t nsteps, float* edge_ms_out, float* cell_ms_out, float* sh_ms_out);
/* Host-only helpers (no GPU): the basis rows Y[(l_max+1)^2][n] at the given (lat, lon) [n][2] in radians, and the
 * inverse normal matrix (Y Y^T)^-1 [(l_max+1)^2][(l_max+1)^2] of the least-squares fit over those points. */
int odis_sh_basis(int32_t n, const double* pos_sph, int32_t l_max, double* Y_out);
int odis_sh_normal_inverse(int32_t n, const double* pos_sph, int32_t l_max, double* Ginv_out);
/* Device -> host, reference numbering and layout. */
int odis_get_field(odis_solver* s, int32_t field, double* out);
/* Area-mean dissipated energy flux after the last step (e_diss of updateEnergy, energy.cpp:60). */
int odis_get_dissipation_avg(odis_solver* s, double* out);
/* Per-step series of the same quantity counted from the last odis_set_state: entry j (first <= j <
 * first+count) is the value for the state after j steps, j = 0 being the state as set. */
int odis_get_dissipation_series(odis_solver* s, int64_t first, int64_t count, double* out);
/* Forgets the series before the current step: entry 0 becomes the current state's, the steps counted since odis_set_state restart at 0.
 * Whole runs (odis_run) call it at every output interval — the reference keeps no per-step series at all (energy.cpp:13-62 overwrites
 * one scalar) — so that the series does not grow with the length of the run. */
int odis_trim_dissipation_series(odis_solver* s);
/* operators — the free functions the reference's loop calls (src/timeIntegrator.cpp:205-313), one call each, for a caller
 * that keeps the reference's own ab3Explicit and swaps single functions (integration/operators_b200.cpp holds the wrappers
 * with the reference's C++ signatures). Host arrays, reference numbering; each call copies its arguments to the device, runs
 * the same kernels odis_step runs (so the arithmetic is identical) and copies the result back. Linear branch
 * (`advection; false`), unpartitioned solver. The solver's device state is used as scratch space: after any odis_op_* call
 * odis_step / odis_get_* return ODIS_ERR_STATE until odis_set_state is called again.
 *   odis_op_update_momentum      updateMomentum, src/updateMomentum.cpp:16-47 (:42): dvdt = -g G eta + C v           [F]
 *   odis_op_update_eta           updateEta, src/updateEta.cpp:7-44 (:39): detadt = h Div v                             [N]
 *   odis_op_forcing              forcing, src/tidalPotentials.cpp:29-328: tidal potential at `time` (the solver's potential type) [N]
 *   odis_op_integrate_ab3_scalar integrateAB3scalar, src/temporalOperators.cpp:17-68: solution[n] and dsolution_dt[n][3] in place,
 *                                n <= number of edges; `iter` selects the start-up formulas (:36,:49,:59)
 *   odis_op_interpolate_velocity interpolateVelocity, src/interpolation.cpp:26-62: east/north components                [F][2]
 *   odis_op_update_energy        updateEnergy, src/energy.cpp:13-62: e_flux [F] and the area-mean flux from v_avg [F][2] and
 *                                the edge areas [F] (grid->face_area) */
int odis_op_update_momentum(odis_solver* s, const double* v, const double* eta, double* dvdt_out);
int odis_op_update_eta(odis_solver* s, const double* v, double* detadt_out);
int odis_op_forcing(odis_solver* s, double time, double* potential_out);
int odis_op_integrate_ab3_scalar(odis_solver* s, double* solution, double* dsolution_dt, int64_t iter, int32_t n);
int odis_op_interpolate_velocity(odis_solver* s, const double* v, double* v_avg_out);
int odis_op_update_energy(odis_solver* s, const double* v_avg, const double* areas, double* e_flux_out, double* avg_flux_out);
/* Output snapshots that overlap with stepping — the DumpData side of the loop (src/timeIntegrator.cpp:280-304, src/outFiles.cpp:522-684)
 * without stalling it. odis_snapshot_begin enqueues, behind the steps taken so far, the diagnostics and the device -> host copy of the
 * requested fields (ODIS_SNAP_* bits; the dissipation average always) into page-locked memory owned by the library, on a second
 * stream, and returns; the caller enqueues the next interval with odis_step and then calls odis_snapshot_wait, which blocks only until
 * that copy has landed and hands out pointers valid until the slot's next odis_snapshot_begin. Two slots (0, 1) for double buffering.
 * Partitioned solvers: every array holds the rank's OWN entries only ([own_cells], [own_edges][2], ...), compact, in the order of
 * odis_get_partition_map (only those bytes cross PCIe); dissipation_avg is the rank's partial sum over the sphere's area, as from
 * odis_get_dissipation_avg. */
#define ODIS_SNAP_ETA 1u          /* p_t0        [N]    */
#define ODIS_SNAP_VELOCITY_EN 2u  /* v_avg       [F][2] */
#define ODIS_SNAP_DISSIPATION 4u  /* energy_diss [F]    */
#define ODIS_SNAP_VELOCITY 8u     /* v_t0        [F]    */
typedef struct odis_snapshot_view {
    const double* eta;
    const double* velocity_en;
    const double* dissipation;
    const double* velocity;        /* NULL for fields that were not requested */
    double dissipation_avg;        /* as odis_get_dissipation_avg */
    int64_t iter;                  /* steps taken when the snapshot was begun */
} odis_snapshot_view;
int odis_snapshot_begin(odis_solver* s, int32_t slot, uint32_t fields);
int odis_snapshot_wait(odis_solver* s, int32_t slot, odis_snapshot_view* out);
/* The input side of the same pipeline: the NEXT state (a restart, the next case of a sweep, coupled-model input — arrays as in
 * odis_set_state, the reference's getInitialConditions / loadInitialConditions arrays, src/initialConditions.cpp:19-144) travels to the
 * device on the second stream while the current interval is still stepping. odis_stage_state returns at once; the host arrays must stay
 * unchanged until odis_commit_state (page-locked memory makes the copies asynchronous). odis_commit_state makes the staged arrays the
 * solver's state behind the steps enqueued so far (same renumbering launches and first potential as odis_set_state) without waiting on
 * the host. One staged state at a time. Partitioned solvers take the GLOBAL arrays (as odis_set_state does), pack their own share on
 * the host inside odis_stage_state — while the device steps; the arrays are free again when the call returns — and upload only that
 * share; every rank calls both functions (collective, like odis_set_state). */
int odis_stage_state(odis_solver* s, const double* v, const double* eta, const double* dvdt /*[F][3]*/, const double* detadt /*[N][3]*/);
int odis_commit_state(odis_solver* s, int64_t iter);
int odis_get_iter(odis_solver* s, int64_t* iter_out);
/* Bytes of device memory held, and the algorithmic HBM bytes one step moves. */
int odis_get_footprint(odis_solver* s, int64_t* device_bytes_out, int64_t* algorithmic_bytes_per_step_out);
/* Number of kernel launches issued by this solver since creation. */
int odis_get_launch_count(odis_solver* s, int64_t* launches_out);
int odis_synchronize(odis_solver* s);
void odis_destroy(odis_solver* s);

/* ------------------------------------------------------------------------------------------
 * ensemble — M independent runs on one grid, advanced together (parameter sweeps over ocean thickness, drag, ...).
 * The reference has no such mode: a sweep is M separate `./ODIS` runs, i.e. M times the loop of ab3Explicit
 * (src/timeIntegrator.cpp:205-313) with a different input.in each. Members share the mesh, dt, omega, radius, shell
 * thickness, potential / friction / surface type and init_load; g, h, alpha, love_reduct, ecc, obl may differ. Every
 * member is bit-identical to an odis_solver run with the same odis_params. Fields cross in reference numbering.
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_ensemble odis_ensemble;
int odis_ensemble_create(const odis_mesh_view* mesh, const odis_params* params /*[n_members]*/, int32_t n_members, int32_t device,
                         odis_ensemble** out);
/* State of one member (member = -1: the same state for every member); NULL pointers mean zeros. `iter` (shared by all
 * members) restarts the step count; call it for the members before stepping. */
int odis_ensemble_set_state(odis_ensemble* e, int32_t member, const double* v, const double* eta, const double* dvdt /*[F][3]*/,
                            const double* detadt /*[N][3]*/, int64_t iter);
/* Self-gravity / shell-pressure term for every member (see odis_enable_self_gravity; the members' own g applies, the
 * per-degree factors are shared). The harmonic analysis and synthesis of all members are one FP64 tensor-core GEMM each
 * (basis matrix x member-innermost state). l_max <= 10. */
int odis_ensemble_enable_self_gravity(odis_ensemble* e, const odis_mesh_view* mesh, int32_t l_max, const double* factor /*[l_max+1]*/);
int odis_ensemble_get_sh_coefficients(odis_ensemble* e, int32_t member, double* out /*[(l_max+1)^2]*/);
int odis_ensemble_step(odis_ensemble* e, int32_t nsteps);
int odis_ensemble_step_timed(odis_ensemble* e, int32_t nsteps, float* elapsed_ms_out);
/* field: ODIS_FIELD_VELOCITY, _ETA, _DVDT, _DETADT or _POTENTIAL of one member. */
int odis_ensemble_get_field(odis_ensemble* e, int32_t member, int32_t field, double* out);
/* as odis_get_dissipation_series, for one member */
int odis_ensemble_get_dissipation_series(odis_ensemble* e, int32_t member, int64_t first, int64_t count, double* out);
/* any pointer may be NULL; algorithmic bytes per batched step = M*(40F + 56N) + 248F + 160N */
int odis_ensemble_get_info(odis_ensemble* e, int32_t* n_members, int64_t* iter, int64_t* launches, int64_t* device_bytes,
                           int64_t* algorithmic_bytes_per_step);
void odis_ensemble_destroy(odis_ensemble* e);

/* ------------------------------------------------------------------------------------------
 * output — replaces the HDF5 side of OutFiles: CreateHDF5Framework (src/outFiles.cpp:138-462: H5Fcreate +
 * one H5Dcreate(H5T_NATIVE_FLOAT, contiguous, fixed shape) per enabled field), DumpGridData (:464-520) and
 * DumpData's H5Sselect_hyperslab/H5Dwrite of one row per dump (:522-684). Writes the HDF5 file format
 * directly (superblock v0, v1 object headers, contiguous f32 datasets); declare all datasets, then write.
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_h5 odis_h5;
int odis_h5_create(const char* path, odis_h5** out);
/* float32 dataset of rank 1 or 2 with fixed dims; fails on a duplicate name (as H5Dcreate does). */
int odis_h5_add_dataset(odis_h5* h, const char* name, int32_t rank, const uint64_t* dims, int32_t* id_out);
/* rows [first_row, first_row+nrows) of a rank-2 dataset (row = dims[1] floats) or elements of a rank-1
 * dataset; fails when the selection leaves the dataset extent (as H5Sselect_hyperslab does). */
int odis_h5_write_rows(odis_h5* h, int32_t dataset, uint64_t first_row, uint64_t nrows, const float* data);
/* Flushes the metadata, closes the file and frees the handle. */
int odis_h5_close(odis_h5* h);

/* ------------------------------------------------------------------------------------------
 * whole run — replaces `./ODIS` started in a run directory: main() (src/main.cpp:46-68) -> solveODIS
 * (src/solver.cpp:15-61) -> ab3Explicit (src/timeIntegrator.cpp:57-322) with its output cadence
 * (:188,280), log line (:296-299), SIGINT handling (:32-37,120,307-312) and restart files (:316).
 * Reads <run_dir>/input.in, input_files/grid_l<L>.txt [, InitialConditions/]; writes DATA/OUTPUT.txt,
 * DATA/ERROR.txt, DATA/data.h5, InitialConditions/{vel,pres}_init.txt.
 * ---------------------------------------------------------------------------------------- */
typedef struct odis_run_options {
    int32_t device;       /* CUDA device ordinal */
    int32_t reorder;      /* as odis_params.reorder (default 1 when options == NULL) */
    int32_t echo;         /* 1: copy OUTPUT.txt lines to stdout */
    int32_t self_gravity; /* 0: as reference HEAD (term commented out). 1: odis_enable_self_gravity with input.in's "sh degree" and
                           * the surface type's per-degree factors (matrix-free kernels); 2: the same with the stored basis */
    int64_t max_steps;    /* > 0: stop after this many steps even if the loop bound is larger */
    int32_t overlap_output; /* 1: dumps go through odis_snapshot_begin/_wait: the next output interval is computed while the previous
                             * dump is copied out and written to data.h5 (same datasets, log lines and restart files). 0: synchronous dumps */
    int32_t n_gpus;       /* 0 or 1: one GPU. N > 1: the grid is cut into N space-filling-curve parts, one partitioned solver per GPU
                           * (devices device .. device + N - 1) driven from this one process; halos and harmonic sums are exchanged by the
                           * step kernels through peer memory. Same files as the one-GPU run (bit-identical without the self-gravity
                           * term), also with overlap_output (every rank's own entries come back compact and are placed by its partition
                           * map). Not with `advection; true`. */
} odis_run_options;

typedef struct odis_run_result {
    int64_t steps;              /* time steps taken */
    int64_t kernel_launches;
    int32_t dumps;              /* rows written to data.h5 (including the initial state) */
    int32_t interrupted;        /* 1 if SIGINT ended the run (ab3Explicit's return value) */
    int32_t n_cells, n_edges;
    int32_t steps_per_period;   /* totalIter */
    int32_t reserved;
    double dt;
    double last_dissipation_avg;
} odis_run_result;

int odis_run(const char* run_dir, const odis_run_options* options, odis_run_result* result_out);

#ifdef __cplusplus
}
#endif
#endif /* ODIS_B200_H */
