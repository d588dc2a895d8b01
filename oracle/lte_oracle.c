/* TEST INFRASTRUCTURE — CPU oracle for the LTE time step. NOT product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * A plain-C restatement of the reference's per-step arithmetic, function by function, in the
 * reference's own evaluation order (Eigen row-major CSR products: columns ascending, per-row
 * accumulator started at 0, scalar*coefficient formed before the multiply; see
 * oracle/ref_build/shim/Eigen/ShimCore.h for the Eigen semantics restated). Each routine names the
 * reference lines it follows (paths under /root/reference). Parity is PINNED: tests/test_oracle_pinned.py
 * checks this file bit-for-bit against the reference's own solver (oracle/_ref, the unmodified
 * reference sources built by oracle/ref_build/Makefile) through the committed fixtures in
 * tests/golden/ and, when /root/reference is present, against live runs.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build_oracle.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_cells, n_edges;
    const int* node_friends;                  /* [N][6] */
    const int* faces;                         /* [N][6] */
    const int* node_face_dir;                 /* [N][6] */
    const int* face_nodes;                    /* [F][2] */
    const int* face_interp_friends;           /* [F][10] */
    const double* face_interp_weights;        /* [F][10] */
    const double* face_len;                   /* [F] */
    const double* face_node_dist;             /* [F] */
    const double* face_centre_m;              /* [F][2] */
    const double* face_centre_pos_sph;        /* [F][2] */
    const double* face_area;                  /* [F] */
    const double* face_normal_vec_map;        /* [F][2] */
    const double* control_volume_surf_area_map; /* [N] */
    const double* node_pos_sph;               /* [N][2] */
} oracle_mesh;

typedef struct {
    double g, h, alpha, dt, radius, omega, love_reduct, ecc, obl, shell_thickness, semimajor_axis;
    int potential, friction, surface, init_load;
} oracle_params;

typedef struct { int nr, nc; int* ptr; int* idx; double* val; } csr;

typedef struct {
    int N, F;
    oracle_mesh m;
    oracle_params p;
    csr grad, div, cor, drag;
    double *trigLat, *trigLon, *trig2Lat, *trig2Lon, *trigSqLat;   /* [N][2] each, mesh.cpp:2132-2145 */
    /* state, names of timeIntegrator.cpp:70-102 */
    double *v_t0, *p_t0, *dv_dt, *dp_dt, *dv_dt_t0, *dp_dt_t0, *drag_term, *forcing_potential, *v_avg, *energy_diss;
    double *tmp_f, *tmp_n;
    long iter;
    double e_diss;
    /* optional spherical-harmonic self-gravity term (oracle/sh_oracle.py builds the matrices) */
    int sh_rows;
    const double* sh_Y;        /* [rows][N] basis */
    const double* sh_T;        /* [rows][rows] factor_l * (Y Y^T)^-1, rows of degree 0 and 1 zero */
    double *sh_b, *sh_s;
    /* optional nonlinear branch (advection; true): operators handed in as CSR, see oracle_set_nonlinear */
    int nl_on, V;
    csr curl, rbf, d2;                  /* operatorCurl VxF, operatorRBFinterp 3NxF, operatorDirectionalSecondDeriv 2FxN */
    const double *vertex_sinlat, *vertex_area, *vertex_R;   /* [V], [V], [V][3] */
    const int *vertex_nodes, *face_vertexes;                /* [V][3], [F][2] */
    double *h_total, *ekin, *vorticity_v, *vorticity_e, *thickness_e, *vel_xyz, *d2v, *flux;
} oracle_ctx;

static const double pi = 3.1415926535897932384626433832795028841971693993751058;  /* mathRoutines.h:10 */

/* ---- Eigen::SparseMatrix<double,RowMajor>::setFromTriplets: rows, then columns ascending,
 *      duplicates summed (none occur here) ---- */
typedef struct { int r, c; double v; long seq; } trip;
static int trip_cmp(const void* a, const void* b) {
    const trip* x = (const trip*)a; const trip* y = (const trip*)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static void csr_from_triplets(csr* A, int nr, int nc, trip* t, long n) {
    long k;
    for (k = 0; k < n; k++) t[k].seq = k;
    qsort(t, (size_t)n, sizeof(trip), trip_cmp);
    A->nr = nr; A->nc = nc;
    A->ptr = (int*)calloc((size_t)nr + 1, sizeof(int));
    A->idx = (int*)malloc((size_t)n * sizeof(int));
    A->val = (double*)malloc((size_t)n * sizeof(double));
    long nnz = 0; int lr = -1, lc = -1;
    for (k = 0; k < n; k++) {
        if (t[k].r == lr && t[k].c == lc) { A->val[nnz - 1] += t[k].v; continue; }
        A->idx[nnz] = t[k].c; A->val[nnz] = t[k].v; nnz++;
        A->ptr[t[k].r + 1]++;
        lr = t[k].r; lc = t[k].c;
    }
    for (k = 0; k < nr; k++) A->ptr[k + 1] += A->ptr[k];
}
static void csr_free(csr* A) { free(A->ptr); free(A->idx); free(A->val); }

/* dst = 0; dst_i += 1.0 * sum_j (s*a_ij) * x_j   — "dst = (s*A)*x" in Eigen */
static void spmv_scaled_assign(const csr* A, double s, const double* x, double* dst) {
    int i, k;
    for (i = 0; i < A->nr; i++) {
        double tmp = 0;
        for (k = A->ptr[i]; k < A->ptr[i + 1]; k++) tmp += (s * A->val[k]) * x[A->idx[k]];
        dst[i] = 0.0;
        dst[i] += 1.0 * tmp;
    }
}
static void spmv_assign(const csr* A, const double* x, double* dst) {
    int i, k;
    for (i = 0; i < A->nr; i++) {
        double tmp = 0;
        for (k = A->ptr[i]; k < A->ptr[i + 1]; k++) tmp += A->val[k] * x[A->idx[k]];
        dst[i] = 0.0;
        dst[i] += 1.0 * tmp;
    }
}
static void spmv_add(const csr* A, const double* x, double* dst) {
    int i, k;
    for (i = 0; i < A->nr; i++) {
        double tmp = 0;
        for (k = A->ptr[i]; k < A->ptr[i + 1]; k++) tmp += A->val[k] * x[A->idx[k]];
        dst[i] += 1.0 * tmp;
    }
}

/* ---- operator assembly ---- */
static void build_operators(oracle_ctx* c) {
    const oracle_mesh* m = &c->m;
    const int N = c->N, F = c->F;
    int i, j;
    long n;
    /* Mesh::CalcGradOperatorCoeffs, mesh.cpp:3059-3084 */
    trip* t = (trip*)malloc((size_t)F * 10 * sizeof(trip));
    n = 0;
    for (i = 0; i < F; i++) {
        const double dist = m->face_node_dist[i];
        t[n].r = i; t[n].c = m->face_nodes[i * 2 + 0]; t[n].v = (-m->face_centre_m[i * 2 + 0]) / dist; n++;
        t[n].r = i; t[n].c = m->face_nodes[i * 2 + 1]; t[n].v = (m->face_centre_m[i * 2 + 1]) / dist; n++;
    }
    csr_from_triplets(&c->grad, F, N, t, n);
    /* Mesh::CalcDivOperatorCoeffs, mesh.cpp:3228-3254 */
    n = 0;
    for (i = 0; i < N; i++) {
        const int f_num = (m->node_friends[i * 6 + 5] == -1) ? 5 : 6;
        for (j = 0; j < f_num; j++) {
            const int face_id = m->faces[i * 6 + j];
            const double edge_len = m->face_len[face_id];
            const int dir = m->node_face_dir[i * 6 + j];
            const double area = m->control_volume_surf_area_map[i];
            t[n].r = i; t[n].c = face_id; t[n].v = -dir * edge_len / area; n++;
        }
    }
    csr_from_triplets(&c->div, N, F, t, n);
    /* Mesh::CalcCoriolisOperatorCoeffs, mesh.cpp:2863-2898 */
    n = 0;
    for (i = 0; i < F; i++) {
        int friend_num = 10;
        const int n1 = m->face_nodes[i * 2 + 0], n2 = m->face_nodes[i * 2 + 1];
        if (m->node_friends[n1 * 6 + 5] < 0) friend_num--;
        if (m->node_friends[n2 * 6 + 5] < 0) friend_num--;
        for (j = 0; j < friend_num; j++) {
            const int f_ID = m->face_interp_friends[i * 10 + j];
            t[n].r = i; t[n].c = f_ID;
            t[n].v = -2.0 * c->p.omega * sin(m->face_centre_pos_sph[i * 2 + 0]) * m->face_interp_weights[i * 10 + j] *
                     m->face_len[f_ID] / m->face_node_dist[i];
            n++;
        }
    }
    csr_from_triplets(&c->cor, F, F, t, n);
    /* Mesh::CalcLinearDragOperatorCoeffs, mesh.cpp:2957-2963 */
    n = 0;
    for (i = 0; i < F; i++) { t[n].r = i; t[n].c = i; t[n].v = -c->p.alpha; n++; }
    csr_from_triplets(&c->drag, F, F, t, n);
    free(t);
}

/* Mesh::CalcTrigFunctions, mesh.cpp:2127-2149 */
static void build_trig(oracle_ctx* c) {
    const int N = c->N;
    int i;
    c->trigLat = (double*)malloc((size_t)N * 2 * sizeof(double));
    c->trigLon = (double*)malloc((size_t)N * 2 * sizeof(double));
    c->trig2Lat = (double*)malloc((size_t)N * 2 * sizeof(double));
    c->trig2Lon = (double*)malloc((size_t)N * 2 * sizeof(double));
    c->trigSqLat = (double*)malloc((size_t)N * 2 * sizeof(double));
    for (i = 0; i < N; i++) {
        const double lat = c->m.node_pos_sph[i * 2], lon = c->m.node_pos_sph[i * 2 + 1];
        c->trigLat[i * 2] = cos(lat);        c->trigLat[i * 2 + 1] = sin(lat);
        c->trigLon[i * 2] = cos(lon);        c->trigLon[i * 2 + 1] = sin(lon);
        c->trig2Lat[i * 2] = cos(2.0 * lat); c->trig2Lat[i * 2 + 1] = sin(2.0 * lat);
        c->trig2Lon[i * 2] = cos(2.0 * lon); c->trig2Lon[i * 2 + 1] = sin(2.0 * lon);
        c->trigSqLat[i * 2] = cos(lat) * cos(lat);
        c->trigSqLat[i * 2 + 1] = sin(lat) * sin(lat);
    }
}

/* forcing(), tidalPotentials.cpp:29-328 (ECC :84-100, OBLIQ :106-113, OBLIQ_WEST :120-128,
 * FULL2 :135-154, FULL :160-172, NONE :283). enum values: include/globals.h:60-76 */
static void forcing(oracle_ctx* c, double* potential, double time) {
    const int N = c->N;
    const oracle_params* p = &c->p;
    double radius = p->radius, omega = p->omega, ecc = p->ecc, obl = p->obl, factor, factor2;
    int i, j;
    if (p->surface == 2 || p->surface == 3) radius += p->shell_thickness;      /* :50-53 */
    const double cosM = cos(omega * time), sinM = sin(omega * time);
    const double cos2M = cos(2 * omega * time);
    const double cos3M = cos(3 * omega * time), cos4M = cos(4 * omega * time);
    const double *cosLon = c->trigLon, *sinLon = c->trigLon + 1, *cosLat = c->trigLat, *sinLat = c->trigLat + 1;
    const double *cos2Lon = c->trig2Lon, *sin2Lon = c->trig2Lon + 1, *sin2Lat = c->trig2Lat + 1, *cos2Lat = c->trig2Lat;
    const double *cosSqLat = c->trigSqLat, *sinSqLat = c->trigSqLat + 1;
    switch (p->potential) {
        case 5: /* ECC */
            factor = 0.75 * p->love_reduct * (omega * omega) * (radius * radius) * ecc;
            for (i = 0; i < N; ++i)
                potential[i] = factor * ((1. - 3. * sinSqLat[i * 2]) * cosM +
                                         cosSqLat[i * 2] * (3. * cosM * cos2Lon[i * 2] + 4. * sinM * sin2Lon[i * 2]));
            break;
        case 0: /* OBLIQ */
            factor = -3. / 2. * p->love_reduct * (omega * omega) * (radius * radius) * obl;
            for (i = 0; i < N; ++i) { j = i * 2; potential[i] = factor * cosM * sin2Lat[j] * cosLon[j]; }
            break;
        case 1: /* OBLIQ_WEST */
            factor = 0.5 * p->love_reduct * (omega * omega) * (radius * radius) * obl;
            for (i = 0; i < N; ++i) { j = i * 2; potential[i] = 3 * factor * sinLat[j] * cosLat[j] * (cosLon[j] * cosM - sinLon[j] * sinM); }
            break;
        case 9: /* FULL2 */
            factor = 1 / 32. * p->love_reduct * (omega * omega) * (radius * radius);
            for (i = 0; i < N; i++) {
                double T1, T2, T3;
                j = i * 2;
                T1 = 3. * ecc * (4. - 7. * obl * obl) * cosM + 6 * (obl * obl + ecc * ecc * (3 - 7 * obl * obl)) * cos2M;
                T1 += 3 * ecc * obl * obl * (7 * cos3M + 17 * ecc * cos4M);
                T1 *= -(1 - 3 * cos2Lat[j]);
                T2 = (4 + 15 * ecc * ecc + 20 * ecc * cosM + 43 * ecc * ecc * cos2M) * cosLon[j];
                T2 += 2 * ecc * (4 + 25 * ecc * cosM) * sinM * sinLon[j];
                T2 *= 24 * obl * cosLat[j] * sinLat[j] * sinM;
                T3 = obl * obl * (2 + 3 * ecc * ecc + 6 * ecc * cosM + 9 * ecc * ecc * cos2M) * (cosM * cosLon[j] + sinM * sinLon[j]);
                T3 += -(obl * obl - 2) * ((6 * ecc * cosM + 17 * ecc * ecc * cos2M) * cos2Lon[j] + 2 * ecc * (4 + 17 * ecc * cosM) * sinM * sin2Lon[j]);
                T3 *= 6 * cosSqLat[j];
                potential[i] = factor * (T1 + T2 + T3);
            }
            break;
        case 8: /* FULL */
            factor = 0.75 * p->love_reduct * (omega * omega) * (radius * radius) * ecc;
            factor2 = -3. / 2. * p->love_reduct * (omega * omega) * (radius * radius) * obl;
            for (i = 0; i < N; ++i) {
                j = i * 2;
                potential[i] = factor * ((1 - 3 * sinSqLat[j]) * cosM + cosSqLat[j] * (3 * cosM * cos2Lon[j] + 4 * sinM * sin2Lon[j])) +
                               factor2 * cosM * sin2Lat[j] * cosLon[j];
            }
            break;
        case 13: { /* PLANET, tidalPotentials.cpp:176-225: a companion moon on the inner 2:1 orbit (hard-wired mass and orbit of Io) */
            const double m2 = 8.931938e+22, a1 = 421800000.0, a2 = p->semimajor_axis;
            const double n2 = omega, n1 = n2 * 2.0, nij = (n1 - n2);
            const double cosnt = cos(nij * time), sinnt = sin(nij * time);
            const double pp = pow(a1, 2.0) + pow(a2, 2.0) - 2. * a1 * a2 * cosnt;
            const double cosphi = (a1 - a2 * cosnt), sinphi = a2 * sinnt;
            factor = 0.5 * 6.67408e-11 * m2 * pow(radius / pp, 2.0) / sqrt(pp);
            for (i = 0; i < N; ++i) {
                j = i * 2;
                const double cosgam = cosLat[j] * (cosLon[j] * cosphi + sinLon[j] * sinphi);
                potential[i] = factor * (3. * pow(cosgam, 2.0) - pp);
            }
            break;
        }
        default: /* NONE */
            break;
    }
}

/* integrateAB3scalar, temporalOperators.cpp:17-68 */
static void integrateAB3scalar(const oracle_ctx* c, double* s, double* ds_dt, long iter, int num) {
    const double dt = c->p.dt;
    const double a = 23. / 12., b = -16. / 12., cc = 5. / 12.;
    int i;
    if ((iter > 1) || c->p.init_load) {
        for (i = 0; i < num; ++i) {
            s[i] += (a * ds_dt[i * 3] + b * ds_dt[i * 3 + 1] + cc * ds_dt[i * 3 + 2]) * dt;
            ds_dt[i * 3 + 2] = ds_dt[i * 3 + 1];
            ds_dt[i * 3 + 1] = ds_dt[i * 3];
        }
    } else if (iter == 0) {
        for (i = 0; i < num; i++) { s[i] += ds_dt[i * 3] * dt; ds_dt[i * 3 + 2] = ds_dt[i * 3]; }
    } else if (iter == 1) {
        for (i = 0; i < num; i++) { s[i] += ds_dt[i * 3] * dt; ds_dt[i * 3 + 1] = ds_dt[i * 3]; }
    }
}

/* interpolateVelocity, interpolation.cpp:26-62 */
static void interpolateVelocity(const oracle_ctx* c, double* interp_vel, const double* normal_vel) {
    const oracle_mesh* m = &c->m;
    int i, j;
    for (i = 0; i < c->F; ++i) {
        double v_tang = 0.0, v_norm;
        int friend_num = 10;
        const int n1 = m->face_nodes[i * 2], n2 = m->face_nodes[i * 2 + 1];
        if (m->node_friends[n1 * 6 + 5] < 0) friend_num--;
        if (m->node_friends[n2 * 6 + 5] < 0) friend_num--;
        for (j = 0; j < friend_num; j++) {
            const int f_ID = m->face_interp_friends[i * 10 + j];
            v_tang += normal_vel[f_ID] * m->face_interp_weights[i * 10 + j] * m->face_len[f_ID];
        }
        v_tang /= m->face_node_dist[i];
        v_norm = normal_vel[i];
        const double nx = m->face_normal_vec_map[i * 2], ny = m->face_normal_vec_map[i * 2 + 1];
        const double ty = -nx, tx = ny;
        interp_vel[i * 2] = nx * v_norm + tx * v_tang;
        interp_vel[i * 2 + 1] = ny * v_norm + ty * v_tang;
    }
}

/* updateEnergy, energy.cpp:13-62 */
static void updateEnergy(const oracle_ctx* c, double* avg_flux_out, double* e_flux, const double* vel, const double* areas) {
    const double drag_coeff = c->p.alpha, h = c->p.h, r = c->p.radius;
    double avg_flux = 0.0;
    int i;
    if (c->p.friction == 0) {
        for (i = 0; i < c->F; i++) {
            e_flux[i] = drag_coeff * 1000.0 * h * (vel[i * 2] * vel[i * 2] + vel[i * 2 + 1] * vel[i * 2 + 1]);
            avg_flux += e_flux[i] * areas[i];
        }
    } else {
        for (i = 0; i < c->F; i++) {
            e_flux[i] = drag_coeff / h * sqrt(vel[i * 2] * vel[i * 2] + vel[i * 2 + 1] * vel[i * 2 + 1]) *
                        (vel[i * 2] * vel[i * 2] + vel[i * 2 + 1] * vel[i * 2 + 1]);
            avg_flux += e_flux[i] * areas[i];
        }
    }
    avg_flux /= (4 * pi * (r * r));
    *avg_flux_out = avg_flux;
}

/* pressureGradientSH, spatialOperators.cpp:387-462 (commented out at reference HEAD; restated from it):
 * coefficients of eta by least squares over the cell centres (sphericalHarmonics.cpp:16-72, extractSHCoeffGG.f95), then
 * forcing_potential += g * sh_matrix * coefficients (:446, dgemv alpha = factor = g, beta = 1) for degrees >= 2, the
 * basis rows carrying factor_l (mesh.cpp:2228-2238). */
static void self_gravity(oracle_ctx* c, double* potential, const double* eta) {
    const int R = c->sh_rows, N = c->N;
    int i, j, k;
    if (!R) return;
    for (k = 0; k < R; k++) {
        long double acc = 0.0L;
        for (i = 0; i < N; i++) acc += (long double)c->sh_Y[(size_t)k * N + i] * eta[i];
        c->sh_b[k] = (double)acc;
    }
    for (j = 0; j < R; j++) {
        long double acc = 0.0L;
        for (k = 0; k < R; k++) acc += (long double)c->sh_T[(size_t)j * R + k] * c->sh_b[k];
        c->sh_s[j] = c->p.g * (double)acc;
    }
    for (i = 0; i < N; i++) {
        double acc = 0.0;
        for (k = 4; k < R; k++) acc = acc + c->sh_Y[(size_t)k * N + i] * c->sh_s[k];
        potential[i] = potential[i] + acc;
    }
}

/* calculateMomentumAdvection, momAdvection.cpp:11-268 (live lines only) */
static void calculateMomentumAdvection(oracle_ctx* c, double* dvdt, const double* vel, const double* thickness_n, double* Ekin) {
    const oracle_mesh* m = &c->m;
    const int N = c->N, F = c->F, V = c->V;
    const double rot_rate = c->p.omega;
    int i, j;
    spmv_assign(&c->curl, vel, c->vorticity_v);                                   /* :31 */
    for (i = 0; i < V; i++) {                                                     /* :39-67 */
        double f = -2 * rot_rate * c->vertex_sinlat[i];
        double thickness = 0.0;
        for (j = 0; j < 3; j++) {
            const int node_ID = c->vertex_nodes[i * 3 + j];
            thickness += thickness_n[node_ID] * m->control_volume_surf_area_map[node_ID] * c->vertex_R[i * 3 + j];
        }
        thickness *= 1.0 / c->vertex_area[i];                                     /* vertex_area_r, mesh.cpp:1147 */
        c->vorticity_v[i] = (c->vorticity_v[i] + f) / thickness;
    }
    for (i = 0; i < F; i++) {                                                     /* :77-91 */
        const int v1 = c->face_vertexes[i * 2], v2 = c->face_vertexes[i * 2 + 1];
        const int n1 = m->face_nodes[i * 2], n2 = m->face_nodes[i * 2 + 1];
        c->vorticity_e[i] = 0.5 * (c->vorticity_v[v1] + c->vorticity_v[v2]);
        c->thickness_e[i] = 0.5 * (thickness_n[n1] + thickness_n[n2]);
    }
    for (i = 0; i < F; ++i) {                                                     /* :100-142 */
        int friend_num = 10;
        const int n1 = m->face_nodes[i * 2], n2 = m->face_nodes[i * 2 + 1];
        double q_e, F_tang_q = 0.0;
        if (m->node_friends[n1 * 6 + 5] < 0) friend_num--;
        if (m->node_friends[n2 * 6 + 5] < 0) friend_num--;
        q_e = c->vorticity_e[i];
        for (j = 0; j < friend_num; j++) {
            const int f_ID = m->face_interp_friends[i * 10 + j];
            const double q_e2 = c->vorticity_e[f_ID];
            const double F_e = c->thickness_e[f_ID] * vel[f_ID];
            const double coeff = m->face_interp_weights[i * 10 + j] * m->face_len[f_ID] * (1.0 / m->face_node_dist[i]);   /* face_node_dist_r, mesh.cpp:624 */
            F_tang_q += coeff * F_e * (q_e + q_e2) * 0.5;
        }
        dvdt[i] -= -F_tang_q;
    }
    spmv_assign(&c->rbf, vel, c->vel_xyz);                                        /* interpolateVelocityCartRBF, interpolation.cpp:116 */
    for (i = 0; i < N; i++)
        Ekin[i] = 0.5 * (c->vel_xyz[i * 3] * c->vel_xyz[i * 3] + c->vel_xyz[i * 3 + 1] * c->vel_xyz[i * 3 + 1] + c->vel_xyz[i * 3 + 2] * c->vel_xyz[i * 3 + 2]);
    {                                                                             /* :267: dvdt += -G * Ekin, -G materialised first */
        int k;
        for (i = 0; i < F; i++) {
            double tmp = 0;
            for (k = c->grad.ptr[i]; k < c->grad.ptr[i + 1]; k++) tmp += (-c->grad.val[k]) * Ekin[c->grad.idx[k]];
            dvdt[i] += 1.0 * tmp;
        }
    }
}

/* interpolateLSQFlux, interpolation.cpp:311-364 */
static void interpolateLSQFlux(oracle_ctx* c, double* flux, const double* edge_vel, const double* node_scalar) {
    const oracle_mesh* m = &c->m;
    const double fact = 1. / 12.0, beta = 1.0;
    int i;
    spmv_assign(&c->d2, node_scalar, c->d2v);
    for (i = 0; i < c->F; ++i) {
        const double dx = m->face_node_dist[i];
        const double dx2 = dx * dx * fact;
        const double vel = edge_vel[i];
        const int inner_ID = m->face_nodes[i * 2], outer_ID = m->face_nodes[i * 2 + 1];
        const double d2_outer = c->d2v[2 * i + 1], d2_inner = c->d2v[2 * i];
        flux[i] = vel * 0.5 * (node_scalar[inner_ID] + node_scalar[outer_ID]) - dx2 * (d2_outer + d2_inner) * vel +
                  dx2 * beta * fabs(vel) * (d2_outer - d2_inner);
    }
}

/* ---------------------------------------------------------------- public ---- */
oracle_ctx* oracle_create(const oracle_mesh* mesh, const oracle_params* params) {
    oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(oracle_ctx));
    c->N = mesh->n_cells; c->F = mesh->n_edges; c->m = *mesh; c->p = *params;
    build_operators(c);
    build_trig(c);
    const size_t N = (size_t)c->N, F = (size_t)c->F;
    c->v_t0 = (double*)calloc(F, 8); c->p_t0 = (double*)calloc(N, 8);
    c->dv_dt = (double*)calloc(F * 3, 8); c->dp_dt = (double*)calloc(N * 3, 8);
    c->dv_dt_t0 = (double*)calloc(F, 8); c->dp_dt_t0 = (double*)calloc(N, 8);
    c->drag_term = (double*)calloc(F, 8); c->forcing_potential = (double*)calloc(N, 8);
    c->v_avg = (double*)calloc(F * 2, 8); c->energy_diss = (double*)calloc(F, 8);
    c->tmp_f = (double*)calloc(F, 8); c->tmp_n = (double*)calloc(N, 8);
    c->iter = 0;
    return c;
}

void oracle_destroy(oracle_ctx* c) {
    if (!c) return;
    csr_free(&c->grad); csr_free(&c->div); csr_free(&c->cor); csr_free(&c->drag);
    free(c->trigLat); free(c->trigLon); free(c->trig2Lat); free(c->trig2Lon); free(c->trigSqLat);
    free(c->v_t0); free(c->p_t0); free(c->dv_dt); free(c->dp_dt); free(c->dv_dt_t0); free(c->dp_dt_t0);
    free(c->drag_term); free(c->forcing_potential); free(c->v_avg); free(c->energy_diss); free(c->tmp_f); free(c->tmp_n);
    free(c->sh_b); free(c->sh_s);
    free(c);
}

/* Nonlinear branch on: the three operators as CSR (row pointers, ascending columns, values: exactly what the reference built,
 * e.g. from oracle/_ref's table dump) and the vertex tables; all arrays stay owned by the caller. Call before oracle_set_state. */
static void csr_borrow_copy(csr* A, int nr, int nc, const int* ptr, const int* idx, const double* val) {
    const size_t nnz = (size_t)ptr[nr];
    A->nr = nr; A->nc = nc;
    A->ptr = (int*)malloc(((size_t)nr + 1) * sizeof(int)); memcpy(A->ptr, ptr, ((size_t)nr + 1) * sizeof(int));
    A->idx = (int*)malloc((nnz + 1) * sizeof(int)); memcpy(A->idx, idx, nnz * sizeof(int));
    A->val = (double*)malloc((nnz + 1) * sizeof(double)); memcpy(A->val, val, nnz * sizeof(double));
}
void oracle_set_nonlinear(oracle_ctx* c, int V, const int* curl_ptr, const int* curl_idx, const double* curl_val, const int* rbf_ptr,
                          const int* rbf_idx, const double* rbf_val, const int* d2_ptr, const int* d2_idx, const double* d2_val,
                          const double* vertex_sinlat, const double* vertex_area, const double* vertex_R, const int* vertex_nodes,
                          const int* face_vertexes) {
    const size_t N = (size_t)c->N, F = (size_t)c->F;
    size_t i;
    c->V = V;
    csr_borrow_copy(&c->curl, V, c->F, curl_ptr, curl_idx, curl_val);
    csr_borrow_copy(&c->rbf, 3 * c->N, c->F, rbf_ptr, rbf_idx, rbf_val);
    csr_borrow_copy(&c->d2, 2 * c->F, c->N, d2_ptr, d2_idx, d2_val);
    c->vertex_sinlat = vertex_sinlat; c->vertex_area = vertex_area; c->vertex_R = vertex_R;
    c->vertex_nodes = vertex_nodes; c->face_vertexes = face_vertexes;
    c->h_total = (double*)calloc(N, 8); c->ekin = (double*)calloc(N, 8); c->vorticity_v = (double*)calloc((size_t)V, 8);
    c->vorticity_e = (double*)calloc(F, 8); c->thickness_e = (double*)calloc(F, 8); c->vel_xyz = (double*)calloc(3 * N, 8);
    c->d2v = (double*)calloc(2 * F, 8); c->flux = (double*)calloc(F, 8);
    for (i = 0; i < N; i++) c->h_total[i] = c->p.h + c->p_t0[i];
    c->nl_on = 1;
}

/* Y, T stay owned by the caller */
void oracle_set_sh(oracle_ctx* c, int rows, const double* Y, const double* T) {
    free(c->sh_b); free(c->sh_s);
    c->sh_rows = rows; c->sh_Y = Y; c->sh_T = T;
    c->sh_b = (double*)calloc((size_t)rows + 1, 8); c->sh_s = (double*)calloc((size_t)rows + 1, 8);
}
void oracle_get_sh_b(const oracle_ctx* c, double* out) { memcpy(out, c->sh_b, (size_t)c->sh_rows * 8); }

/* state as ab3Explicit holds it after getInitialConditions (timeIntegrator.cpp:168-178) */
void oracle_set_state(oracle_ctx* c, const double* v, const double* eta, const double* dvdt, const double* detadt, long iter) {
    const size_t N = (size_t)c->N, F = (size_t)c->F;
    if (v) memcpy(c->v_t0, v, F * 8); else memset(c->v_t0, 0, F * 8);
    if (eta) memcpy(c->p_t0, eta, N * 8); else memset(c->p_t0, 0, N * 8);
    if (dvdt) memcpy(c->dv_dt, dvdt, F * 24); else memset(c->dv_dt, 0, F * 24);
    if (detadt) memcpy(c->dp_dt, detadt, N * 24); else memset(c->dp_dt, 0, N * 24);
    c->iter = iter;
    if (c->nl_on) { size_t i; for (i = 0; i < N; i++) c->h_total[i] = c->p.h + c->p_t0[i]; }   /* timeIntegrator.cpp:203 */
    interpolateVelocity(c, c->v_avg, c->v_t0);
    updateEnergy(c, &c->e_diss, c->energy_diss, c->v_avg, c->m.face_area);
}

/* the while-loop body, timeIntegrator.cpp:205-277 (linear path: advection false). diss_series, if
 * not NULL, receives e_diss after each step. */
void oracle_step(oracle_ctx* c, int nsteps, double* diss_series) {
    const int N = c->N, F = c->F;
    const double g = c->p.g, h = c->p.h, dt = c->p.dt;
    int k, i;
    for (k = 0; k < nsteps; k++) {
        const double current_time = dt * c->iter;                       /* :187,277 */
        /* updateMomentum, updateMomentum.cpp:42: dvdt = -g*(1-GAMMA*IMPLICIT)*G*eta + C*v */
        spmv_scaled_assign(&c->grad, -g * (1 - 0.5 * 0), c->p_t0, c->dv_dt_t0);
        if (c->nl_on) calculateMomentumAdvection(c, c->dv_dt_t0, c->v_t0, c->h_total, c->ekin);   /* updateMomentum.cpp:37-38 */
        else spmv_add(&c->cor, c->v_t0, c->dv_dt_t0);
        for (i = 0; i < F; ++i) c->dv_dt[i * 3] = c->dv_dt_t0[i];       /* :215 */
        forcing(c, c->forcing_potential, current_time + dt);             /* :218 */
        self_gravity(c, c->forcing_potential, c->p_t0);
        spmv_assign(&c->drag, c->v_t0, c->drag_term);                    /* :219 */
        spmv_add(&c->grad, c->forcing_potential, c->drag_term);
        integrateAB3scalar(c, c->v_t0, c->dv_dt, c->iter, F);            /* :239 */
        for (i = 0; i < F; ++i) c->v_t0[i] += dt * c->drag_term[i];      /* :242 */
        if (c->nl_on) {                                                  /* updateEta.cpp:32-33 */
            interpolateLSQFlux(c, c->flux, c->v_t0, c->h_total);
            spmv_assign(&c->div, c->flux, c->dp_dt_t0);
        } else
            spmv_scaled_assign(&c->div, h, c->v_t0, c->dp_dt_t0);        /* updateEta.cpp:39 */
        for (i = 0; i < N; ++i) c->dp_dt[i * 3] = c->dp_dt_t0[i];       /* :251 */
        integrateAB3scalar(c, c->p_t0, c->dp_dt, c->iter, N);            /* :253 */
        interpolateVelocity(c, c->v_avg, c->v_t0);                       /* :261 */
        updateEnergy(c, &c->e_diss, c->energy_diss, c->v_avg, c->m.face_area);   /* :262 */
        if (c->nl_on) for (i = 0; i < N; i++) c->h_total[i] = h + c->p_t0[i];    /* :266-269 */
        c->iter++;                                                       /* :276 */
        if (diss_series) diss_series[k] = c->e_diss;
    }
}

/* field ids match enum odis_field of include/odis_b200.h */
void oracle_get_field(const oracle_ctx* c, int field, double* out) {
    const size_t N = (size_t)c->N, F = (size_t)c->F;
    switch (field) {
        case 0: memcpy(out, c->v_t0, F * 8); break;
        case 1: memcpy(out, c->p_t0, N * 8); break;
        case 2: memcpy(out, c->dv_dt, F * 24); break;
        case 3: memcpy(out, c->dp_dt, N * 24); break;
        case 4: memcpy(out, c->v_avg, F * 16); break;
        case 5: memcpy(out, c->energy_diss, F * 8); break;
        case 6: memcpy(out, c->forcing_potential, N * 8); break;
        default: break;
    }
}
double oracle_get_dissipation_avg(const oracle_ctx* c) { return c->e_diss; }
long oracle_get_iter(const oracle_ctx* c) { return c->iter; }

/* CSR operator access for operator-level tests: which = 0 grad, 1 div, 2 coriolis, 3 drag */
void oracle_get_operator(const oracle_ctx* c, int which, int* nr, int* nc, const int** ptr, const int** idx, const double** val) {
    const csr* A = which == 0 ? &c->grad : which == 1 ? &c->div : which == 2 ? &c->cor : &c->drag;
    *nr = A->nr; *nc = A->nc; *ptr = A->ptr; *idx = A->idx; *val = A->val;
}

/* The loop-level functions one at a time, on caller arrays (parity checks of the odis_op_* entry points). Same routines
 * oracle_step runs, so they are pinned through it (tests/test_oracle_pinned.py::test_operator_calls_compose_to_a_step). */
void oracle_op_update_momentum(oracle_ctx* c, const double* v, const double* eta, double* dvdt) {   /* updateMomentum.cpp:42 */
    spmv_scaled_assign(&c->grad, -c->p.g * (1 - 0.5 * 0), eta, dvdt);
    spmv_add(&c->cor, v, dvdt);
}
void oracle_op_update_eta(oracle_ctx* c, const double* v, double* detadt) { spmv_scaled_assign(&c->div, c->p.h, v, detadt); }   /* updateEta.cpp:39 */
void oracle_op_forcing(oracle_ctx* c, double time, double* potential) { forcing(c, potential, time); }
void oracle_op_drag_forcing(oracle_ctx* c, const double* v, const double* potential, double* drag_term) {   /* timeIntegrator.cpp:219 */
    spmv_assign(&c->drag, v, drag_term);
    spmv_add(&c->grad, potential, drag_term);
}
void oracle_op_integrate_ab3_scalar(const oracle_ctx* c, double* s, double* ds_dt, long iter, int num) { integrateAB3scalar(c, s, ds_dt, iter, num); }
void oracle_op_interpolate_velocity(const oracle_ctx* c, const double* v, double* v_avg) { interpolateVelocity(c, v_avg, v); }
double oracle_op_update_energy(const oracle_ctx* c, const double* v_avg, const double* areas, double* e_flux) {
    double avg = 0.0;
    updateEnergy(c, &avg, e_flux, v_avg, areas);
    return avg;
}
