// See odis_analytic.h. The closed form is the two-mode truncation of the LTE response to the degree-2 order-1 westward
// obliquity potential: a stream-function mode Psi_11 and a velocity-potential mode Phi_21 coupled through the Coriolis
// recurrence coefficients p_n, q_n and the admittances K_n, L_n (analyticalLTE.cpp:29-45, :133-173). The arithmetic below is
// kept in the reference's operation order (std::complex, left to right) so that the loaded state agrees with the reference's
// to the last bit; the position-independent factors are formed once instead of once per sample (same operations, same bits).
#include "odis_analytic.h"

#include <cmath>
#include <complex>

#include "odis_sphere.h"

namespace odis {

namespace {

using cplx = std::complex<double>;
const cplx kI(0.0, 1.0);

double coupling_p(double n, double m) { return (n + 1) * (n + m) / (n * (2 * n + 1)); }          // analyticalLTE.cpp:29-31
double coupling_q(double n, double m) { return n * (n + 1 - m) / ((n + 1) * (2 * n + 1)); }      // :33-35
cplx admittance_K(double n, double m, double lambs, double lam, double alpha, double Omega) {    // :37-39
    return lam + m / (n * (n + 1)) - n * (n + 1) / (lambs * lam) + kI * alpha / (2 * Omega);
}
cplx admittance_L(double n, double m, double, double lam, double alpha, double Omega) {          // :41-43
    return lam + m / (n * (n + 1)) + kI * alpha / (2 * Omega);
}

struct Modes {            // what does not depend on the sample point
    double radius, h, w;
    cplx Psi11, Phi21;
};

Modes modes_of(const AnalyticParams& p) {
    Modes m;
    m.radius = p.radius; m.h = p.h;
    const double Omega = p.omega;
    m.w = -Omega;                                                   // westward: frequency -Omega (:71)
    const double lam = m.w / (2 * Omega);
    const double lambs = 4 * Omega * Omega * p.radius * p.radius / (p.g * p.h);      // Lamb parameter (:74)
    const cplx L1 = admittance_L(1, 1, lambs, lam, p.alpha, Omega);
    const cplx K2 = admittance_K(2, 1, lambs, lam, p.alpha, Omega);
    const cplx U21 = 0.5 * Omega * Omega * p.radius * p.radius * p.obl;              // :139
    m.Psi11 = U21 / (2 * Omega) / (coupling_q(1, 1) - K2 * L1 / coupling_p(2, 1));   // :140
    m.Phi21 = -kI * L1 / coupling_p(2, 1) * m.Psi11;                                   // :141
    return m;
}

// {u_east, v_north, du/dt, dv/dt, eta, deta/dt} at (lat, lon, t), analyticalLTE.cpp:144-173
void sample(const Modes& m, double lat, double lon, double t, double out[6]) {
    const double colat = kPi * 0.5 - lat;
    const double radius = m.radius, w = m.w;
    const cplx Y31 = 1.5 * (5 * std::pow(std::cos(colat), 2.0) - 1) * std::sin(colat) * std::exp(kI * 1.0 * lon);
    const cplx Y21 = 3 * std::cos(colat) * std::sin(colat) * std::exp(kI * 1.0 * lon);
    const cplx Y11 = std::sin(colat) * std::exp(kI * 1.0 * lon);
    const cplx phase = std::exp(-kI * w * t);

    cplx U = (-1 / 3. * (m.Psi11 * Y21) + kI * m.Phi21 * Y21) * phase;
    cplx dUdt = -kI * w * U;
    U += std::conj(U);
    U /= 2 * radius * std::sin(colat);
    dUdt += std::conj(dUdt);
    dUdt /= 2 * radius * std::sin(colat);

    cplx V = -(kI * m.Psi11 * Y11 + m.Phi21 * (4.0 * Y31 - 9.0 * Y11) / 5.) * phase;
    cplx dVdt = -kI * w * V;
    V += std::conj(V);
    V /= 2 * radius * std::sin(colat);
    dVdt += std::conj(dVdt);
    dVdt /= 2 * radius * std::sin(colat);

    cplx ETA = kI * 6.0 / std::pow(radius, 2.0) * m.h / w * m.Phi21 * Y21 * phase;
    cplx dETAdt = -kI * w * ETA;
    ETA += std::conj(ETA);
    ETA *= 0.5;
    dETAdt += std::conj(dETAdt);
    dETAdt *= 0.5;

    out[0] = U.real(); out[1] = V.real(); out[2] = dUdt.real(); out[3] = dVdt.real(); out[4] = ETA.real(); out[5] = dETAdt.real();
}

}  // namespace

void analytical_state_obliq_west(const AnalyticParams& p, int n_cells, int n_edges, const double* node_pos_sph, const double* face_centre_pos_sph,
                                 const double* face_normal_vec_map, double* v, double* dvdt, double* eta, double* detadt) {
    const Modes m = modes_of(p);
    const double t = 0.0 * 2 * kPi / p.omega;                       // initialConditions.cpp:152
    const double dt = p.dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_edges; i++) {                             // :157-181: edge midpoints, projected on the edge normal
        const double lat = face_centre_pos_sph[(size_t)i * 2], lon = face_centre_pos_sph[(size_t)i * 2 + 1];
        const double nx = face_normal_vec_map[(size_t)i * 2], ny = face_normal_vec_map[(size_t)i * 2 + 1];
        double a[6];
        sample(m, lat, lon, t, a);
        v[i] = a[0] * nx + a[1] * ny;
        dvdt[(size_t)i * 3] = a[2] * nx + a[3] * ny;
        sample(m, lat, lon, t - dt, a);
        dvdt[(size_t)i * 3 + 1] = a[2] * nx + a[3] * ny;
        sample(m, lat, lon, t - 2 * dt, a);
        dvdt[(size_t)i * 3 + 2] = a[2] * nx + a[3] * ny;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_cells; i++) {                             // :183-203: cell centres
        const double lat = node_pos_sph[(size_t)i * 2], lon = node_pos_sph[(size_t)i * 2 + 1];
        double a[6];
        sample(m, lat, lon, t, a);
        eta[i] = a[4];
        detadt[(size_t)i * 3] = a[5];
        sample(m, lat, lon, t - dt, a);
        detadt[(size_t)i * 3 + 1] = a[5];
        sample(m, lat, lon, t - 2 * dt, a);
        detadt[(size_t)i * 3 + 2] = a[5];
    }
}

}  // namespace odis
