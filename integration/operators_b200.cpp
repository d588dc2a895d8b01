// Drop-ins for the free functions GeodesicODIS's time loop calls (src/timeIntegrator.cpp:205-313), with the reference's
// own C++ signatures, each forwarding to one odis_op_* call of libodis_b200.so. For a maintainer who keeps the reference's
// ab3Explicit as it is and moves single functions to the GPU:
//
//   updateMomentum       include/updateMomentum.h:8-16      src/updateMomentum.cpp:16-47
//   updateEta            include/updateEta.h:8-13           src/updateEta.cpp:7-44
//   forcing              include/tidalPotentials.h:16       src/tidalPotentials.cpp:29-328
//   integrateAB3scalar   include/temporalOperators.h:14     src/temporalOperators.cpp:17-68
//   interpolateVelocity  include/interpolation.h:14         src/interpolation.cpp:26-62
//   updateEnergy         include/energy.h:10                src/energy.cpp:13-62
//
// Return values follow the reference (1 = OK). Every call copies its arrays to the device and back, so this level is for
// validation and incremental adoption; the fast path is integration/timeIntegrator_b200.cpp. Linear branch only: with
// `advection; true` the calls fail loudly through Output->TerminateODIS (there is no CPU fallback).
//
// Build: compile the reference's TUs of these functions with the symbol renamed (e.g. -DinterpolateVelocity=
// interpolateVelocity_reference) or leave the single-function TUs out, add this file + odis_b200_bridge.cpp, link -lodis_b200
// (oracle/ref_build/Makefile target `hybridops`).
#include "odis_b200_bridge.h"

#include "array1d.h"
#include "array2d.h"
#include "energy.h"
#include "interpolation.h"
#include "temporalOperators.h"
#include "tidalPotentials.h"
#include "updateEta.h"
#include "updateMomentum.h"

using odis_bridge::check;

int updateMomentum(Globals* constants, Mesh* grid, Array1D<double>& dvdt, Array1D<double>& v_tm1, Array1D<double>& p_tm1,
                   Array1D<double>& /*h_total*/, Array1D<double>& /*ekin*/, double GAMMA, double IMPLICIT) {
    if (GAMMA * IMPLICIT != 0.0) check(constants, ODIS_ERR_UNSUPPORTED, "updateMomentum with the semi-implicit pressure split");
    odis_solver* s = odis_bridge::solver(constants, grid);
    check(constants, odis_op_update_momentum(s, &v_tm1(0), &p_tm1(0), &dvdt(0)), "odis_op_update_momentum");
    return 1;
}

int updateEta(Globals* globals, Mesh* grid, Array1D<double>& deta_dt, Array1D<double>& v_t0, Array1D<double>& /*eta*/,
              Array1D<double>& /*h_total*/) {
    odis_solver* s = odis_bridge::solver(globals, grid);
    check(globals, odis_op_update_eta(s, &v_t0(0), &deta_dt(0)), "odis_op_update_eta");
    return 1;
}

void forcing(Globals* consts, Mesh* grid, Array1D<double>& potential, int forcing_type, double time, double ecc, double obl) {
    // the device solver evaluates the potential it was created for (globals->tide_type, ->e, ->theta)
    if (forcing_type != (int)consts->tide_type || ecc != consts->e.Value() || obl != consts->theta.Value())
        check(consts, ODIS_ERR_UNSUPPORTED, "forcing with a potential other than the run's own");
    if (consts->tide_type == NONE) return;                      // the reference leaves the array untouched (:283)
    odis_solver* s = odis_bridge::solver(consts, grid);
    check(consts, odis_op_forcing(s, time, &potential(0)), "odis_op_forcing");
}

int integrateAB3scalar(Globals* globals, Mesh* grid, Array1D<double>& s, Array2D<double>& ds_dt, int iter, int num) {
    odis_solver* dev = odis_bridge::solver(globals, grid);
    check(globals, odis_op_integrate_ab3_scalar(dev, &s(0), &ds_dt(0, 0), iter, num), "odis_op_integrate_ab3_scalar");
    return 1;
}

int interpolateVelocity(Globals* globals, Mesh* mesh, Array2D<double>& interp_vel, Array1D<double>& normal_vel) {
    odis_solver* s = odis_bridge::solver(globals, mesh);
    check(globals, odis_op_interpolate_velocity(s, &normal_vel(0), &interp_vel(0, 0)), "odis_op_interpolate_velocity");
    return 1;
}

void updateEnergy(Globals* globals, double& avg_flux, Array1D<double>& e_flux, Array2D<double>& vel, Array1D<double>& areas) {
    odis_solver* s = odis_bridge::solver(globals, nullptr);     // no Mesh argument: the solver interpolateVelocity created
    check(globals, odis_op_update_energy(s, &vel(0, 0), &areas(0), &e_flux(0), &avg_flux), "odis_op_update_energy");
}
