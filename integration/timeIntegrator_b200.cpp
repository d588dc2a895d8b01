// Drop-in for GeodesicODIS's src/timeIntegrator.cpp: the same `int ab3Explicit(Globals*, Mesh*)` that solveODIS calls
// (src/solver.cpp:34), with the while-loop body (src/timeIntegrator.cpp:205-313) running on a B200 through libodis_b200.so.
// Everything around the loop stays the reference's own code: getInitialConditions (src/initialConditions.cpp), the
// DumpData / OUTPUT.txt cadence (src/outFiles.cpp:522-684), writeInitialConditions for the restart files, SIGINT handling.
//
// Build: add this file and odis_b200_bridge.cpp to the reference's sources in place of src/timeIntegrator.cpp, add
// -I<repo>/include, link -lodis_b200 (oracle/ref_build/Makefile target `hybrid` does exactly that for the tests).
#include "odis_b200_bridge.h"

#include "array1d.h"
#include "array2d.h"
#include "gridConstants.h"
#include "initialConditions.h"
#include "interpolation.h"
#include "mathRoutines.h"
#include "outFiles.h"
#include "timeIntegrator.h"

#include <signal.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

static volatile sig_atomic_t odis_b200_interrupted = 0;

static void odis_b200_on_sigint(int) {
    std::printf("%s\n", "Caught Terminate Signal...");
    odis_b200_interrupted = 1;
}

int ab3Explicit(Globals* globals, Mesh* grid) {
    using odis_bridge::check;
    OutFiles* Output = globals->Output;
    std::ostringstream outstring;

    // Host mirrors of the state and of everything DumpData may be asked for. DumpData receives raw pointers ONCE, before
    // the loop (src/timeIntegrator.cpp:143-152), so these arrays keep their addresses and are refreshed from the device
    // before every dump.
    Array1D<double> v_t0(FACE_NUM), p_t0(NODE_NUM), energy_diss(FACE_NUM), cv_mass(NODE_NUM);
    Array2D<double> dv_dt(FACE_NUM, 3), dp_dt(NODE_NUM, 3), v_avg(FACE_NUM, 2), v_xyz(NODE_NUM, 3);
    for (int i = 0; i < NODE_NUM; i++) cv_mass(i) = 0.0;
    double total_diss = 0.0, current_time = 0.0;

    const std::vector<std::string>& tags = globals->out_tags;
    std::vector<double*> pp(tags.size(), nullptr);
    bool want_cartesian = false, want_diss_field = false, want_velocity = false;
    for (size_t i = 0; i < tags.size(); i++) {
        if (tags[i] == "velocity output") { pp[i] = &v_avg(0, 0); want_velocity = true; }
        else if (tags[i] == "velocity cartesian output") { pp[i] = &v_xyz(0, 0); want_cartesian = true; }
        else if (tags[i] == "displacement output") pp[i] = &p_t0(0);
        else if (tags[i] == "dissipation output") { pp[i] = &energy_diss(0); want_diss_field = true; }
        else if (tags[i] == "dissipation avg output") pp[i] = &total_diss;
        else if (tags[i] == "kinetic avg output") pp[i] = &current_time;
        else if (tags[i] == "dummy1 output") pp[i] = &cv_mass(0);
    }

    signal(SIGINT, odis_b200_on_sigint);

    const double dt = globals->timeStep.Value();
    const double orbit_period = globals->period.Value();
    const double r = globals->radius.Value();

    getInitialConditions(globals, grid, v_t0, dv_dt, p_t0, dp_dt);       // zeros, restart files or the analytical solution

    // one solver, or ODIS_B200_GPUS=N partitioned ones driven from this process (odis_b200_bridge.h)
    const odis_bridge::Group& dev = odis_bridge::group(globals, grid);
    odis_bridge::set_state_all(globals, dev, &v_t0(0), &p_t0(0), &dv_dt(0, 0), &dp_dt(0, 0), /*iter*/ 0);
    if (dev.world > 1) {
        outstring << "grid partitioned over " << dev.world << " GPUs" << std::endl;
        Output->Write(OUT_MESSAGE, &outstring);
    }

    // device -> the arrays DumpData reads (what interpolateVelocity / updateEnergy leave behind every step in the
    // reference, src/timeIntegrator.cpp:261-263; here only when somebody looks)
    auto refresh_outputs = [&]() {
        odis_bridge::get_field_all(globals, dev, ODIS_FIELD_ETA, &p_t0(0), NODE_NUM);
        if (want_velocity) odis_bridge::get_field_all(globals, dev, ODIS_FIELD_VELOCITY_EN, &v_avg(0, 0), (size_t)FACE_NUM * 2);
        if (want_diss_field) odis_bridge::get_field_all(globals, dev, ODIS_FIELD_DISSIPATION, &energy_diss(0), FACE_NUM);
        if (want_cartesian) {       // output-cadence only: the reference's own RBF reconstruction on the host copy of v
            odis_bridge::get_field_all(globals, dev, ODIS_FIELD_VELOCITY, &v_t0(0), FACE_NUM);
            interpolateVelocityCartRBF(globals, grid, v_xyz, v_t0);
        }
        total_diss = odis_bridge::dissipation_avg_all(globals, dev);
    };
    int out_count = 1;
    auto log_and_dump = [&]() {
        outstring << std::fixed << "DUMPING DATA AT " << current_time / orbit_period;
        outstring << " AVG DISS: " << std::scientific << total_diss * 4 * pi * r * r / 1e9 << " GW" << out_count;
        Output->Write(OUT_MESSAGE, &outstring);
        Output->DumpData(globals, out_count, pp.data());
        out_count++;
    };

    long iter = 0;
    refresh_outputs();
    log_and_dump();                                                       // slice 1: the initial state

    const int out_freq = globals->totalIter.Value() / globals->outputTime.Value();
    const double bound = globals->totalIter.Value() * globals->endTime.Value();      // `iter < bound` as the reference writes it
    const long last = (long)std::ceil(bound);
    const int kChunk = 512;                                               // SIGINT is honoured between chunks
    while ((double)iter < bound) {
        long n = std::min<long>(out_freq - iter % out_freq, last - iter);
        n = std::min<long>(std::max<long>(n, 1), kChunk);
        odis_bridge::step_all(globals, dev, (int32_t)n);
        iter += n;
        current_time = dt * iter;
        if (iter % out_freq == 0) {
            refresh_outputs();
            log_and_dump();
        }
        if (odis_b200_interrupted) {
            outstring << "Terminate signal caught..." << std::endl;
            Output->Write(OUT_MESSAGE, &outstring);
            break;
        }
    }

    // full state back for the restart files (src/timeIntegrator.cpp:316)
    odis_bridge::get_field_all(globals, dev, ODIS_FIELD_VELOCITY, &v_t0(0), FACE_NUM);
    odis_bridge::get_field_all(globals, dev, ODIS_FIELD_ETA, &p_t0(0), NODE_NUM);
    odis_bridge::get_field_all(globals, dev, ODIS_FIELD_DVDT, &dv_dt(0, 0), (size_t)FACE_NUM * 3);
    odis_bridge::get_field_all(globals, dev, ODIS_FIELD_DETADT, &dp_dt(0, 0), (size_t)NODE_NUM * 3);
    writeInitialConditions(globals, grid, v_t0, dv_dt, p_t0, dp_dt);
    Output->Write(OUT_MESSAGE, &outstring);
    odis_bridge::release();
    return odis_b200_interrupted ? 1 : 0;
}
