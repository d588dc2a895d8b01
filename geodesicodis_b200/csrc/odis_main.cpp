// `ODIS` — run from a directory that holds input.in and input_files/grid_l<L>.txt, like the reference
// executable (/root/reference/src/main.cpp:46-68, Makefile:7 EXE = ./ODIS). Everything is in the library.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/odis_b200.h"

int main(int argc, char** argv) {
    odis_run_options opt{};
    opt.reorder = 1;
    opt.echo = 1;
    const char* dir = ".";
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--device") && i + 1 < argc) opt.device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--max-steps") && i + 1 < argc) opt.max_steps = std::atoll(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) opt.echo = 0;
        else if (!std::strcmp(argv[i], "--self-gravity")) opt.self_gravity = 1;   // pressureGradientSH, off at reference HEAD
        else if (!std::strcmp(argv[i], "--overlap-output")) opt.overlap_output = 1;  // dumps written while the next interval is computed
        else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) opt.n_gpus = std::atoi(argv[++i]);   // grid partitioned over N GPUs
        else if (!std::strcmp(argv[i], "--dir") && i + 1 < argc) dir = argv[++i];
        else { std::fprintf(stderr, "usage: ODIS [--dir RUN_DIR] [--device N] [--gpus N] [--max-steps K] [--quiet] [--self-gravity] [--overlap-output]\n"); return 2; }
    }
    odis_run_result res{};
    const int rc = odis_run(dir, &opt, &res);
    if (rc != ODIS_OK) {
        std::printf("ODIS HAS FOUND AN ERROR. TERMINATING PROGRAM.\n%s\n", odis_last_error());   // outFiles.cpp:127
        return 0;   // the reference exits with status 0 from TerminateODIS (outFiles.cpp:129)
    }
    return res.interrupted ? 1 : 0;
}
