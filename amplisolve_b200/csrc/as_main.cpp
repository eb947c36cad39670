// The drop-in executables: same names and argv grammar as the reference's programs
// (README.md:60-61; AmpliSolveErrorEstimation.cpp:241, AmpliSolveVariantCalling.cpp:199) and as its pre-processing
// binary computeCounts / ASEQ PILEUP mode (Execution_examples.md:27).
// Everything happens in libamplisolve_b200.so behind the C ABI.
#include <unistd.h>

#include <cstdio>
#include <iostream>

#include "amplisolve_b200.h"

int main(int argc, char** argv) {
#if defined(AS_MAIN_SERVE)
    const int rc = as_serve_main(argc, argv);
#else
#if defined(AS_MAIN_EE)
    const int prog = 0;
#elif defined(AS_MAIN_VC)
    const int prog = 1;
#elif defined(AS_MAIN_CC)
    const int prog = 2;
#else
#error "define AS_MAIN_EE, AS_MAIN_VC, AS_MAIN_CC or AS_MAIN_SERVE"
#endif
    int rc = 0;
    // AS_SERVER=<socket> and a service listening there: the program runs in that process, on its warm CUDA context
    if (!as_client_run(prog, argc, argv, &rc))
        rc = prog == 0 ? as_error_estimation_main(argc, argv) : prog == 1 ? as_variant_calling_main(argc, argv) : as_compute_counts_main(argc, argv);
#endif
    // every output file is closed and the context destroyed by now; leave without the CUDA runtime's exit handlers
    // (a few tenths of a second of a program that otherwise runs for two)
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    _exit(rc);
}
