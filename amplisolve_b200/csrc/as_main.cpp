// The drop-in executables: same names and argv grammar as the reference's programs
// (README.md:60-61; AmpliSolveErrorEstimation.cpp:241, AmpliSolveVariantCalling.cpp:199) and as its pre-processing
// binary computeCounts / ASEQ PILEUP mode (Execution_examples.md:27).
// Everything happens in libamplisolve_b200.so behind the C ABI.
#include <unistd.h>

#include <cstdio>
#include <iostream>

#include "amplisolve_b200.h"

int main(int argc, char** argv) {
#if defined(AS_MAIN_EE)
    const int rc = as_error_estimation_main(argc, argv);
#elif defined(AS_MAIN_VC)
    const int rc = as_variant_calling_main(argc, argv);
#elif defined(AS_MAIN_CC)
    const int rc = as_compute_counts_main(argc, argv);
#else
#error "define AS_MAIN_EE, AS_MAIN_VC or AS_MAIN_CC"
#endif
    // every output file is closed and the context destroyed by now; leave without the CUDA runtime's exit handlers
    // (a few tenths of a second of a program that otherwise runs for two)
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    _exit(rc);
}
