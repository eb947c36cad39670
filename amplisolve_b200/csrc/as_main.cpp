// The two drop-in executables: same names and argv grammar as the reference's programs
// (README.md:60-61; AmpliSolveErrorEstimation.cpp:241, AmpliSolveVariantCalling.cpp:199).
// Everything happens in libamplisolve_b200.so behind the C ABI.
#include "amplisolve_b200.h"

int main(int argc, char** argv) {
#if defined(AS_MAIN_EE)
    return as_error_estimation_main(argc, argv);
#elif defined(AS_MAIN_VC)
    return as_variant_calling_main(argc, argv);
#else
#error "define AS_MAIN_EE or AS_MAIN_VC"
#endif
}
