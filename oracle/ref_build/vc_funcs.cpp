// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libvc_ref_funcs.so from the UNMODIFIED
// reference source: AmpliSolveVariantCalling.cpp is #included from /root/reference (path via
// -DREF_VC_CPP=...) with main() renamed, and its own functions are exposed through plain-C
// wrappers so tests can obtain function-level golden vectors:
//   kf_gammaq                          AmpliSolveVariantCalling.cpp:3726
//   mutationRulesPoissonQualityScore   AmpliSolveVariantCalling.cpp:3834
//   fisherTest                         AmpliSolveVariantCalling.cpp:3797 (Boost stand-in, see shim/)
//   homopolymerTest                    AmpliSolveVariantCalling.cpp:3615
#define main ampli_reference_main_vc
#include REF_VC_CPP
#undef main

extern "C" {

double ref_kf_gammaq(double s, double z) { return kf_gammaq(s, z); }

// Q as the reference's long double, plus its double rounding (what `double(Q_fw+Q_bw)` sees).
void ref_poisson_q(int k, int rd, float err, long double* q_ld, double* q_d) {
    long double q = mutationRulesPoissonQualityScore(k, rd, err);
    *q_ld = q;
    *q_d = (double)q;
}

// Vectorised variants for grids (avoid ctypes per-call overhead).
void ref_kf_gammaq_vec(const double* s, const double* z, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = kf_gammaq(s[i], z[i]);
}

void ref_poisson_q_vec(const int* k, const int* rd, const float* err, double* q_d, long n) {
    for (long i = 0; i < n; ++i) q_d[i] = (double)mutationRulesPoissonQualityScore(k[i], rd[i], err[i]);
}

// The decision exactly as AmpliSolveVariantCalling.cpp:898 evaluates it (long double compare).
void ref_call_decision_vec(const int* kfw, const int* fw, const float* efw, const int* kbw, const int* bw,
                           const float* ebw, int cut, unsigned char* call, long n) {
    for (long i = 0; i < n; ++i) {
        long double qf = mutationRulesPoissonQualityScore(kfw[i], fw[i], efw[i]);
        long double qb = mutationRulesPoissonQualityScore(kbw[i], bw[i], ebw[i]);
        call[i] = (fw[i] >= cut && bw[i] >= cut && qf >= 5 && qb >= 5) ? 1 : 0;
    }
}

double ref_fisher(int a, int b, int c, int d) { return fisherTest(a, b, c, d); }

int ref_homopolymer(const char* down, const char* up, char sub) {
    return homopolymerTest((char*)down, (char*)up, sub);
}

}  // extern "C"
