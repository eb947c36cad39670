// TEST INFRASTRUCTURE ONLY -- stand-in for Boost.Math 1.61's hypergeometric.hpp.
//
// The reference's AmpliSolveVariantCalling.cpp:135 includes
// <boost/math/distributions/hypergeometric.hpp>, but Boost is a missing blob of
// /root/reference (.MISSING_LARGE_BLOBS:2) and is not installed in this image.  The header
// is used at one call site only, fisherTest (AmpliSolveVariantCalling.cpp:3797-3814):
//     hypergeometric_distribution<> hgd(r, n, N);   pdf(hgd, k)
// This stand-in supplies exactly that surface so the UNMODIFIED reference source can be
// compiled into oracle/_ref/.  pdf = C(r,k) C(N-r,n-k) / C(N,n) through lgamma in double.
// Everything on the Poisson path (kfunc, Q score, call decision) is in the reference's own
// file and untouched by this header.  Fisher p-values are therefore "parity unpinned"
// against real Boost beyond printed precision.
#pragma once
#include <cmath>

#include "../../../../../factorials.h"

namespace boost { namespace math {

template <class RealType = double>
struct hypergeometric_distribution {
    unsigned r_, n_, N_;
    hypergeometric_distribution(unsigned r, unsigned n, unsigned N) : r_(r), n_(n), N_(N) {}
};

namespace standin_detail {
// operands in a canonical order (smaller of k, n - k first): C(n,k) == C(n,n-k) bitwise, so that terms of a symmetric
// table that are tied mathematically are tied here as well (Boost's own pdf treats them alike: tests/golden/fisher_boost.npz)
inline double log_choose(double n, double k) {
    const double lo = k < n - k ? k : n - k, hi = k < n - k ? n - k : k;
    return (std::lgamma(n + 1.0) - std::lgamma(lo + 1.0)) - std::lgamma(hi + 1.0);
}
}  // namespace standin_detail

// N <= 170: Boost's own factorial-table method (hypergeometric_pdf_factorial_imp), restated; bit-identical to SciPy's Boost
inline double pdf_factorial(unsigned r, unsigned n, unsigned N, unsigned k) {
    double result = ASO_FACTORIAL[n];
    const double num[3] = {ASO_FACTORIAL[r], ASO_FACTORIAL[N - n], ASO_FACTORIAL[N - r]};
    const double den[5] = {ASO_FACTORIAL[N], ASO_FACTORIAL[k], ASO_FACTORIAL[n - k], ASO_FACTORIAL[r - k], ASO_FACTORIAL[N - n - r + k]};
    int i = 0, j = 0;
    while (i < 3 || j < 5) {
        while (j < 5 && (result >= 1 || i >= 3)) result /= den[j++];
        while (i < 3 && (result <= 1 || j >= 5)) result *= num[i++];
    }
    return result > 1 ? 1.0 : result;
}

template <class RealType, class K>
inline RealType pdf(const hypergeometric_distribution<RealType>& d, K k) {
    if (d.N_ <= 170) return (RealType)pdf_factorial(d.r_, d.n_, d.N_, (unsigned)k);
    const double r = d.r_, n = d.n_, N = d.N_, x = (double)k;
    return (RealType)std::exp(standin_detail::log_choose(r, x) +
                              standin_detail::log_choose(N - r, n - x) -
                              standin_detail::log_choose(N, n));
}

}}  // namespace boost::math
