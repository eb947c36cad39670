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

namespace boost { namespace math {

template <class RealType = double>
struct hypergeometric_distribution {
    unsigned r_, n_, N_;
    hypergeometric_distribution(unsigned r, unsigned n, unsigned N) : r_(r), n_(n), N_(N) {}
};

namespace standin_detail {
inline double log_choose(double n, double k) {
    return std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0);
}
}  // namespace standin_detail

template <class RealType, class K>
inline RealType pdf(const hypergeometric_distribution<RealType>& d, K k) {
    const double r = d.r_, n = d.n_, N = d.N_, x = (double)k;
    return (RealType)std::exp(standin_detail::log_choose(r, x) +
                              standin_detail::log_choose(N - r, n - x) -
                              standin_detail::log_choose(N, n));
}

}}  // namespace boost::math
