"""Critical Poisson means of the caller's second exact screen (AS_MCRIT in amplisolve_b200/csrc/as_kernels.cu):
m*(k) with P(X >= k | m*) = P*, P* = the largest double p whose long-double Q = -10 log10(p) reaches 5
(0x3FD43D136248490E, tests/test_oracle_golden.py), solved with mpmath to 40 digits and multiplied by 1 + 1e-9.

    python scripts/critical_means.py            # prints the table as C initialisers
"""
import struct

import mpmath as mp

mp.mp.dps = 50
P_STAR = mp.mpf(struct.unpack("<d", struct.pack("<Q", 0x3FD43D136248490E))[0])
MARGIN = mp.mpf("1e-9")


def critical_means(k_max=64):
    out = []
    for k in range(1, k_max + 1):
        f = lambda m: mp.gammainc(k, 0, m, regularized=True) - P_STAR  # noqa: E731  P(X >= k | m) = P(k, m)
        r = mp.findroot(f, (mp.mpf("1e-6"), mp.mpf(k)), solver="anderson", tol=1e-40)
        r = mp.findroot(f, r, tol=1e-45)
        assert abs(f(r)) < 1e-40
        out.append(float(r * (1 + MARGIN)))
    return out


if __name__ == "__main__":
    v = critical_means()
    for i in range(0, len(v), 4):
        print("    " + ", ".join(repr(x) for x in v[i:i + 4]) + ",")
