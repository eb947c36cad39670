#!/usr/bin/env python
"""Pins the Fisher strand-bias test (VC:3797-3814) against a REAL Boost.Math hypergeometric pdf.

Boost.Math 1.61 is a missing blob of the reference tree, so the compiled reference under
oracle/_ref uses a stand-in header.  SciPy, however, ships Boost.Math compiled in: scipy.special's
`_hypergeom_pmf` ufunc is `boost::math::pdf(hypergeometric_distribution<double>(r, n, N), k)` -- the very
call of VC:3805 / VC:3810 (a newer Boost than 1.61; hypergeometric_pdf.hpp -- factorial table for N <= 170,
prime factorisation up to N = 104,723, Lanczos beyond -- has not changed its method since).  This script
evaluates the reference's own loop (cutoff = pdf(c); sequential sum of the pdf(k) <= cutoff, k ascending)
over that pdf and stores the p-values, for
  * every call row of the three golden cases (tables recovered from the Summary columns RD_fw, RD_bw,
    Reads_fw, Reads_bw: fisherTest(FW, BW, alt_fw, alt_bw), VC:902), and
  * seeded random tables at the depths of BASELINE.json's configs (200x ... 100,000x; balanced, strand-biased,
    germline-like, alt = 0 on one strand, symmetric tables where pdf ties with the cutoff).
Runs in the dev container only (needs scipy); the fixture tests/golden/fisher_boost.npz is committed.

    python tests/golden/make_fisher_boost.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))


def boost_fisher(a, b, c, d):
    """VC:3797-3814 over scipy's Boost pdf.  np.cumsum adds sequentially (k ascending) like the reference's loop."""
    import scipy.special._ufuncs as scu
    N, r, n = a + b + c + d, a + c, c + d
    hi, lo = min(r, n), max(0, r + n - N)
    k = np.arange(lo, hi + 1, dtype=np.float64)
    # scipy's argument order: _hypergeom_pmf(k, n, N, M) = boost pdf(hypergeometric_distribution(r=n, n=N, N=M), k)
    pdf = scu._hypergeom_pmf(k, float(r), float(n), float(N))
    cutoff = pdf[c - lo]
    kept = pdf[pdf <= cutoff]
    return float(np.cumsum(kept)[-1]) if kept.size else 0.0


def tables_from_summary(text):
    rows = []
    for line in str(text).splitlines()[1:]:
        f = line.split("\t")
        if len(f) < 14:
            continue
        fw, bw, afw, abw = int(f[5]), int(f[6]), int(f[8]), int(f[9])
        rows.append((fw, bw, afw, abw, f[13]))
    return rows


def random_tables(rng, n):
    out = []
    for _ in range(n):
        depth = int(rng.choice([200, 500, 1000, 2000, 5000, 10000, 50000, 100000]) * rng.lognormal(0, 0.4))
        depth = max(40, depth)
        fw = int(rng.binomial(depth, rng.choice([0.5, 0.5, 0.3, 0.1])))
        bw = depth - fw
        kind = rng.integers(0, 6)
        if kind == 0:      # low-VAF somatic, balanced
            vaf = rng.uniform(0.003, 0.05)
            afw, abw = rng.binomial(fw, vaf), rng.binomial(bw, vaf)
        elif kind == 1:    # strand-biased
            vaf = rng.uniform(0.005, 0.2)
            afw, abw = rng.binomial(fw, vaf), rng.binomial(bw, vaf * rng.choice([0.0, 0.1, 0.3]))
        elif kind == 2:    # germline-like
            vaf = rng.choice([0.5, 1.0]) * rng.uniform(0.9, 1.0)
            afw, abw = rng.binomial(fw, vaf), rng.binomial(bw, vaf)
        elif kind == 3:    # tiny counts
            afw, abw = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        elif kind == 4:    # symmetric table: pdf(k) ties with pdf(c) at the mirror k up to rounding
            fw = bw = depth // 2
            afw = int(rng.integers(0, min(fw, 60)))
            abw = int(rng.integers(0, min(bw, 60)))
        else:              # small N (Boost's factorial-table branch, N <= 170)
            fw, bw = int(rng.integers(10, 80)), int(rng.integers(10, 80))
            afw, abw = int(rng.integers(0, min(fw, 10) + 1)), int(rng.integers(0, min(bw, 10) + 1))
        afw, abw = int(min(afw, fw)), int(min(abw, bw))
        out.append((fw, bw, afw, abw))
    return out


def tie_tables():
    """Tables whose hypergeometric distribution is symmetric, so that pdf(k) ties exactly with pdf(mirror k): r = N - r
    (a + c = b + d) or n = N - n (c + d = a + b).  Exhaustive for N <= 28 (Boost's factorial-table branch), plus larger ones
    on the prime-factorisation and Lanczos branches.  Whether the mirror term is <= cutoff decides a whole term of p."""
    out = []
    for N in range(4, 29, 2):
        for c in range(0, N // 2 + 1):
            for d in range(0, N // 2 + 1 - c):
                for a in range(0, N - c - d + 1):
                    b = N - a - c - d
                    if (a + c == b + d or c + d == a + b) and a >= 0 and b >= 0:
                        out.append((a, b, c, d))
    rng = np.random.default_rng(20187)
    for _ in range(400):
        half = int(rng.choice([60, 90, 200, 1000, 5000, 40000, 60000, 150000]))
        c, d = int(rng.integers(0, min(half, 40))), int(rng.integers(0, min(half, 40)))
        if rng.random() < 0.5:      # r = N - r
            a = half - c
            b = half - d
        else:                        # n = N - n  (alt reads = half of everything: only for small halves)
            half = min(half, 200)
            c, d = int(rng.integers(0, half + 1)), 0
            d = half - c
            a = int(rng.integers(0, half + 1))
            b = half - a
        if a >= 0 and b >= 0:
            out.append((a, b, c, d))
    return out


def main():
    import scipy
    rows, printed = [], []
    for name in ("toy_full", "toy_slice", "synth_small"):
        case = np.load(HERE / f"{name}.npz", allow_pickle=True)
        for fw, bw, afw, abw, txt in tables_from_summary(case["summary"]):
            rows.append((fw, bw, afw, abw))
            printed.append(txt)
    n_fixture = len(rows)
    rows += random_tables(np.random.default_rng(20186), 1500)
    n_random = len(rows)
    rows += tie_tables()
    printed += [""] * (len(rows) - n_fixture)
    t = np.array(rows, dtype=np.int64)
    p = np.array([boost_fisher(*map(int, r)) for r in rows])
    np.savez_compressed(HERE / "fisher_boost.npz", fw=t[:, 0], bw=t[:, 1], alt_fw=t[:, 2], alt_bw=t[:, 3], p_boost=p,
                        printed_by_standin=np.array(printed), n_from_fixtures=np.int64(n_fixture), n_before_ties=np.int64(n_random),
                        scipy_version=np.array(scipy.__version__))
    print(f"fisher_boost.npz: {n_fixture} call rows of the golden cases + {n_random - n_fixture} seeded tables + {len(rows) - n_random} "
          f"symmetric (tie-prone) tables, scipy {scipy.__version__}")


if __name__ == "__main__":
    main()
