"""Seeded numpy generator of small dense count tensors with the edge cases the parity tests need.

Layout everywhere: uint32 counts[sample][strand][slot][base] (include/amplisolve_b200.h); ABSENT rows are
0xFFFFFFFF in all eight words.  TEST INFRASTRUCTURE (shared by CPU and GPU tests).
"""
from __future__ import annotations

import numpy as np

ABSENT = np.uint32(0xFFFFFFFF)


def make_panel(n_amplicons=40, amp_len=(60, 130), overlap_frac=0.3, seed=1, chroms=("chr1", "chr7", "chrX")):
    """Amplicons on a few contigs; a fraction overlaps its predecessor by 2..10 bp (-> duplicated positions).
    Returns (bed_rows [(chrom,start,end)], slots [(chrom,pos)], pos_id [P] int32, U)."""
    rng = np.random.default_rng(seed)
    bed, slots = [], []
    cursor = {c: 1000 + 1000 * i for i, c in enumerate(chroms)}
    last_end = {c: None for c in chroms}
    for a in range(n_amplicons):
        c = chroms[a * len(chroms) // n_amplicons]
        L = int(rng.integers(amp_len[0], amp_len[1] + 1))
        if last_end[c] is not None and rng.random() < overlap_frac:
            start = last_end[c] - int(rng.integers(2, 11)) + 1
        else:
            start = cursor[c] + int(rng.integers(5, 200))
        end = start + L - 1
        bed.append((c, start, end))
        for p in range(start, end + 1):
            slots.append((c, p))
        last_end[c] = end
        cursor[c] = end
    uniq: dict = {}
    pos_id = np.empty(len(slots), dtype=np.int32)
    for i, key in enumerate(slots):
        pos_id[i] = uniq.setdefault(key, len(uniq))
    return bed, slots, pos_id, len(uniq)


def make_counts(n_samples, P, *, depth=2000, seed=2, ref=None, pos_id=None, somatic_rate=0.0, absent_rate=0.03,
                low_cov_rate=0.05, edge_rate=0.02, big_rate=0.0, ragged_twins=False):
    """Dense counts with: log-normal depth, per-slot strand-specific error rates, germline SNPs, optional
    spiked SNVs, absent rows, low-coverage rows (around the cutoff of 100), rows whose alt fraction sits
    exactly on / next to the 5 % boundary (5/100, 50/1000, 51/1000 ...), zero-depth strands, and
    (big_rate) counts beyond 2^24.  Twin slots of one position get identical rows (as real ASEQ files do)
    except for a few where only the first twin is present."""
    rng = np.random.default_rng(seed)
    if ref is None:
        ref = rng.integers(0, 4, size=P).astype(np.uint8)
        if pos_id is not None:  # one reference base per position: twin slots share it
            first_of: dict = {}
            for p in range(P):
                ref[p] = ref[first_of.setdefault(int(pos_id[p]), p)]
    err = np.where(rng.random((P, 2, 4)) < 0.6, 0.0, np.minimum(0.02, 3e-4 * np.exp(rng.normal(size=(P, 2, 4)))))
    snp = rng.random(P) < 0.01
    snp_alt = (ref + 1 + rng.integers(0, 3, size=P)) % 4
    counts = np.zeros((n_samples, 2, P, 4), dtype=np.uint32)
    slot_mult = np.exp(rng.normal(scale=0.6, size=P))
    for s in range(n_samples):
        rd = np.maximum(0, np.rint(depth * slot_mult * np.exp(rng.normal(scale=0.5, size=P)))).astype(np.int64)
        low = rng.random(P) < low_cov_rate
        rd[low] = rng.integers(150, 260, size=int(low.sum()))
        fw_d = rng.binomial(rd, 0.5)
        d = np.stack([fw_d, rd - fw_d])                       # [2][P]
        zero_strand = rng.random(P) < 0.003
        d[1, zero_strand] = 0
        vaf = np.zeros((P, 4))
        carriers = snp & (rng.random(P) < 0.3)
        vaf[carriers, snp_alt[carriers]] = rng.choice([0.5, 1.0], size=int(carriers.sum()))
        if somatic_rate > 0:
            som = rng.random(P) < somatic_rate
            alt = (ref + 1 + rng.integers(0, 3, size=P)) % 4
            vaf[som, alt[som]] = np.maximum(vaf[som, alt[som]], rng.uniform(0.01, 0.2, size=int(som.sum())))
        # a third of the variants are strand-biased (exercises the Fisher flag, VC:902-910)
        strand_f = np.where(rng.random((P, 2, 1)) < 0.33, rng.uniform(0.25, 1.0, (P, 2, 1)), 1.0)
        rate = np.minimum(1.0, err + vaf[:, None, :] * strand_f)   # [P][2][4]
        rate[np.arange(P), :, ref] = 0.0
        c = np.zeros((2, P, 4), dtype=np.int64)
        for t in range(2):
            remaining = d[t].copy()
            for b in range(4):
                k = np.minimum(rng.binomial(d[t], rate[:, t, b]), remaining)
                c[t, :, b] = k
                remaining -= k
            c[t, np.arange(P), ref] += remaining
        # rows pinned to the 5 % boundary of the AF filter (EE:1613): alt/depth in {5/100, 50/1000, 51/1000, 49/1000}
        edge = np.nonzero(rng.random(P) < edge_rate)[0]
        for p in edge:
            t = int(rng.integers(0, 2))
            D, k = [(100, 5), (1000, 50), (1000, 51), (1000, 49), (120, 6), (2000, 100), (2001, 100), (1999, 100),
                    (10000, 500), (10000, 501)][int(rng.integers(0, 10))]
            a = int((ref[p] + 1 + rng.integers(0, 3)) % 4)
            c[t, p, :] = 0
            c[t, p, a] = k
            c[t, p, ref[p]] = D - k
        if big_rate > 0:
            big = np.nonzero(rng.random(P) < big_rate)[0]
            for p in big:
                base = int(rng.integers(1 << 24, 1 << 27))
                a = int((ref[p] + 1 + rng.integers(0, 3)) % 4)
                for t in range(2):
                    c[t, p, :] = 0
                    c[t, p, a] = int(base * rng.uniform(0.0495, 0.0505))
                    c[t, p, ref[p]] = base - c[t, p, a]
        counts[s] = c.astype(np.uint32)
        absent = rng.random(P) < absent_rate
        counts[s][:, absent, :] = ABSENT
    if pos_id is not None:  # twins carry identical rows; occasionally only the first twin has a row
        first: dict = {}
        for p in range(P):
            u = int(pos_id[p])
            if u in first:
                counts[:, :, p, :] = counts[:, :, first[u], :]
                drop = rng.random(n_samples) < 0.1
                counts[drop, :, p, :] = ABSENT
                if ragged_twins:
                    # API-level stress (an ASEQ file cannot express it): rows that differ between the twins, and samples
                    # where only the SECOND twin has a row, so that the first qualifying Germ_Max record of the pair can
                    # come from either slot
                    jitter = rng.random(n_samples) < 0.3
                    sel = jitter & (counts[:, 0, p, 0] != ABSENT)
                    counts[sel, 0, p, (ref[p] + 1) % 4] += rng.integers(0, 4, size=int(sel.sum())).astype(np.uint32)
                    gone = rng.random(n_samples) < 0.15
                    counts[gone, :, first[u], :] = ABSENT
            else:
                first[u] = p
    return counts, ref


def to_oracle_layout(counts):
    """[S][2][P][4] -> [S][P][8] (fw ACGT, bw ACGT) as oracle.pyoracle.dense_to_rows expects."""
    return np.ascontiguousarray(np.concatenate([counts[:, 0], counts[:, 1]], axis=-1))

