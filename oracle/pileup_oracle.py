"""TEST INFRASTRUCTURE -- checker for the pileup path, never imported by the product.

PARITY UNPINNED: the reference ships its BAM -> ASEQ step only as a binary (Pre-compiled_binaries/computeCounts, Mach-O;
Execution_examples.md:16-54), so there is no source to restate and no way to run it here.  This module states the
conventions the product follows (amplisolve_b200/csrc/as_pileup.cu) as a few lines of plain Python and renders the
expected *.PILEUP.ASEQ text (the 15-column format of Toy_data/*/*.PILEUP.ASEQ, which both reference programs parse:
EE:1114-1149, VC:723-752).  The BAM files the tests feed to bin/computeCounts are written by tests/bam_io.py.

A read is used when mapq >= mrq and (flag & skip_flags) == 0; a base counts when its CIGAR op is M, = or X, its quality is
>= mbq (no qualities = 0xFF = counts) and it is A, C, G or T; reverse strand = flag 0x10.  A row is written for every
line of the position file whose counted depth is >= mdc, in file order.
"""

SKIP_FLAGS = 0x704  # unmapped, secondary, QC fail, duplicate


def pileup(refs, reads, mbq=20, mrq=20, skip_flags=SKIP_FLAGS):
    """{(contig name, 0-based position): [[A, C, G, T] forward, [A, C, G, T] reverse]} over every covered position"""
    counts = {}
    for r in reads:
        if r["ref_id"] < 0 or r["ref_id"] >= len(refs) or r["pos"] < 0:
            continue
        if (r["flag"] & skip_flags) or r["mapq"] < mrq:
            continue
        chrom = refs[r["ref_id"]][0]
        strand = 1 if r["flag"] & 0x10 else 0
        q, g = 0, r["pos"]
        qual = r.get("qual")
        for op, n in r["cigar"]:
            if op in "M=X":
                for i in range(n):
                    base = "ACGT".find(r["seq"][q + i])
                    qv = 0xFF if qual is None else qual[q + i]
                    if base >= 0 and qv >= mbq:
                        counts.setdefault((chrom, g + i), [[0] * 4, [0] * 4])[strand][base] += 1
                q += n
                g += n
            elif op in "IS":
                q += n
            elif op in "DN":
                g += n
    return counts


def render_aseq(position_lines, counts, mdc=20):
    """position_lines: [(chrom, 1-based position, id, ref, alt)] in file order -> the text of <name>.PILEUP.ASEQ"""
    out = ["chr\tpos\tdbsnp\tMAF\tref\talt\tA\tC\tG\tT\tRD\tArs\tCrs\tGrs\tTrs\n"]
    for chrom, pos, ident, ref, alt in position_lines:
        fw, bw = counts.get((chrom, pos - 1), [[0] * 4, [0] * 4])
        tot = [fw[i] + bw[i] for i in range(4)]
        rd = sum(tot)
        if rd < mdc:
            continue
        out.append("\t".join([chrom, str(pos), ident, ".", ref, alt] + [str(v) for v in tot] + [str(rd)] + [str(v) for v in bw]) + "\n")
    return "".join(out)
