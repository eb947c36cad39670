"""TEST INFRASTRUCTURE -- checker for the pileup path, never imported by the product.

PARITY UNPINNED: the reference ships its BAM -> ASEQ step only as a binary (Pre-compiled_binaries/computeCounts, Mach-O;
Execution_examples.md:16-54), so there is no source to restate and no way to run it here.  This module states the
conventions the product follows (amplisolve_b200/csrc/as_pileup.cu) as a few lines of plain Python, writes the BAM
files the tests feed to bin/computeCounts (BGZF container, SAM/BAM specification v1 section 4), and renders the
expected *.PILEUP.ASEQ text (the 15-column format of Toy_data/*/*.PILEUP.ASEQ, which both reference programs parse:
EE:1114-1149, VC:723-752).

A read is used when mapq >= mrq and (flag & skip_flags) == 0; a base counts when its CIGAR op is M, = or X, its quality is
>= mbq (no qualities = 0xFF = counts) and it is A, C, G or T; reverse strand = flag 0x10.  A row is written for every
line of the position file whose counted depth is >= mdc, in file order.
"""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"
SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
SKIP_FLAGS = 0x704  # unmapped, secondary, QC fail, duplicate
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bam_record(read):
    """read: dict(ref_id, pos (0-based), mapq, flag, cigar [(op letter, length)], seq, qual (bytes or None), name)"""
    name = read.get("name", "r").encode() + b"\0"
    seq = read["seq"]
    l_seq = len(seq)
    packed = bytearray((l_seq + 1) // 2)
    for i, ch in enumerate(seq):
        packed[i >> 1] |= SEQ_CODE[ch] << (0 if i & 1 else 4)
    qual = read.get("qual")
    qual = bytes([0xFF]) * l_seq if qual is None else bytes(qual)
    assert len(qual) == l_seq
    cigar = b"".join(struct.pack("<I", (n << 4) | CIGAR_OPS.index(op)) for op, n in read["cigar"])
    body = struct.pack("<iiBBHHHiiii", read["ref_id"], read["pos"], len(name), read["mapq"], 4680, len(read["cigar"]), read["flag"],
                       l_seq, -1, -1, 0) + name + cigar + bytes(packed) + qual + read.get("aux", b"")
    return struct.pack("<i", len(body)) + body


def bam_stream(refs, reads, text="@HD\tVN:1.6\tSO:unsorted\n"):
    """the uncompressed BAM byte stream: header + records.  refs: [(name, length)]"""
    t = text.encode()
    out = [b"BAM\1", struct.pack("<i", len(t)), t, struct.pack("<i", len(refs))]
    for name, length in refs:
        n = name.encode() + b"\0"
        out += [struct.pack("<i", len(n)), n, struct.pack("<i", length)]
    out += [bam_record(r) for r in reads]
    return b"".join(out)


def bgzf_compress(data, block_bytes=0xFF00, level=6):
    out = []
    for o in range(0, len(data), block_bytes):
        chunk = data[o:o + block_bytes]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25  # 12 header + 6 extra + data + 8 trailer, minus 1
        assert bsize < 65536
        out.append(struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, bsize) + comp +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(BGZF_EOF)
    return b"".join(out)


def write_bam(path, refs, reads, block_bytes=0xFF00, level=6):
    with open(path, "wb") as f:
        f.write(bgzf_compress(bam_stream(refs, reads), block_bytes, level))


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
