"""Test-data generator: writes BAM files (BGZF container + BAM records, SAM/BAM specification v1 section 4) for the
computeCounts tests and scripts/pileup_bench.py.  Not a checker and not part of the product."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"
SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
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
