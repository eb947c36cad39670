"""computeCounts (BAM -> *.PILEUP.ASEQ) against oracle/pileup_oracle.py on synthetic BAM files: every CIGAR
operation, both strands, quality / mapping-quality / flag filters, reads across BGZF block and piece boundaries, positions
listed twice, contigs missing on either side, and the C ABI in pieces.  The reference has no source for
this step (binary only): parity with it is unpinned, the conventions are stated in as_pileup.cu."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import pileup_oracle as po
from tests import bam_io

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "amplisolve_b200" / "bin"

REFS = [("chr1", 100000), ("chr2", 50000), ("chrUn", 30000)]


def panel_lines():
    """position file: amplicons as runs of consecutive positions, two overlapping amplicons (positions listed twice), an
    isolated position, a contig that the BAM does not have"""
    lines = []
    for chrom, start, length in (("chr1", 1000, 120), ("chr1", 1100, 110), ("chr1", 5000, 90), ("chr2", 300, 140), ("chr2", 2000, 1),
                                 ("chrZ", 10, 5)):
        lines += [(chrom, start + i, ".", ".", ".") for i in range(length)]
    lines.append(("chr2", 301, "rs1", "A", "G"))  # an annotated line for a position already listed
    return lines


def random_reads(n, seed):
    rng = np.random.default_rng(seed)
    starts = [(0, 1000), (0, 1100), (0, 5000), (1, 300), (1, 2000), (2, 100)]
    reads = []
    for i in range(n):
        ref_id, s = starts[rng.integers(len(starts))]
        pos = max(0, s - 1 + int(rng.integers(-60, 100)))
        cigar, q_len = [], 0
        if rng.random() < 0.2:
            cigar.append(("H", int(rng.integers(1, 10))))
        if rng.random() < 0.3:
            k = int(rng.integers(1, 12)); cigar.append(("S", k)); q_len += k
        for b in range(int(rng.integers(1, 5))):
            if b:
                op = "IDN"[rng.integers(3)]
                k = int(rng.integers(1, 8 if op != "N" else 40)); cigar.append((op, k))
                if op == "I":
                    q_len += k
            k = int(rng.integers(5, 70)); cigar.append(("M=X"[rng.integers(3)] if rng.random() < 0.3 else "M", k)); q_len += k
        if rng.random() < 0.3:
            k = int(rng.integers(1, 12)); cigar.append(("S", k)); q_len += k
        seq = "".join(rng.choice(list("ACGTN"), size=q_len, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        qual = None if rng.random() < 0.05 else rng.integers(0, 42, size=q_len).astype(np.uint8).tobytes()
        flag = int(rng.choice([0, 16, 0, 16, 4, 256, 512, 1024, 2048, 16 | 2048, 1 | 64, 1 | 16 | 128]))
        reads.append(dict(ref_id=ref_id, pos=pos, mapq=int(rng.choice([0, 10, 19, 20, 37, 60])), flag=flag, cigar=cigar, seq=seq, qual=qual,
                          name=f"read{i}"))
    reads.append(dict(ref_id=-1, pos=-1, mapq=0, flag=4, cigar=[], seq="ACGT", qual=None, name="unplaced"))
    return reads


def run_counts(td, bam, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([str(BIN / "computeCounts"), "vcf=positions.txt", f"bam={bam}", "out=aseq"] + args, cwd=td, capture_output=True,
                       text=True, env=e)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def write_positions(td, lines):
    (Path(td) / "positions.txt").write_text("#CHROM\tPOS\tID\tREF\tALT\n" + "".join(f"{c}\t{p}\t{i}\t{r}\t{a}\t.\t.\t.\n" for c, p, i, r, a in lines))


@pytest.mark.parametrize("block_bytes,piece_mb", [(0xFF00, None), (700, "0"), (97, "0")])
@pytest.mark.parametrize("mbq,mrq,mdc", [(20, 20, 20), (0, 0, 1), (30, 37, 5)])
def test_program_output_equals_the_oracle(ctx, tmp_path, block_bytes, piece_mb, mbq, mrq, mdc):
    lines = panel_lines()
    write_positions(tmp_path, lines)
    reads = random_reads(6000, seed=block_bytes + mbq)
    bam_io.write_bam(tmp_path / "S1.bam", REFS, reads, block_bytes=block_bytes)
    env = {"AS_BAM_PIECE_MB": piece_mb} if piece_mb is not None else None  # "0": one BGZF block per piece, records carried over
    out = run_counts(tmp_path, "S1.bam", [f"mbq={mbq}", f"mrq={mrq}", f"mdc={mdc}", "threads=3"], env)
    want = po.render_aseq(lines, po.pileup(REFS, reads, mbq=mbq, mrq=mrq), mdc=mdc)
    got = (tmp_path / "aseq" / "S1.PILEUP.ASEQ").read_text()
    assert want.count("\n") > 300, "the case must produce rows"
    assert got == want
    assert "S1.PILEUP.ASEQ" in out


def test_several_bams_in_one_run(ctx, tmp_path):
    """bam=a.bam,b.bam (an extension: one context and one position file for many samples): each file as if run alone"""
    lines = panel_lines()
    write_positions(tmp_path, lines)
    sets = {name: random_reads(2500, seed=k) for k, name in enumerate(("N1", "N2", "T1"))}
    for name, reads in sets.items():
        bam_io.write_bam(tmp_path / f"{name}.bam", REFS, reads, block_bytes=4000)
    run_counts(tmp_path, "N1.bam,N2.bam,T1.bam", ["mdc=3"])
    for name, reads in sets.items():
        assert (tmp_path / "aseq" / f"{name}.PILEUP.ASEQ").read_text() == po.render_aseq(lines, po.pileup(REFS, reads), mdc=3)


def test_empty_and_headers_only(ctx, tmp_path):
    lines = panel_lines()
    write_positions(tmp_path, lines)
    bam_io.write_bam(tmp_path / "empty.bam", REFS, [])
    run_counts(tmp_path, "empty.bam", ["mdc=1"])
    assert (tmp_path / "aseq" / "empty.PILEUP.ASEQ").read_text() == po.render_aseq(lines, {}, mdc=1)
    run_counts(tmp_path, "empty.bam", ["mdc=0"])  # every line of the position file, all zero
    assert (tmp_path / "aseq" / "empty.PILEUP.ASEQ").read_text().count("\n") == len(lines) + 1
    (tmp_path / "notbam.bam").write_bytes(b"not a BAM file at all, not even gzip" * 3)
    r = subprocess.run([str(BIN / "computeCounts"), "vcf=positions.txt", "bam=notbam.bam", "out=aseq"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "not a BGZF" in r.stdout
    bad = bytearray((tmp_path / "empty.bam").read_bytes())
    bad[30] ^= 0xFF  # inside the deflate stream of the header block
    (tmp_path / "corrupt.bam").write_bytes(bytes(bad))
    r = subprocess.run([str(BIN / "computeCounts"), "vcf=positions.txt", "bam=corrupt.bam", "out=aseq"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and ("does not inflate" in r.stdout or "corrupt" in r.stdout)


def test_c_abi_counts_are_the_count_tensor_of_one_sample(ctx):
    """as_pileup_* through the C ABI (two pieces): the counts come back in the [strand][slot][base] layout of the count tensor"""
    import struct
    from amplisolve_b200 import AmpliSolveError
    reads = random_reads(3000, seed=5)
    stream = bam_io.bam_stream(REFS, reads)
    o = 8 + struct.unpack_from("<i", stream, 4)[0]  # strip the header, index the records
    n_ref = struct.unpack_from("<i", stream, o)[0]
    o += 4
    for _ in range(n_ref):
        o += 4 + struct.unpack_from("<i", stream, o)[0] + 4
    rec = np.frombuffer(stream[o:], np.uint8).copy()
    offs, p = [], 0
    while p < len(rec):
        offs.append(p)
        p += 4 + struct.unpack_from("<i", rec, p)[0]
    offs = np.array(offs, np.int64)
    pos = {}
    for c, q, *_ in panel_lines():
        if c != "chrZ":
            pos.setdefault(c, set()).add(q - 1)
    contigs = sorted(pos)
    slot_pos = np.concatenate([np.array(sorted(pos[c]), np.int32) for c in contigs])
    first = np.cumsum([0] + [len(pos[c]) for c in contigs]).astype(np.int64)
    ref_contig = np.array([contigs.index(n) if n in contigs else -1 for n, _ in REFS], np.int32)
    half = len(offs) // 2
    cut = int(offs[half])
    pieces = [(rec[:cut], offs[:half]), (rec[cut:], offs[half:] - cut)]
    counts, (n_reads, n_bases) = ctx.pileup(pieces, ref_contig, first, slot_pos)
    want = po.pileup(REFS, reads)
    exp = np.zeros_like(counts)
    j = 0
    for c in contigs:
        for q in sorted(pos[c]):
            fw, bw = want.get((c, q), [[0] * 4, [0] * 4])
            exp[0, j], exp[1, j] = fw, bw
            j += 1
    assert np.array_equal(counts, exp)
    assert n_bases == int(exp.sum()) and 0 < n_reads <= len(reads)
    unsorted = slot_pos.copy()
    unsorted[[0, 1]] = unsorted[[1, 0]]
    with pytest.raises(AmpliSolveError):  # argument errors are reported, not executed
        ctx.pileup(pieces, ref_contig, first, unsorted)
    with pytest.raises(AmpliSolveError):
        ctx.pileup([(rec[:cut], offs[:half] + len(rec))], ref_contig, first, slot_pos)
