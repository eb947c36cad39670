#!/usr/bin/env python
"""computeCounts at the size of one sample of BASELINE.json configs[1]: a synthetic amplicon BAM (~41 k panel positions at
~5000x: 1.4 M reads of 150 bases), written here with numpy, piled up by bin/computeCounts and checked against a vectorised
numpy pileup of the same reads (np.add.at) -- every row of the ASEQ file.  Prints one JSON object: wall time, AS_TIMING phases
(inflate on the host threads, pileup on the GPU), reads/s and bases/s.

    python scripts/pileup_bench.py [out.json]        (on a GPU box; ~1 minute, most of it writing the BAM)
"""
import json
import os
import re
import struct
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import bam_io  # noqa: E402

BIN = ROOT / "amplisolve_b200" / "bin"
READ_LEN = 150


def make_reads(n_amplicons, amp_len, depth, seed):
    """fixed-length reads, one 150M block each, starting within +-20 bases of an amplicon start"""
    rng = np.random.default_rng(seed)
    starts = 10_000 + np.arange(n_amplicons, dtype=np.int64) * 1000
    n_reads = int(n_amplicons * amp_len * depth / READ_LEN)
    amp = rng.integers(0, n_amplicons, n_reads)
    pos = (starts[amp] + rng.integers(-20, 21, n_reads)).astype(np.int32)
    flag = np.where(rng.random(n_reads) < 0.5, 16, 0).astype(np.uint16)
    flag[rng.random(n_reads) < 0.01] |= 1024            # duplicates: skipped
    mapq = np.where(rng.random(n_reads) < 0.03, 5, 60).astype(np.uint8)
    seq = rng.integers(0, 4, (n_reads, READ_LEN), dtype=np.uint8)   # base index per read position
    qual = rng.integers(10, 41, (n_reads, READ_LEN), dtype=np.uint8)
    return starts, pos, flag, mapq, seq, qual


def bam_bytes(refs, pos, flag, mapq, seq, qual):
    n = len(pos)
    name_len = 8
    rec_len = 32 + name_len + 4 + READ_LEN // 2 + READ_LEN
    rec = np.zeros((n, 4 + rec_len), np.uint8)
    hdr = np.zeros(n, dtype=[("bs", "<i4"), ("ref", "<i4"), ("pos", "<i4"), ("ln", "u1"), ("mq", "u1"), ("bin", "<u2"), ("nc", "<u2"),
                             ("fl", "<u2"), ("ls", "<i4"), ("nr", "<i4"), ("np", "<i4"), ("tl", "<i4")])
    hdr["bs"], hdr["ref"], hdr["pos"], hdr["ln"], hdr["mq"], hdr["bin"], hdr["nc"] = rec_len, 0, pos, name_len, mapq, 4680, 1
    hdr["fl"], hdr["ls"], hdr["nr"], hdr["np"] = flag, READ_LEN, -1, -1
    rec[:, :36] = hdr.view(np.uint8).reshape(n, 36)
    rec[:, 36:36 + name_len] = np.frombuffer(b"readnam\0", np.uint8)
    rec[:, 44:48] = np.frombuffer(struct.pack("<I", READ_LEN << 4), np.uint8)
    code = np.array([1, 2, 4, 8], np.uint8)[seq]
    rec[:, 48:48 + READ_LEN // 2] = (code[:, 0::2] << 4) | code[:, 1::2]
    rec[:, 48 + READ_LEN // 2:] = qual
    return bam_io.bam_stream(refs, []) + rec.tobytes()


def numpy_pileup(starts, amp_len, pos, flag, mapq, seq, qual, mbq, mrq):
    """counts [2][P][4] over the panel positions (amplicon a covers starts[a] .. starts[a] + amp_len - 1, 0-based)"""
    P = len(starts) * amp_len
    counts = np.zeros((2, P, 4), np.int64)
    ok = ((flag & 0x704) == 0) & (mapq >= mrq)
    g = pos[ok, None].astype(np.int64) + np.arange(READ_LEN)[None, :]
    a = (g - 10_000) // 1000
    off = (g - 10_000) % 1000
    use = (a >= 0) & (a < len(starts)) & (off < amp_len) & (qual[ok] >= mbq)
    slot = a * amp_len + off
    strand = ((flag[ok] & 16) != 0).astype(np.int64)[:, None] * np.ones(READ_LEN, np.int64)[None, :]
    np.add.at(counts, (strand[use], slot[use], seq[ok][use]), 1)
    return counts


def run(out_json=None, n_amplicons=330, amp_len=125, depth=5000, threads=None):
    res = {"workload": f"{n_amplicons} amplicons x {amp_len} bases at ~{depth}x, {READ_LEN}-base reads (one sample of configs[1])"}
    with tempfile.TemporaryDirectory(prefix="pile_", dir="/tmp") as td:
        td = Path(td)
        t = time.time()
        starts, pos, flag, mapq, seq, qual = make_reads(n_amplicons, amp_len, depth, seed=99)
        refs = [("chr1", int(starts[-1]) + 10_000)]
        raw = bam_bytes(refs, pos, flag, mapq, seq, qual)
        if os.environ.get("PILEUP_SORTED"):  # coordinate-sorted, as aligners + samtools sort deliver it
            order = np.argsort(pos, kind="stable")
            pos, flag, mapq, seq, qual = pos[order], flag[order], mapq[order], seq[order], qual[order]
            raw = bam_bytes(refs, pos, flag, mapq, seq, qual)
            res["sorted"] = True
        (td / "S.bam").write_bytes(bam_io.bgzf_compress(raw, level=1))
        with open(td / "positions.txt", "w") as f:
            for s in starts:
                f.write("".join(f"chr1\t{s + 1 + i}\t.\t.\t.\n" for i in range(amp_len)))
        res.update(reads=len(pos), bam_bytes=(td / "S.bam").stat().st_size, uncompressed_bytes=len(raw), setup_s=round(time.time() - t, 1))
        args = [str(BIN / "computeCounts"), "vcf=positions.txt", "bam=S.bam", "out=o", "mdc=1"] + ([f"threads={threads}"] if threads else [])
        walls = []
        for _ in range(3):  # the first run also pays the page cache and the driver
            t = time.perf_counter()
            r = subprocess.run(args, cwd=td, capture_output=True, text=True, env=dict(os.environ, AS_TIMING="1"))
            walls.append(time.perf_counter() - t)
            assert r.returncode == 0, r.stdout + r.stderr
        res["wall_s_runs"] = walls
        res["phases_s"] = {m.group(1): float(m.group(2)) for m in re.finditer(r"AS_TIMING (\S+) ([0-9.]+)", r.stderr)}
        res["stdout"] = r.stdout.strip()
        if os.environ.get("PILEUP_NCU"):  # one --set full capture of the pileup kernel on this BAM (after the timed runs)
            subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:pileup_kernel", "-c", "1",
                            "-f", "-o", os.environ["PILEUP_NCU"]] + args[:1] + args[1:], cwd=td, capture_output=True, text=True)
        # eight samples in one process (bam=a,b,...: the context is created once)
        for k in range(1, 8):
            os.link(td / "S.bam", td / f"S{k}.bam")
        many = args[:2] + ["bam=" + ",".join(["S.bam"] + [f"S{k}.bam" for k in range(1, 8)])] + args[3:]
        t = time.perf_counter()
        r8 = subprocess.run(many, cwd=td, capture_output=True, text=True, env=dict(os.environ, AS_TIMING="1"))
        res["eight_samples_one_process_wall_s"] = time.perf_counter() - t
        assert r8.returncode == 0, r8.stdout + r8.stderr
        res["eight_samples_phases_s"] = {}
        for m in re.finditer(r"AS_TIMING (\S+) ([0-9.]+)", r8.stderr):
            res["eight_samples_phases_s"][m.group(1)] = res["eight_samples_phases_s"].get(m.group(1), 0.0) + float(m.group(2))
        want = numpy_pileup(starts, amp_len, pos, flag, mapq, seq, qual, 20, 20)
        rows = (td / "o" / "S.PILEUP.ASEQ").read_text().splitlines()[1:]
        got = np.array([[int(x) for x in row.split("\t")[6:]] for row in rows], np.int64)   # A C G T RD Ars Crs Grs Trs
        tot = want[0] + want[1]
        keep = tot.sum(1) >= 1
        exp = np.concatenate([tot[keep], tot[keep].sum(1, keepdims=True), want[1][keep]], 1)
        res["rows"] = len(rows)
        res["identical_to_numpy_pileup"] = bool(got.shape == exp.shape and np.array_equal(got, exp))
        res["bases_counted"] = int(tot.sum())
        work = res["phases_s"].get("inflate_and_pileup", min(walls))
        res["reads_per_s_inflate_and_pileup"] = len(pos) / work
        k_s = res["phases_s"].get("capi.pileup.kernel", 0.0)
        if k_s > 0:
            res["pileup_kernel"] = {"ms": 1e3 * k_s, "reads_per_s": len(pos) / k_s, "counted_bases_per_s": res["bases_counted"] / k_s,
                                    "record_bytes_per_s": len(raw) / k_s}
        res["eight_samples_s_per_sample_after_start_up"] = res["eight_samples_phases_s"].get("inflate_and_pileup", 0.0) / 8
        res["inflate_bytes_per_s"] = len(raw) / max(res["phases_s"].get("inflate_busy", 1e-9), 1e-9)
    print(json.dumps(res, indent=1))
    if out_json:
        Path(out_json).write_text(json.dumps(res, indent=1) + "\n")
    assert res["identical_to_numpy_pileup"], "computeCounts differs from the numpy pileup"
    return res


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else None)
