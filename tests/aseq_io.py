"""TEST INFRASTRUCTURE -- text formats of the reference, read and written from Python for fixtures.

BED enumeration (both ends inclusive, AmpliSolveErrorEstimation.cpp:633-637), .PILEUP.ASEQ rows
(EE:1149, 15 columns) <-> the dense count tensor uint32 [sample][strand][slot][base] of
include/amplisolve_b200.h (k-th row of a position in a file -> k-th slot of that position).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

ABSENT = np.uint32(0xFFFFFFFF)
ASEQ_HEADER = "chr\tpos\tdbsnp\tMAF\tref\talt\tA\tC\tG\tT\tRD\tArs\tCrs\tGrs\tTrs"


def enumerate_bed_text(bed_text: str):
    slots = []
    for line in bed_text.splitlines():
        f = line.split()
        if len(f) < 3:
            continue
        for p in range(int(f[1]), int(f[2]) + 1):
            slots.append((f[0], p))
    return slots


def slot_index(slots):
    """position -> list of its slots (panel order); pos_id per slot; number of unique positions."""
    where: dict = {}
    uniq: dict = {}
    pos_id = np.empty(len(slots), dtype=np.int32)
    for i, key in enumerate(slots):
        where.setdefault(key, []).append(i)
        pos_id[i] = uniq.setdefault(key, len(uniq))
    return where, pos_id, len(uniq)


def read_aseq_dense(path, where, P):
    """One ASEQ file -> uint32 [2][P][4]; rows outside the panel or beyond the position's slots are reported."""
    out = np.full((2, P, 4), ABSENT, dtype=np.uint32)
    seen: dict = {}
    extra = 0
    with open(path, "rb") as fh:
        fh.readline()
        for raw in fh:
            f = raw.decode("ascii", "replace").split()
            if len(f) < 15:
                continue
            key = (f[0], int(f[1]))
            k = seen.get(key, 0)
            seen[key] = k + 1
            lst = where.get(key)
            if lst is None or k >= len(lst):
                extra += 1
                continue
            v = [int(x) for x in f[6:15]]
            tot, rd, rs = v[0:4], v[4], v[5:9]
            assert sum(tot) == rd, f"RD column differs from the sum of the counts in {path}: {raw!r}"
            s = lst[k]
            out[0, s] = [tot[b] - rs[b] for b in range(4)]
            out[1, s] = rs
    return out, extra


def write_aseq(path, slots, counts_s):
    """uint32 [2][P][4] of one sample -> ASEQ text in panel order (absent rows skipped)."""
    present = counts_s[0, :, 0] != ABSENT
    lines = [ASEQ_HEADER]
    fw = counts_s[0].astype(np.int64)
    bw = counts_s[1].astype(np.int64)
    for i in np.nonzero(present)[0]:
        tot = fw[i] + bw[i]
        c, p = slots[i]
        lines.append(f"{c}\t{p}\t.\t.\t.\t.\t{tot[0]}\t{tot[1]}\t{tot[2]}\t{tot[3]}\t{int(tot.sum())}\t"
                     f"{bw[i][0]}\t{bw[i][1]}\t{bw[i][2]}\t{bw[i][3]}")
    Path(path).write_text("\n".join(lines) + "\n")


def write_ref_tables(workdir, slots, ref_letters, stem="rb"):
    """<seed>_panelReferenceBases.txt / <seed>_ampliconDuplicatedPositions.txt stand-ins (EE:657-665)."""
    workdir = Path(workdir)
    seen: dict = {}
    with open(workdir / f"{stem}_ref.txt", "w") as fh:
        for (c, p), r in zip(slots, ref_letters):
            fh.write(f"{c}\t{p}\t{r}\n")
            seen[(c, p)] = seen.get((c, p), 0) + 1
    with open(workdir / f"{stem}_dup.txt", "w") as fh:
        for (c, p), n in sorted(seen.items()):
            if n >= 2:
                fh.write(f"{c}\t{p}\n")
    return f"{stem}_ref.txt", f"{stem}_dup.txt"


def stage_case(workdir, case):
    """Write a fixture case (dict from load_case) as the text inputs the programs read: panel.bed, N/, T/,
    rb_ref.txt, rb_dup.txt."""
    workdir = Path(workdir)
    (workdir / "panel.bed").write_text(case["bed"])
    slots = enumerate_bed_text(case["bed"])
    for sub, names, counts in (("N", case["normal_names"], case["normals"]), ("T", case["tumour_names"], case["tumours"])):
        (workdir / sub).mkdir(exist_ok=True)
        for i, nm in enumerate(names):
            write_aseq(workdir / sub / f"{nm}.PILEUP.ASEQ", slots, counts[i])
    write_ref_tables(workdir, slots, list(case["ref_letters"]))
    return slots


def load_case(npz_path):
    z = np.load(npz_path, allow_pickle=False)
    case = {k: z[k] for k in z.files}
    for k in ("bed", "ref_letters", "noise_table", "summary", "default_table"):
        case[k] = str(case[k])
    case["normal_names"] = [str(x) for x in case["normal_names"]]
    case["tumour_names"] = [str(x) for x in case["tumour_names"]]
    case["vcfs"] = {nm: str(v) for nm, v in zip(case["tumour_names"], case["vcf_bodies"])}
    return case


def write_fasta(workdir, slots, ref_letters, name="ref.fa"):
    """A .fai-indexed FASTA holding the given base at every panel position (sparse file: one line per
    chromosome, only the panel positions are written, everything else reads as NUL)."""
    workdir = Path(workdir)
    top: dict = {}
    for (c, p) in slots:
        top[c] = max(top.get(c, 0), p)
    offsets = {}
    off = 0
    fai = []
    for c, length in top.items():
        off += len(c) + 2                   # ">chrom\n"
        offsets[c] = off
        fai.append(f"{c}\t{length}\t{off}\t{length}\t{length + 1}")
        off += length + 1
    with open(workdir / name, "wb") as fh:
        fh.truncate(off)
        pos = 0
        for c, length in top.items():
            fh.seek(pos)
            fh.write(f">{c}\n".encode())
            pos = offsets[c] + length
            fh.seek(pos)
            fh.write(b"\n")
            pos += 1
        for (c, p), r in zip(slots, ref_letters):
            fh.seek(offsets[c] + p - 1)
            fh.write(r.encode())
    (workdir / (name + ".fai")).write_text("\n".join(fai) + "\n")
    return name
