"""TEST INFRASTRUCTURE ONLY -- harness around oracle/_ref (the real reference compiled from /root/reference).

Nothing here is imported by the product (amplisolve_b200/).  Used by tests/ (CPU side), by
tests/golden/make_golden.py (fixture generation) and by bench.py's reference arm.

The reference programs keep paths in fixed 50-byte buffers (AmpliSolveVariantCalling.cpp:317,
:340; AmpliSolveErrorEstimation.cpp:1084), so every run happens with cwd = a scratch dir and
short relative paths.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REFERENCE_ROOT = Path("/root/reference")


def have_ref() -> bool:
    return (REF_DIR / "ee_ref").exists() and (REF_DIR / "AmpliSolveVariantCalling").exists()


def build_ref(opt: str = "-O2") -> None:
    """Compile the reference from where it lies (only possible where /root/reference exists)."""
    if not (REFERENCE_ROOT / "source_codes").exists():
        raise RuntimeError("/root/reference is not present: oracle/_ref can only be built in the dev container")
    subprocess.run(["make", "-C", str(HERE), "ref", f"REF_OPT={opt}"], check=True, capture_output=True)


def enumerate_bed(bed_path) -> list[tuple[str, int]]:
    """Panel slots in BED order, both ends inclusive (AmpliSolveErrorEstimation.cpp:633-637, :2595-2606)."""
    slots = []
    with open(bed_path, "rb") as fh:
        for raw in fh:
            f = raw.decode("ascii", "replace").split()
            if len(f) < 3:
                continue
            chrom, start, end = f[0], int(f[1]), int(f[2])
            for p in range(start, end + 1):
                slots.append((chrom, p))
    return slots


def read_aseq(path) -> list[tuple[str, int, list[int]]]:
    """Rows of a .PILEUP.ASEQ file: (chrom, pos, [A,C,G,T,RD,Ars,Crs,Grs,Trs]); first line is the header."""
    rows = []
    with open(path, "rb") as fh:
        fh.readline()
        for raw in fh:
            f = raw.decode("ascii", "replace").split()
            if len(f) < 15:
                continue
            rows.append((f[0], int(f[1]), [int(x) for x in f[6:15]]))
    return rows


def consensus_refbases(slots, aseq_files) -> dict[tuple[str, int], str]:
    """no hg19 FASTA exists here, so the reference base of every panel
    position is taken as the first maximum of the A,C,G,T totals summed over the given ASEQ
    files ('N' when the position occurs in none)."""
    tot: dict[tuple[str, int], list[int]] = {}
    for path in aseq_files:
        for chrom, pos, c in read_aseq(path):
            t = tot.setdefault((chrom, pos), [0, 0, 0, 0])
            for i in range(4):
                t[i] += c[i]
    out = {}
    for key in dict.fromkeys(slots):
        t = tot.get(key)
        out[key] = "N" if t is None or sum(t) == 0 else "ACGT"[t.index(max(t))]
    return out


def write_ref_tables(workdir, slots, refmap, stem="rb") -> tuple[str, str]:
    """The two files generateReferenceBases would have left behind (EE:657-665):
    <seed>_panelReferenceBases.txt (chrom\\tpos\\tbase per enumerated slot) and
    <seed>_ampliconDuplicatedPositions.txt (chrom\\tpos of positions enumerated >= 2 times)."""
    workdir = Path(workdir)
    ref_name, dup_name = f"{stem}_ref.txt", f"{stem}_dup.txt"
    seen: dict[tuple[str, int], int] = {}
    with open(workdir / ref_name, "w") as fh:
        for chrom, pos in slots:
            fh.write(f"{chrom}\t{pos}\t{refmap[(chrom, pos)]}\n")
            seen[(chrom, pos)] = seen.get((chrom, pos), 0) + 1
    with open(workdir / dup_name, "w") as fh:
        for (chrom, pos), n in sorted(seen.items()):
            if n >= 2:
                fh.write(f"{chrom}\t{pos}\n")
    return ref_name, dup_name


_TIMING = re.compile(r"EE_REF_TIMING setup=(\S+) parse=(\S+) estimate=(\S+) write=(\S+) records=(\d+)")


def run_ee_ref(workdir, panel_rel, ref_rel, dup_rel, germline_dir_rel, c_value, cutoff, out_rel="o"):
    """Run the reference noise model (fast driver).  Returns (noise_table_path, timing dict)."""
    workdir = Path(workdir)
    (workdir / out_rel).mkdir(exist_ok=True)
    cmd = [str(REF_DIR / "ee_ref"), panel_rel, ref_rel, dup_rel, germline_dir_rel, str(c_value), str(cutoff), out_rel,
           f"{out_rel}/list.txt"]
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, check=True)
    m = _TIMING.search(r.stderr)
    timing = dict(zip(("setup", "parse", "estimate", "write"), map(float, m.groups()[:4]))) if m else {}
    if m:
        timing["records"] = int(m.group(5))
    outs = sorted((workdir / out_rel).glob("positionSpecificNoise_*.txt"))
    if not outs:
        raise RuntimeError("ee_ref produced no noise table:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return outs[0], timing


def run_ee_ref_default(workdir, panel_rel, ref_rel, dup_rel, default_error, out_rel="od"):
    workdir = Path(workdir)
    (workdir / out_rel).mkdir(exist_ok=True)
    cmd = [str(REF_DIR / "ee_ref"), "default", panel_rel, ref_rel, dup_rel, str(default_error), out_rel]
    subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, check=True)
    return workdir / out_rel / "positionSpecificNoise_default.txt"


def run_vc_ref(workdir, error_file_rel, tumour_dir_rel, out_rel="v", cutoff=100, p_value=0.05):
    """Run the UNMODIFIED reference caller binary.  Returns the output directory."""
    workdir = Path(workdir)
    if (workdir / out_rel).exists():
        shutil.rmtree(workdir / out_rel)
    (workdir / out_rel).mkdir()
    cmd = [str(REF_DIR / "AmpliSolveVariantCalling"), f"errorFile={error_file_rel}", f"tumour_dir={tumour_dir_rel}",
           f"output_dir={out_rel}", f"coverage_cutoff={cutoff}", f"p_value={p_value}"]
    r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, check=True)
    if not (workdir / out_rel / "Summary_Variant_Info.txt").exists():
        raise RuntimeError("reference caller produced no summary:\n" + r.stdout[-2000:])
    return workdir / out_rel


def vcf_body(path) -> bytes:
    """VCF bytes without the time-dependent ##fileDate line (AmpliSolveVariantCalling.cpp:688)."""
    return b"".join(l for l in open(path, "rb") if not l.startswith(b"##fileDate="))


def stage_toy(workdir) -> dict:
    """Copy the toy inputs into a scratch dir under short names (N/, T/, panel.bed) and derive the
    Appendix D.3 reference-base tables.  Scratch only -- nothing lands in the repo."""
    workdir = Path(workdir)
    toy = REFERENCE_ROOT / "Toy_data"
    shutil.copy(toy / "AmpliSeq_30genes_Designed-1.bed", workdir / "panel.bed")
    for sub, dst in (("NORMAL_ASEQ_DIR", "N"), ("TUMOUR_ASEQ_DIR", "T")):
        (workdir / dst).mkdir(exist_ok=True)
        for f in sorted((toy / sub).glob("*.ASEQ")):
            shutil.copy(f, workdir / dst / f.name)
    os.system(f"chmod -R u+w {workdir}")
    slots = enumerate_bed(workdir / "panel.bed")
    files = sorted((workdir / "N").glob("*.ASEQ")) + sorted((workdir / "T").glob("*.ASEQ"))
    refmap = consensus_refbases(slots, files)
    ref_rel, dup_rel = write_ref_tables(workdir, slots, refmap)
    return {"slots": slots, "refmap": refmap, "ref": ref_rel, "dup": dup_rel}
