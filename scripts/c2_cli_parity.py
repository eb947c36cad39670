#!/usr/bin/env python
"""Program against program on the same text files: the two drop-in executables and the compiled reference.

Default = BASELINE.json configs[1] at full size: synthetic 30-gene-panel-sized run, ~41 k slots x 40 normals x 96 ctDNA
tumours at ~5000x (text ASEQ / BED / FASTA inputs).
  ours      : amplisolve_b200/bin/AmpliSolveErrorEstimation + AmpliSolveVariantCalling (AS_DEVICES selects the GPUs), AS_TIMING phases
  reference : oracle/_ref/ee_ref (the reference's own functions minus the samtools fork loop) + AmpliSolveVariantCalling
  check     : noise table, Summary_Variant_Info.txt and every VCF byte-identical (VCF minus ##fileDate)

Used by tests/test_gpu_configs.py (pytest -m gpu) and by bench.py's e2e_text leg (which also runs a configs[2]-shaped
slice).  Run by hand on a GPU box (needs oracle/_ref, which travels with the tree):  python scripts/c2_cli_parity.py [out.json]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import refrun  # noqa: E402
from tests import aseq_io, synth  # noqa: E402

N_AMPLICONS, N_NORMALS, N_TUMOURS, DEPTH = int(os.environ.get("C2_AMPLICONS", 330)), 40, 96, 5000
BIN = ROOT / "amplisolve_b200" / "bin"


def write_aseq_fast(path, chrom, pos, counts_s):
    present = counts_s[0, :, 0] != 0xFFFFFFFF
    fw = counts_s[0, present].astype(np.int64)
    bw = counts_s[1, present].astype(np.int64)
    tot = fw + bw
    df = pd.DataFrame({"chr": chrom[present], "pos": pos[present], "dbsnp": ".", "MAF": ".", "ref": ".", "alt": ".",
                       "A": tot[:, 0], "C": tot[:, 1], "G": tot[:, 2], "T": tot[:, 3], "RD": tot.sum(axis=1),
                       "Ars": bw[:, 0], "Crs": bw[:, 1], "Grs": bw[:, 2], "Trs": bw[:, 3]})
    df.to_csv(path, sep="\t", index=False)
    return int(present.sum())


def timings(stderr):
    return {m.group(1): float(m.group(2)) for m in re.finditer(r"AS_TIMING (\S+) ([0-9.]+)", stderr)}


def stage(td, n_amplicons=N_AMPLICONS, n_normals=N_NORMALS, n_tumours=N_TUMOURS, depth=DEPTH, amp_len=(110, 140), seed=20182,
          chroms=("chr1", "chr3", "chr7", "chr12", "chr17", "chrX"), somatic_rate=0.002, sample_block=128):
    """Writes panel.bed, ref.fa(.fai), rb_ref.txt / rb_dup.txt (what the reference's fast driver reads instead of forking
    samtools), N/*.ASEQ and T/*.ASEQ under td.  Samples are generated in blocks to bound memory."""
    t0 = time.time()
    td = Path(td)
    bed, slots, pos_id, U = synth.make_panel(n_amplicons, amp_len=amp_len, overlap_frac=0.25, seed=seed, chroms=chroms)
    P = len(slots)
    chrom = np.array([c for c, _ in slots])
    pos = np.array([p for _, p in slots])
    (td / "panel.bed").write_text("".join(f"{c}\t{s}\t{e}\tAMPL{i}\trs{i}\tG{i % 30}\n" for i, (c, s, e) in enumerate(bed)))
    (td / "N").mkdir()
    (td / "T").mkdir()
    ref = None
    rows_n = rows_t = 0
    for s0 in range(0, n_normals, sample_block):
        n = min(sample_block, n_normals - s0)
        blk, ref = synth.make_counts(n, P, depth=depth, seed=seed + 7 * s0, ref=ref, pos_id=pos_id)
        rows_n += sum(write_aseq_fast(td / "N" / f"NORM{s0 + i:03d}.PILEUP.ASEQ", chrom, pos, blk[i]) for i in range(n))
    for s0 in range(0, n_tumours, sample_block):
        n = min(sample_block, n_tumours - s0)
        blk, _ = synth.make_counts(n, P, depth=depth, seed=seed + 1 + 7 * s0, ref=ref, pos_id=pos_id, somatic_rate=somatic_rate)
        rows_t += sum(write_aseq_fast(td / "T" / f"PT{s0 + i:03d}_ctDNA.PILEUP.ASEQ", chrom, pos, blk[i]) for i in range(n))
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    letters = ["ACGT"[r] for r in ref_u[pos_id]]
    aseq_io.write_ref_tables(td, slots, letters)
    aseq_io.write_fasta(td, slots, letters)
    return {"slots": P, "unique_positions": U, "normals": n_normals, "tumours": n_tumours, "depth": depth, "normal_rows": rows_n,
            "tumour_rows": rows_t, "setup_s": round(time.time() - t0, 1)}


def start_service(sock, devices=None):
    """amplisolve_b200_serve on a UNIX socket (a resident CUDA context for the programs); returns the process once it listens"""
    args = [str(BIN / "amplisolve_b200_serve"), f"socket={sock}"]
    if devices is not None:
        args.append("devices=" + ",".join(map(str, devices)))
    srv = subprocess.Popen(args, cwd="/", stderr=subprocess.PIPE, text=True)
    line = srv.stderr.readline()
    if "ready" not in line:
        srv.kill()
        raise RuntimeError("amplisolve_b200_serve did not start: " + line)
    return srv


def run_ours(td, out_ee="o", out_vc="v", devices=None, server=None):
    """Both programs on the staged inputs.  Returns wall times and AS_TIMING phases.  server: socket of a resident service."""
    env = dict(os.environ, AS_TIMING="1")
    env.pop("AS_SERVER", None)
    if server is not None:
        env["AS_SERVER"] = server
    if devices is not None:
        env["AS_DEVICES"] = ",".join(map(str, devices))
    t = time.perf_counter()
    r = subprocess.run([str(BIN / "AmpliSolveErrorEstimation"), "panel_design=panel.bed", "reference_genome=ref.fa",
                        "germline_dir=N", "C_value=0.002", "coverage_cutoff=100", "default_error=0.01", f"output_dir={out_ee}"],
                       cwd=td, capture_output=True, text=True, env=env)
    ee_wall = time.perf_counter() - t
    assert r.returncode == 0, r.stdout[-2000:]
    ee_t = timings(r.stderr)
    t = time.perf_counter()
    r = subprocess.run([str(BIN / "AmpliSolveVariantCalling"), f"errorFile={out_ee}/positionSpecificNoise_0.0020.txt", "tumour_dir=T",
                        f"output_dir={out_vc}", "coverage_cutoff=100", "p_value=0.05"], cwd=td, capture_output=True, text=True, env=env)
    vc_wall = time.perf_counter() - t
    assert r.returncode == 0, r.stdout[-2000:]
    return {"error_estimation_wall_s": ee_wall, "variant_calling_wall_s": vc_wall, "ee_phases_s": ee_t, "vc_phases_s": timings(r.stderr)}


def run_reference(td, out_ee="ro", out_vc="rv"):
    t = time.perf_counter()
    noise_path, ee_ref_t = refrun.run_ee_ref(td, "panel.bed", "rb_ref.txt", "rb_dup.txt", "N", "0.002", "100", out_rel=out_ee)
    ref_ee_wall = time.perf_counter() - t
    t = time.perf_counter()
    refrun.run_vc_ref(td, f"{out_ee}/positionSpecificNoise_0.0020.txt", "T", out_vc, cutoff=100, p_value=0.05)
    ref_vc_wall = time.perf_counter() - t
    return {"error_estimation_wall_s": ref_ee_wall, "variant_calling_wall_s": ref_vc_wall, "ee_phases_s": ee_ref_t,
            "note": "single-threaded programs; EE through the fast driver that skips the samtools fork loop "
                    "(one fork per position in the real program: ~200 s more on 41 k positions)"}


def compare(td, a_ee="o", a_vc="v", b_ee="ro", b_vc="rv"):
    """Byte identity of the noise table, the summary and every VCF (minus ##fileDate) of two runs."""
    td = Path(td)
    name = "positionSpecificNoise_0.0020.txt"
    same_noise = (td / a_ee / name).read_bytes() == (td / b_ee / name).read_bytes()
    same_summary = (td / a_vc / "Summary_Variant_Info.txt").read_bytes() == (td / b_vc / "Summary_Variant_Info.txt").read_bytes()
    vcfs = sorted(p.name for p in (td / b_vc).glob("*.vcf"))
    same_vcf = all(refrun.vcf_body(td / a_vc / n) == refrun.vcf_body(td / b_vc / n) for n in vcfs)
    n_calls = len((td / b_vc / "Summary_Variant_Info.txt").read_text().splitlines()) - 1
    return {"noise_table_identical": same_noise, "summary_identical": same_summary, "vcfs_identical": same_vcf, "n_vcfs": len(vcfs),
            "calls": n_calls, "noise_table_bytes": (td / b_ee / name).stat().st_size}


def run(out_json=None, devices=None, **shape):
    with tempfile.TemporaryDirectory(prefix="c2_", dir="/tmp") as td:
        td = Path(td)
        res = stage(td, **shape)
        res["ours"] = run_ours(td, devices=devices)
        res["ours"]["aseq_rows_per_s_normals"] = res["normal_rows"] / max(res["ours"]["ee_phases_s"].get("parse_normals", 1e-9), 1e-9)
        res["reference"] = run_reference(td)
        res["parity"] = compare(td)
        res["speedup_wall"] = {"error_estimation": res["reference"]["error_estimation_wall_s"] / res["ours"]["error_estimation_wall_s"],
                               "variant_calling": res["reference"]["variant_calling_wall_s"] / res["ours"]["variant_calling_wall_s"]}
    print(json.dumps(res, indent=1))
    if out_json:
        Path(out_json).write_text(json.dumps(res, indent=1) + "\n")
    par = res["parity"]
    assert par["noise_table_identical"] and par["summary_identical"] and par["vcfs_identical"], "outputs differ from the reference"
    return res


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else None)
