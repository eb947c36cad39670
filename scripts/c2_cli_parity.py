#!/usr/bin/env python
"""BASELINE.json configs[1] at full size through the two drop-in programs, against the compiled reference.

  synthetic 30-gene-panel-sized run: ~41 k slots x 40 normals x 96 ctDNA tumours at ~5000x (text ASEQ/BED/FASTA inputs)
  ours      : amplisolve_b200/bin/AmpliSolveErrorEstimation + AmpliSolveVariantCalling (1 B200), AS_TIMING phases
  reference : oracle/_ref/ee_ref (the reference's own functions minus the samtools fork loop) + AmpliSolveVariantCalling
  check     : noise table, Summary_Variant_Info.txt and every VCF byte-identical (VCF minus ##fileDate)

Run on a GPU box (needs oracle/_ref, which travels with the tree):  python scripts/c2_cli_parity.py [out.json]
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


def run(out_json=None):
    t0 = time.time()
    bed, slots, pos_id, U = synth.make_panel(N_AMPLICONS, amp_len=(110, 140), overlap_frac=0.25, seed=20182,
                                             chroms=("chr1", "chr3", "chr7", "chr12", "chr17", "chrX"))
    P = len(slots)
    normals, ref = synth.make_counts(N_NORMALS, P, depth=DEPTH, seed=20182, pos_id=pos_id)
    tumours, _ = synth.make_counts(N_TUMOURS, P, depth=DEPTH, seed=20183, ref=ref, pos_id=pos_id, somatic_rate=0.002)
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    letters = ["ACGT"[r] for r in ref_u[pos_id]]
    chrom = np.array([c for c, _ in slots])
    pos = np.array([p for _, p in slots])
    res = {"slots": P, "unique_positions": U, "normals": N_NORMALS, "tumours": N_TUMOURS, "depth": DEPTH}
    with tempfile.TemporaryDirectory(prefix="c2_", dir="/tmp") as td:
        td = Path(td)
        (td / "panel.bed").write_text("".join(f"{c}\t{s}\t{e}\tAMPL{i}\trs{i}\tG{i % 30}\n" for i, (c, s, e) in enumerate(bed)))
        (td / "N").mkdir()
        (td / "T").mkdir()
        rows_n = sum(write_aseq_fast(td / "N" / f"NORM{i:03d}.PILEUP.ASEQ", chrom, pos, normals[i]) for i in range(N_NORMALS))
        rows_t = sum(write_aseq_fast(td / "T" / f"PT{i:03d}_ctDNA.PILEUP.ASEQ", chrom, pos, tumours[i]) for i in range(N_TUMOURS))
        aseq_io.write_ref_tables(td, slots, letters)
        aseq_io.write_fasta(td, slots, letters)
        res.update(normal_rows=rows_n, tumour_rows=rows_t, setup_s=round(time.time() - t0, 1))
        env = dict(os.environ, AS_TIMING="1")
        binp = ROOT / "amplisolve_b200" / "bin"
        # ---- ours
        t = time.perf_counter()
        r = subprocess.run([str(binp / "AmpliSolveErrorEstimation"), "panel_design=panel.bed", "reference_genome=ref.fa",
                            "germline_dir=N", "C_value=0.002", "coverage_cutoff=100", "default_error=0.01", "output_dir=o"],
                           cwd=td, capture_output=True, text=True, env=env)
        ee_wall = time.perf_counter() - t
        assert r.returncode == 0, r.stdout[-2000:]
        ee_t = timings(r.stderr)
        t = time.perf_counter()
        r = subprocess.run([str(binp / "AmpliSolveVariantCalling"), "errorFile=o/positionSpecificNoise_0.0020.txt", "tumour_dir=T",
                            "output_dir=v", "coverage_cutoff=100", "p_value=0.05"], cwd=td, capture_output=True, text=True, env=env)
        vc_wall = time.perf_counter() - t
        assert r.returncode == 0, r.stdout[-2000:]
        vc_t = timings(r.stderr)
        res["ours"] = {"error_estimation_wall_s": ee_wall, "variant_calling_wall_s": vc_wall, "ee_phases_s": ee_t, "vc_phases_s": vc_t,
                       "aseq_rows_per_s_normals": rows_n / max(ee_t.get("parse_normals", 1e-9), 1e-9),
                       "aseq_rows_per_s_tumours": rows_t / max(vc_t.get("parse_tumours", 1e-9), 1e-9)}
        # ---- reference (compiled from /root/reference into oracle/_ref)
        t = time.perf_counter()
        noise_path, ee_ref_t = refrun.run_ee_ref(td, "panel.bed", "rb_ref.txt", "rb_dup.txt", "N", "0.002", "100", out_rel="ro")
        ref_ee_wall = time.perf_counter() - t
        t = time.perf_counter()
        out = refrun.run_vc_ref(td, "ro/positionSpecificNoise_0.0020.txt", "T", "rv", cutoff=100, p_value=0.05)
        ref_vc_wall = time.perf_counter() - t
        res["reference"] = {"error_estimation_wall_s": ref_ee_wall, "variant_calling_wall_s": ref_vc_wall, "ee_phases_s": ee_ref_t,
                            "note": "single-threaded programs; EE through the fast driver that skips the samtools fork loop "
                                    "(one fork per position in the real program: ~200 s more on 41 k positions)"}
        # ---- parity
        same_noise = (td / "o" / "positionSpecificNoise_0.0020.txt").read_bytes() == noise_path.read_bytes()
        same_summary = (td / "v" / "Summary_Variant_Info.txt").read_bytes() == (out / "Summary_Variant_Info.txt").read_bytes()
        vcfs = sorted(p.name for p in out.glob("*.vcf"))
        same_vcf = all(refrun.vcf_body(td / "v" / n) == refrun.vcf_body(out / n) for n in vcfs)
        n_calls = len((out / "Summary_Variant_Info.txt").read_text().splitlines()) - 1
        res["parity"] = {"noise_table_identical": same_noise, "summary_identical": same_summary, "vcfs_identical": same_vcf,
                         "n_vcfs": len(vcfs), "calls": n_calls, "noise_table_bytes": noise_path.stat().st_size}
        res["speedup_wall"] = {"error_estimation": ref_ee_wall / ee_wall, "variant_calling": ref_vc_wall / vc_wall}
    print(json.dumps(res, indent=1))
    if out_json:
        Path(out_json).write_text(json.dumps(res, indent=1) + "\n")
    assert same_noise and same_summary and same_vcf, "outputs differ from the reference"
    return res


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else None)
