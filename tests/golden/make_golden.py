#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the REAL reference (oracle/_ref, compiled from
/root/reference by `make -C oracle ref`).  Runs only in the dev container; the fixtures are committed.

    python tests/golden/make_golden.py

  vc_function_grid.npz   kf_gammaq / mutationRulesPoissonQualityScore / call decision / fisherTest /
                         homopolymerTest of AmpliSolveVariantCalling.cpp evaluated on seeded grids
  toy_full.npz           the whole of Toy_data (BASELINE.json configs[0]; used by the GPU end-to-end test)
  toy_slice.npz          a slice of Toy_data (BED lines, the 5 normals and 3 tumours as dense counts) with the
                         reference's positionSpecificNoise table, Summary_Variant_Info.txt and VCF bodies
  synth_small.npz        a small synthetic panel (tests/synth.py) with the same outputs
Reference bases: no hg19 FASTA exists here, so they follow SURVEY.md Appendix D.3 (consensus of the pileups) for
the toy slice and the generator's own reference for the synthetic case.
"""
from __future__ import annotations

import ctypes as C
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import refrun  # noqa: E402
from tests import aseq_io, synth  # noqa: E402


def function_grids():
    L = C.CDLL(str(ROOT / "oracle" / "_ref" / "libvc_ref_funcs.so"))
    rng = np.random.default_rng(20181)
    # kf_gammaq
    s = np.concatenate([np.arange(1, 301), rng.integers(300, 60000, 1700)]).astype(np.float64)
    ratio = rng.choice([1e-3, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.97, 0.999, 1.0, 1.0000001, 1.001, 1.03, 1.2, 1.5, 2.0, 5.0, 10.0], size=s.size)
    z = s * ratio
    z[:50] = rng.uniform(0, 1, 50)          # z <= 1 branch
    gq = np.empty_like(s)
    L.ref_kf_gammaq_vec(s.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), gq.ctypes.data_as(C.c_void_p), C.c_long(s.size))
    # Q score
    n = 6000
    rd = rng.integers(20, 60000, n).astype(np.int32)
    err = rng.choice(np.array([0.001, 0.002, 0.002189, 0.0035, 0.005, 0.01, 0.02, 0.0, -1.0, 0.000057], np.float32), n)
    lam = rd * np.where(err <= 0, 0.0010008, err)
    k = np.maximum(0, np.rint(lam * rng.choice([0.0, 0.3, 0.9, 1.0, 1.1, 1.3, 1.6, 2.0, 3.0, 8.0], n) + rng.integers(0, 3, n))).astype(np.int32)
    q = np.empty(n, np.float64)
    L.ref_poisson_q_vec(k.ctypes.data_as(C.c_void_p), rd.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p),
                        q.ctypes.data_as(C.c_void_p), C.c_long(n))
    # decisions on strand pairs
    m = 6000
    fw = rng.integers(50, 30000, m).astype(np.int32)
    bw = rng.integers(50, 30000, m).astype(np.int32)
    efw = rng.choice(np.array([0.002, 0.0021, 0.003, 0.01, 0.0], np.float32), m)
    ebw = rng.choice(np.array([0.002, 0.0025, 0.004, 0.01, 0.0], np.float32), m)
    mult = rng.choice([0.5, 1.0, 1.3, 1.6, 2.0, 2.5, 4.0], m)
    kfw = np.maximum(0, np.rint(fw * np.where(efw == 0, 0.0010008, efw) * mult)).astype(np.int32)
    kbw = np.maximum(0, np.rint(bw * np.where(ebw == 0, 0.0010008, ebw) * mult * rng.choice([0.8, 1.0, 1.2], m))).astype(np.int32)
    dec = np.empty(m, np.uint8)
    L.ref_call_decision_vec(kfw.ctypes.data_as(C.c_void_p), fw.ctypes.data_as(C.c_void_p), efw.ctypes.data_as(C.c_void_p),
                            kbw.ctypes.data_as(C.c_void_p), bw.ctypes.data_as(C.c_void_p), ebw.ctypes.data_as(C.c_void_p),
                            C.c_int(100), dec.ctypes.data_as(C.c_void_p), C.c_long(m))
    # Fisher (Boost stand-in) and homopolymer
    L.ref_fisher.restype = C.c_double
    fa = rng.integers(100, 5000, 300)
    fb = rng.integers(100, 5000, 300)
    fc = rng.integers(0, 60, 300)
    fd = rng.integers(0, 60, 300)
    fp = np.array([L.ref_fisher(int(a), int(b), int(c), int(d)) for a, b, c, d in zip(fa, fb, fc, fd)])
    np.savez_compressed(HERE / "vc_function_grid.npz", gq_s=s, gq_z=z, gq=gq, q_k=k, q_rd=rd, q_err=err, q=q,
                        d_kfw=kfw, d_fw=fw, d_efw=efw, d_kbw=kbw, d_bw=bw, d_ebw=ebw, d_cut=np.int32(100), d_call=dec,
                        f_a=fa, f_b=fb, f_c=fc, f_d=fd, f_p=fp)
    print("vc_function_grid.npz:", s.size, "gammaq,", n, "Q,", m, "decisions (", int(dec.sum()), "calls ),", fp.size, "fisher")


def run_reference(workdir, c_value="0.002", cutoff="100"):
    """ee_ref + the unmodified caller on the staged text inputs; returns (noise table, summary, {sample: vcf body})."""
    noise_path, _ = refrun.run_ee_ref(workdir, "panel.bed", "rb_ref.txt", "rb_dup.txt", "N", c_value, cutoff)
    rel = str(noise_path.relative_to(workdir))
    out = refrun.run_vc_ref(workdir, rel, "T", "v", cutoff=int(cutoff), p_value=0.05)
    summary = (out / "Summary_Variant_Info.txt").read_text()
    vcfs = {p.stem: refrun.vcf_body(p).decode() for p in sorted(out.glob("*.vcf"))}
    return noise_path.read_text(), summary, vcfs


def save_case(name, bed_text, ref_letters, normal_names, normals, tumour_names, tumours, c_value, cutoff):
    case = {"bed": bed_text, "ref_letters": "".join(ref_letters), "normal_names": normal_names, "normals": normals,
            "tumour_names": tumour_names, "tumours": tumours}
    with tempfile.TemporaryDirectory(prefix="asg_", dir="/tmp") as td:
        aseq_io.stage_case(td, case)
        noise, summary, vcfs = run_reference(Path(td), c_value, cutoff)
        default_table = refrun.run_ee_ref_default(Path(td), "panel.bed", "rb_ref.txt", "rb_dup.txt", "0.02").read_text()
    np.savez_compressed(HERE / f"{name}.npz", bed=np.array(bed_text), ref_letters=np.array("".join(ref_letters)),
                        normal_names=np.array(normal_names), normals=normals, tumour_names=np.array(tumour_names),
                        tumours=tumours, c_value=np.float32(float(c_value)), cutoff=np.int32(int(cutoff)),
                        noise_table=np.array(noise), summary=np.array(summary), default_table=np.array(default_table),
                        vcf_bodies=np.array([vcfs[n] for n in tumour_names]))
    ncalls = len(summary.splitlines()) - 1
    print(f"{name}.npz: {normals.shape[2]} slots, {len(normal_names)} normals, {len(tumour_names)} tumours, {ncalls} calls,"
          f" noise table {len(noise)} bytes")


def toy_slice(name="toy_slice", n_bed_lines=(0, 22), extra=(150, 160, 330, 345)):
    toy = Path("/root/reference/Toy_data")
    bed_lines = (toy / "AmpliSeq_30genes_Designed-1.bed").read_bytes().decode().splitlines()
    sel = bed_lines[n_bed_lines[0]:n_bed_lines[1]] + bed_lines[extra[0]:extra[1]] + bed_lines[extra[2]:extra[3]]
    if name == "toy_full":      # BASELINE.json configs[0]: the whole of Toy_data
        sel = bed_lines
    bed_text = "\r\n".join(l.rstrip("\r") for l in sel) + "\r\n"      # the toy BED has CRLF line ends
    slots = aseq_io.enumerate_bed_text(bed_text)
    where, pos_id, U = aseq_io.slot_index(slots)
    P = len(slots)
    nfiles = sorted((toy / "NORMAL_ASEQ_DIR").glob("*.ASEQ"))
    tfiles = sorted((toy / "TUMOUR_ASEQ_DIR").glob("*.ASEQ"))
    normals = np.stack([aseq_io.read_aseq_dense(f, where, P)[0] for f in nfiles])
    tumours = np.stack([aseq_io.read_aseq_dense(f, where, P)[0] for f in tfiles])
    # Appendix D.3 reference bases: first maximum of the A,C,G,T totals over all 8 files, N if never seen
    tot = np.zeros((P, 4), dtype=np.int64)
    for arr in (normals, tumours):
        a = np.where(arr == aseq_io.ABSENT, 0, arr).astype(np.int64)
        tot += a.sum(axis=(0, 1))
    upos = np.zeros((U, 4), dtype=np.int64)
    seen = set()
    for i in range(P):              # a twin pair carries the row twice: count each position once
        if pos_id[i] not in seen:
            upos[pos_id[i]] = tot[i]
            seen.add(pos_id[i])
    letters = ["N" if upos[pos_id[i]].sum() == 0 else "ACGT"[int(np.argmax(upos[pos_id[i]]))] for i in range(P)]
    names_n = [f.name[:-len(".PILEUP.ASEQ")] for f in nfiles]
    names_t = [f.name[:-len(".PILEUP.ASEQ")] for f in tfiles]
    save_case(name, bed_text, letters, names_n, normals, names_t, tumours, "0.002", "100")


def synth_small():
    bed, slots, pos_id, U = synth.make_panel(26, seed=404, chroms=("chr2", "chr9", "chrX"))
    P = len(slots)
    normals, ref = synth.make_counts(12, P, depth=5000, seed=404, pos_id=pos_id)
    tumours, _ = synth.make_counts(6, P, depth=1300, seed=405, ref=ref, pos_id=pos_id, somatic_rate=0.03)
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    letters = ["ACGT"[r] for r in ref_u[pos_id]]
    for i in range(0, P, 211):       # a few positions whose reference base is N / lower case (never masked, never called)
        for j in np.nonzero(pos_id == pos_id[i])[0]:
            letters[j] = "N" if (i // 211) % 2 == 0 else "a"
    bed_text = "".join(f"{c}\t{s}\t{e}\tAMPL{i}\trs{i}\tGENE{i % 7}\n" for i, (c, s, e) in enumerate(bed))
    names_n = [f"NS{i:02d}" for i in range(12)]
    names_t = [f"P{i}_TS{i}" for i in range(6)]
    save_case("synth_small", bed_text, letters, names_n, normals, names_t, tumours, "0.0035", "150")


if __name__ == "__main__":
    if not refrun.have_ref():
        refrun.build_ref()
    function_grids()
    toy_slice()
    toy_slice("toy_full")
    synth_small()
