#!/usr/bin/env python
"""Fixture for rows the dense count tensor has no natural place for (VERDICT r01 "missing" item 4):

  * a row listed more often than the panel enumerates its position (the reference inserts every row into its multimap,
    AmpliSolveErrorEstimation.cpp:1241-1245, and tests every row, AmpliSolveVariantCalling.cpp:869-3288), and
  * a row whose RD column is not A+C+G+T (the reference divides by the column: EE:1229-1232, VC:814-817).

Takes the synth_small golden case, edits a few ASEQ rows (PATCHES below -- tests/test_gpu_golden.py applies the same edits),
runs the compiled reference (oracle/_ref) and stores its outputs in irregular_rows.npz.  Run in the dev container:
    python tests/golden/make_irregular.py
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import refrun  # noqa: E402
from tests import aseq_io  # noqa: E402
from tests import golden_util as gu  # noqa: E402


def called_rows(case):
    """(tumour name -> list of (chrom, pos)) of the golden calls of the unpatched case"""
    out = {}
    for r in gu.golden_call_rows(case):
        out.setdefault(r[0], []).append((r[1], r[2]))
    return out


def apply_patches(workdir, case, ee_rd=False):
    """Edits the staged ASEQ files in place; deterministic.  Returns a description of what was done.
    ee_rd: also break the RD column of one NORMAL row (the one deviation this implementation keeps: Germ_Max there)."""
    workdir = Path(workdir)
    where, _, _ = aseq_io.slot_index(case["slots"])
    calls = called_rows(case)
    done = []

    def edit(path, fn):
        lines = Path(path).read_text().split("\n")
        fn(lines)
        Path(path).write_text("\n".join(lines))

    def find(lines, key, nth=0):
        hits = [i for i, l in enumerate(lines) if l.split("\t")[:2] == [key[0], str(key[1])]]
        return hits[nth]

    single = [k for k, v in where.items() if len(v) == 1]
    twin = [k for k, v in where.items() if len(v) == 2]
    n0, n1 = case["normal_names"][1], case["normal_names"][4]
    # 1. a second row for a single-slot position in one normal (counted twice by the reference's noise model)
    key = single[37]
    edit(workdir / "N" / f"{n0}.PILEUP.ASEQ", lambda L: L.insert(find(L, key) + 1, L[find(L, key)]))
    done.append(("normal extra row", n0, key))
    # 2. a third row for a position the panel enumerates twice, with other counts, at the end of the file
    key2 = twin[3]

    def third(L):
        f = L[find(L, key2)].split("\t")
        f[6:15] = ["3", "1200", "2", "0", "1205", "1", "610", "1", "0"]
        L.insert(len(L) - 1, "\t".join(f))
    edit(workdir / "N" / f"{n1}.PILEUP.ASEQ", third)
    done.append(("normal third row of a duplicated position", n1, key2))
    # 3. tumours: a called row repeated (its calls come out twice, at the second row's place in the file)
    t0 = sorted(calls)[0]
    ckey = calls[t0][0]
    edit(workdir / "T" / f"{t0}.PILEUP.ASEQ", lambda L: L.insert(find(L, ckey) + 3, L[find(L, ckey)]))
    done.append(("tumour extra row (called)", t0, ckey))
    # 4. tumours: RD column of a called row off by +9 (AF and RD columns of its calls use the column)
    t1 = sorted(calls)[-1]
    ckey1 = calls[t1][-1]

    def bump(L):
        i = find(L, ckey1)
        f = L[i].split("\t")
        f[10] = str(int(f[10]) + 9)
        L[i] = "\t".join(f)
    edit(workdir / "T" / f"{t1}.PILEUP.ASEQ", bump)
    done.append(("tumour RD column + 9", t1, ckey1))
    if ee_rd:
        # 5. normals: RD column halved on a row with alt reads (Germ_Max of that position uses the column in the reference)
        n2 = case["normal_names"][7]

        def halve(L):
            for i, l in enumerate(L[1:], 1):
                f = l.split("\t")
                if len(f) == 15 and int(f[10]) > 1500 and sorted(int(x) for x in f[6:10])[2] >= 2:
                    f[10] = str(int(f[10]) // 2)
                    L[i] = "\t".join(f)
                    done.append(("normal RD column halved", n2, (f[0], int(f[1]))))
                    return
        edit(workdir / "N" / f"{n2}.PILEUP.ASEQ", halve)
    return done


def shuffle_rows(workdir, case):
    """Files whose rows do not follow the panel enumeration: one tumour file reversed row by row, one with its second half
    first, one normal file reversed.  The reference does not care (it looks every row up by position; its outputs follow the
    file's row order); the loader's cursor does, and the writer must order the calls by rows, not by slots."""
    workdir = Path(workdir)

    def edit(path, fn):
        lines = Path(path).read_text().split("\n")
        head, rows = lines[0], [l for l in lines[1:] if l]
        Path(path).write_text("\n".join([head] + fn(rows)) + "\n")

    t = case["tumour_names"]
    edit(workdir / "T" / f"{t[1]}.PILEUP.ASEQ", lambda r: r[::-1])
    edit(workdir / "T" / f"{t[3]}.PILEUP.ASEQ", lambda r: r[len(r) // 2:] + r[:len(r) // 2])
    edit(workdir / "N" / f"{case['normal_names'][2]}.PILEUP.ASEQ", lambda r: r[::-1])
    return [("tumour rows reversed", t[1]), ("tumour halves swapped", t[3]), ("normal rows reversed", case["normal_names"][2])]


def run_reference(workdir, case):
    noise, _ = refrun.run_ee_ref(workdir, "panel.bed", "rb_ref.txt", "rb_dup.txt", "N", f"{float(case['c_value']):.4f}", str(int(case["cutoff"])))
    out = refrun.run_vc_ref(workdir, str(noise.relative_to(workdir)), "T", "v", cutoff=int(case["cutoff"]), p_value=0.05)
    vcfs = {p.name[:-4]: refrun.vcf_body(p).decode() for p in sorted(out.glob("*.vcf"))}
    return noise.read_text(), (out / "Summary_Variant_Info.txt").read_text(), vcfs


def main():
    case = gu.load("synth_small")
    res = {}
    for tag, ee_rd in (("a", False), ("b", True)):
        with tempfile.TemporaryDirectory(prefix="irr_", dir="/tmp") as td:
            aseq_io.stage_case(td, case)
            done = apply_patches(td, case, ee_rd=ee_rd)
            table, summary, vcfs = run_reference(Path(td), case)
            res[tag] = (table, summary, vcfs, done)
    with tempfile.TemporaryDirectory(prefix="irr_", dir="/tmp") as td:
        aseq_io.stage_case(td, case)
        shuffle_rows(td, case)
        res["c"] = run_reference(Path(td), case) + (None,)
    cn = sorted(res["c"][2])
    names = sorted(res["a"][2])
    np.savez_compressed(HERE / "irregular_rows.npz", shuffled_noise_table=np.array(res["c"][0]), shuffled_summary=np.array(res["c"][1]),
                        shuffled_vcf_names=np.array(cn), shuffled_vcf_bodies=np.array([res["c"][2][n] for n in cn]), noise_table=np.array(res["a"][0]), summary=np.array(res["a"][1]),
                        vcf_names=np.array(names), vcf_bodies=np.array([res["a"][2][n] for n in names]),
                        patches=np.array(repr(res["a"][3])), noise_table_rd=np.array(res["b"][0]), patches_rd=np.array(repr(res["b"][3])))
    base = case["noise_table"].splitlines()
    print("shuffled rows: summary equals the unshuffled one:", res["c"][1] == case["summary"], "| noise table equal:", res["c"][0] == case["noise_table"])
    for tag in ("a", "b"):
        new = res[tag][0].splitlines()
        diff = [i for i, (x, y) in enumerate(zip(base, new)) if x != y]
        print(tag, "noise-table lines that differ from the unpatched case:", len(diff), [new[i].split("\t")[:2] for i in diff[:6]])
    print("calls:", len(case["summary"].splitlines()) - 1, "->", len(res["a"][1].splitlines()) - 1)
    print(res["b"][3])


if __name__ == "__main__":
    main()
