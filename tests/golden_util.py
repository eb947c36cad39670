"""TEST INFRASTRUCTURE -- shared helpers of the golden-fixture tests (CPU and GPU)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle import pyoracle
from tests import aseq_io

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["toy_slice", "synth_small"]
LETTER_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def load(name):
    case = aseq_io.load_case(GOLDEN / f"{name}.npz")
    slots = aseq_io.enumerate_bed_text(case["bed"])
    where, pos_id, U = aseq_io.slot_index(slots)
    case.update(slots=slots, pos_id=pos_id, U=U)
    # the reference walks its file lists in libstdc++ hash order of "<dir>/<file>" (EE:1081, VC:672)
    nkeys = [f"N/{n}.PILEUP.ASEQ" for n in case["normal_names"]]
    tkeys = [f"T/{n}.PILEUP.ASEQ" for n in case["tumour_names"]]
    case["normal_order"] = pyoracle.hash_iteration_order(nkeys)
    case["tumour_order"] = pyoracle.hash_iteration_order(tkeys)
    # only upper-case A,C,G,T are callable / masked (EE:2668-2673, VC:3290-3293)
    case["ref_code"] = np.array([LETTER_CODE.get(ch, 255) for ch in case["ref_letters"]], dtype=np.uint8)
    return case


def noise_table_lines(case, thr, germ_val, germ_state):
    """Format per-slot noise outputs as the reference's table (EE:2561, EE:2606-2849).  thr float32 [P][4][2]
    (NaN = "-1_-1"), germ_val [P][4], germ_state [P][4] (0 = "-")."""
    where, _, _ = aseq_io.slot_index(case["slots"])
    lines = ["chrom\tposition\treference\tduplicate\tThres_A\tThres_C\tThres_G\tThres_T\tGerm_Max_A\tGerm_Max_C\tGerm_Max_G\tGerm_Max_T"]
    for i, (c, p) in enumerate(case["slots"]):
        letter = case["ref_letters"][i]
        dup = "YES" if len(where[(c, p)]) >= 2 else "NO"
        cells = [pyoracle.format_thr_cell(thr[i, b, 0], thr[i, b, 1], letter == "ACGT"[b]) for b in range(4)]
        germ = [pyoracle.format_germ_cell(float(germ_val[i, b]), int(germ_state[i, b] > 0)) for b in range(4)]
        lines.append("\t".join([c, str(p), letter, dup] + cells + germ))
    return lines


def parse_noise_thresholds(case):
    """thr_view float32 [P][4][2] exactly as the caller parses the golden noise table (std::stof, VC:889-890)."""
    rows = case["noise_table"].splitlines()[1:]
    out = np.empty((len(rows), 4, 2), dtype=np.float32)
    for i, line in enumerate(rows):
        f = line.split("\t")
        for b in range(4):
            a, c = f[4 + b].split("_")
            out[i, b, 0] = np.float32(a)
            out[i, b, 1] = np.float32(c)
    return out


def golden_call_rows(case):
    """Summary rows of the golden run: (sample name, chrom, pos, ref letter, alt letter, RD, FW, BW, k_fw, k_bw,
    Qscore_fw text, Qscore_bw text, FisherPvalue text)."""
    rows = []
    for line in case["summary"].splitlines()[1:]:
        f = line.split("\t")
        rows.append((f[0], f[1], int(f[2]), f[3][0], f[3][3], int(f[4]), int(f[5]), int(f[6]), int(f[8]), int(f[9]),
                     f[14], f[15], f[13]))
    return rows


def fmt_g(x, prec=4):
    """C++ ostream << with setprecision(prec) in default float format == %.{prec}g."""
    return "%.*g" % (prec, x)
