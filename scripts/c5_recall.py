#!/usr/bin/env python
"""BASELINE.json configs[4] on a slice: ultra-deep (50,000x) samples with 0.5-1 % spiked SNVs.

A slice of the c5 shape (slots x tumours small enough for the reference to finish in seconds) goes through the
compiled reference programs (oracle/_ref: noise model on the normals, then the caller) and through the CUDA path
(C ABI, device-resident).  Reported: the call sets are identical, and the recall of the spiked SNVs (the same for
both, by identity).  The full-shape caller throughput is bench.py's job
(`--slots 100000 --tumours 10000 --depth 50000 --somatic-rate 0.0005 --vaf 0.005 0.01`).

Run on a GPU box:  python scripts/c5_recall.py [out.json]
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from amplisolve_b200 import Context, twin_links  # noqa: E402
from oracle import pyoracle, refrun  # noqa: E402
from tests import aseq_io, synth  # noqa: E402
from tests import golden_util as gu  # noqa: E402

P_AMPLICONS, NORMALS, TUMOURS, DEPTH, SPIKE_RATE = 24, 30, 40, 50000, 1 / 150


def run(out_json=None):
    rng = np.random.default_rng(20185)
    bed, slots, pos_id, U = synth.make_panel(P_AMPLICONS, amp_len=(125, 125), overlap_frac=0.17, seed=20185, chroms=("chr4",))
    P = len(slots)
    normals, ref = synth.make_counts(NORMALS, P, depth=DEPTH, seed=20185, pos_id=pos_id, absent_rate=0.0, low_cov_rate=0.0, edge_rate=0.0)
    tumours, _ = synth.make_counts(TUMOURS, P, depth=DEPTH, seed=20186, ref=ref, pos_id=pos_id, absent_rate=0.0, low_cov_rate=0.0,
                                   edge_rate=0.0, somatic_rate=0.0)
    # spike SNVs at exactly 0.5-1 % VAF (binomial read counts on both strands), truth kept per (sample, position, alt)
    truth = set()
    first_slot = {}
    for i, u in enumerate(pos_id):
        first_slot.setdefault(int(u), i)
    for t in range(TUMOURS):
        for u in np.nonzero(rng.random(U) < SPIKE_RATE)[0]:
            s0 = first_slot[int(u)]
            alt = int((ref[s0] + 1 + rng.integers(0, 3)) % 4)
            vaf = rng.uniform(0.005, 0.01)
            for s in np.nonzero(pos_id == u)[0]:
                if tumours[t, 0, s, 0] == 0xFFFFFFFF:
                    continue
                for strand in range(2):
                    d = int(tumours[t, strand, s].sum())
                    k = int(np.random.default_rng([20187, t, int(u), strand]).binomial(d, vaf))   # same on both twin slots
                    k = min(k, int(tumours[t, strand, s, ref[s]]))
                    tumours[t, strand, s, ref[s]] -= k
                    tumours[t, strand, s, alt] += k
            truth.add((t, int(u), alt))
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    letters = "".join("ACGT"[r] for r in ref_u[pos_id])
    case = {"bed": "".join(f"{c}\t{s}\t{e}\tA{i}\t.\tG\n" for i, (c, s, e) in enumerate(bed)), "ref_letters": letters,
            "normal_names": [f"N{i:02d}" for i in range(NORMALS)], "normals": normals,
            "tumour_names": [f"S{i:02d}_ct" for i in range(TUMOURS)], "tumours": tumours}
    res = {"slots": P, "positions": U, "normals": NORMALS, "tumours": TUMOURS, "depth": DEPTH, "spiked": len(truth)}
    with tempfile.TemporaryDirectory(prefix="c5_", dir="/tmp") as td:
        td = Path(td)
        aseq_io.stage_case(td, case)
        noise_path, _ = refrun.run_ee_ref(td, "panel.bed", "rb_ref.txt", "rb_dup.txt", "N", "0.002", "100")
        out = refrun.run_vc_ref(td, str(noise_path.relative_to(td)), "T", "v", cutoff=100, p_value=0.05)
        case["summary"] = (out / "Summary_Variant_Info.txt").read_text()
        case["noise_table"] = noise_path.read_text()
    ref_rows = gu.golden_call_rows(case)
    name_idx = {n: i for i, n in enumerate(case["tumour_names"])}
    key_of = {(c, p): int(pos_id[i]) for i, (c, p) in enumerate(slots)}
    ref_calls = {(name_idx[r[0]], key_of[(r[1], r[2])], "ACGT".index(r[4])) for r in ref_rows}
    # ---- CUDA path: noise model on the normals in the reference's file order, caller on the tumours
    n_order = pyoracle.hash_iteration_order([f"N/{n}.PILEUP.ASEQ" for n in case["normal_names"]])
    nxt, head = twin_links(pos_id)
    with Context(0) as ctx:
        noise = ctx.estimate_thresholds(np.ascontiguousarray(normals[n_order]), 0.002, 100, nxt, head, with_view=True)
        calls = ctx.call_variants(tumours, ref_u[pos_id], noise["thr_view"], 100)
    ours = {(int(c["sample"]), int(pos_id[c["slot"]]), int(c["alt"])) for c in calls}
    res["reference_calls"] = len(ref_calls)
    res["cuda_calls"] = len(ours)
    res["call_sets_identical"] = ours == ref_calls
    res["recall_reference"] = len(truth & ref_calls) / max(1, len(truth))
    res["recall_cuda"] = len(truth & ours) / max(1, len(truth))
    res["calls_outside_truth"] = len(ours - truth)
    print(json.dumps(res, indent=1))
    if out_json:
        Path(out_json).write_text(json.dumps(res, indent=1) + "\n")
    assert res["call_sets_identical"]
    return res


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else None)
