"""GPU tests at the level of BASELINE.json's configs (the driver runs these with `pytest -m gpu`):

  configs[1]  synthetic 30-gene panel, 40 normals x 96 tumours at ~5000x, FULL SIZE, through the two drop-in programs:
              noise table, Summary_Variant_Info.txt and all 96 VCFs byte-identical to the compiled reference
  configs[3]  a shard with configs[3]'s dimensions (200 normals x 1000 tumours, C_value 0.001 .. 0.005, 250,000 slots):
              fused sweep == one pass per value == oracle on sampled slots
  configs[4]  a slice of the ultra-deep shape (50,000x, 0.5-1 % spiked SNVs): call set and recall equal the reference's
"""
import numpy as np
import pytest

from oracle import pyoracle, refrun
from tests.test_gpu_parity import bits, check_calls, check_noise, oracle_calls, oracle_noise

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref (the compiled reference) is not present")


@needs_ref
def test_config1_full_size_programs_are_byte_identical_to_the_reference(tmp_path):
    from scripts import c2_cli_parity
    res = c2_cli_parity.run(str(tmp_path / "c2.json"))
    assert res["slots"] > 40_000 and res["normals"] == 40 and res["tumours"] == 96
    par = res["parity"]
    assert par["noise_table_identical"] and par["summary_identical"] and par["vcfs_identical"]
    assert par["n_vcfs"] == 96 and par["calls"] > 10_000


@needs_ref
def test_config4_slice_call_set_and_recall_equal_the_reference(tmp_path):
    from scripts import c5_recall
    res = c5_recall.run(str(tmp_path / "c5.json"))
    assert res["depth"] == 50000 and res["spiked"] > 500
    assert res["call_sets_identical"] and res["cuda_calls"] == res["reference_calls"] > 0
    assert res["recall_cuda"] == res["recall_reference"] > 0.95


C3_SLOTS, C3_NORMALS, C3_TUMOURS = 250_000, 200, 1000
C_VALUES = [0.001, 0.002, 0.003, 0.004, 0.005]


def test_config3_dimension_shard_fused_sweep_equals_per_value_and_oracle(ctx):
    """200 normals x 1000 tumours x 250,000 slots resident in HBM (9.6 GB), the five-value noise-floor sweep as the step."""
    import torch
    from amplisolve_b200 import CALL_DTYPE
    P, S, T = C3_SLOTS, C3_NORMALS, C3_TUMOURS
    gen = dict(seed=20184, mean_depth=2000.0, twin_period=6, absent_rate=0.005)
    normals, ref = ctx.synth_counts_dev(S, P, **gen)
    tumours, _ = ctx.synth_counts_dev(T, P, somatic_rate=2e-4, sample_offset=1 << 20, want_ref=False, **gen)
    nxt, head = ctx.synth_twin_links_dev(P, seed=20184, twin_period=6)
    n_c = len(C_VALUES)
    # ---- noise model: fused sweep vs one pass per value
    thr = torch.empty((n_c, P, 4, 2), dtype=torch.float32, device="cuda")
    out = ctx.alloc_noise_outputs(P)
    ctx.estimate_thresholds_sweep_dev(normals, C_VALUES, 100, thr, out, nxt, head)
    views = ctx.thresholds_caller_view_dev(thr)
    torch.cuda.synchronize()
    one = ctx.alloc_noise_outputs(P)
    for ci, c in enumerate(C_VALUES):
        ctx.estimate_thresholds_dev(normals, c, 100, one, nxt, head)
        torch.cuda.synchronize()
        assert torch.equal(thr[ci].view(torch.int32), one["thr"].view(torch.int32)), c
        for k in ("germ_val", "germ_state", "count", "nrec"):
            assert torch.equal(out[k].view(torch.uint8), one[k].view(torch.uint8)), (c, k)
    # ---- caller: one pass over the tumours for all five tables vs one pass per table
    cap = 1 << 20
    s_calls = torch.zeros(n_c * cap * CALL_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    s_n = torch.zeros(n_c, dtype=torch.int64, device="cuda")
    ctx.call_variants_sweep_dev(tumours, ref, views, 100, s_calls, s_n)
    torch.cuda.synchronize()
    n = s_n.cpu().numpy()
    assert (n < cap).all() and n[0] > n[-1] > 1000
    lists = s_calls.cpu().numpy().view(CALL_DTYPE).reshape(n_c, cap)
    fused = [np.sort(lists[ci, :n[ci]], order=["sample", "slot", "alt"]) for ci in range(n_c)]
    calls = torch.zeros(cap * CALL_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    n1 = torch.zeros(1, dtype=torch.int64, device="cuda")
    for ci in range(n_c):
        n1.zero_()
        ctx.call_variants_dev(tumours, ref, views[ci], 100, calls, n1)
        torch.cuda.synchronize()
        plain = np.sort(calls.cpu().numpy().view(CALL_DTYPE)[: int(n1.item())], order=["sample", "slot", "alt"])
        assert plain.tobytes() == fused[ci].tobytes(), C_VALUES[ci]
    # ---- the oracle on sampled slots (whole twin groups), both ends of the sweep
    rng = np.random.default_rng(7)
    h_next, h_head = nxt.cpu().numpy(), head.cpu().numpy()
    twins = np.nonzero(h_next >= 0)[0]
    pick = np.unique(np.concatenate([rng.choice(P, 160, replace=False), twins[:20], h_next[twins[:20]]]))
    pick = np.unique(np.concatenate([pick, h_head[pick], np.where(h_next[pick] >= 0, h_next[pick], pick)]))
    idx = torch.from_numpy(pick).cuda()
    h_norm = normals[:, :, idx, :].cpu().numpy().view(np.uint32)
    h_tum = tumours[:, :, idx, :].cpu().numpy().view(np.uint32)
    h_ref = ref[idx].cpu().numpy()
    uniq, pos_id = np.unique(h_head[pick], return_inverse=True)
    pos_id = pos_id.astype(np.int32)
    present = h_tum[:, 0, :, 0] != 0xFFFFFFFF
    rows_slot = [np.nonzero(present[s])[0] for s in range(T)]
    for ci in (0, n_c - 1):
        want = oracle_noise(h_norm, pos_id, len(uniq), np.float32(C_VALUES[ci]), 100)
        got = {k: v[idx].cpu().numpy() for k, v in out.items()}
        got["thr"] = thr[ci][idx].cpu().numpy()
        got["count"], got["nrec"] = got["count"].view(np.uint32), got["nrec"].view(np.uint32)
        check_noise(got, want, pos_id)
        thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(want["thr"]), np.float32(0.01), want["thr"]))
        ref_u = np.zeros(len(uniq), np.uint8)
        ref_u[pos_id] = h_ref
        wcalls, _, _ = oracle_calls(h_tum, pos_id, len(uniq), ref_u, thr_u, 100)
        sub = fused[ci][np.isin(fused[ci]["slot"], pick)].copy()
        sub["slot"] = np.searchsorted(pick, sub["slot"])
        sub = np.sort(sub, order=["sample", "slot", "alt"])
        assert len(wcalls) > 0
        check_calls(sub, wcalls, rows_slot)
