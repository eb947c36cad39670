"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): noise tables and call sets bit-exact (thresholds compared as float bit
patterns, counts/states as integers, call keys as integers); p-values within 1e-12 relative where the
reference's own 1-(1-x) arithmetic is conditioned to that (p >= 1e-4), and within two
quanta of 2^-53 below.
"""
import numpy as np
import pytest
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent

from oracle import pyoracle
from tests import synth

pytestmark = pytest.mark.gpu

P_REL_TOL = 1e-12     # relative tolerance on p where well conditioned (p >= 1e-4)
P_ABS_QUANTA = 2.3e-16  # two quanta of the 1-(1-x) grid


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def oracle_noise(counts, pos_id, U, C, cut):
    rows, row_off = pyoracle.dense_to_rows(synth.to_oracle_layout(counts), pos_id)
    return pyoracle.noise_estimate(rows, row_off, U, C, cut)


def check_noise(got, want, pos_id):
    thr_w = want["thr"][pos_id]          # per slot
    assert np.array_equal(np.isnan(got["thr"]), np.isnan(thr_w))
    m = ~np.isnan(thr_w)
    assert np.array_equal(bits(got["thr"])[m], bits(thr_w)[m]), "threshold floats differ"
    assert np.array_equal(got["count"].astype(np.int64), want["count"][pos_id].astype(np.int64))
    assert np.array_equal(got["nrec"].astype(np.int64), want["nrec"][pos_id].astype(np.int64))
    present_w = want["germ_present"][pos_id]
    assert np.array_equal(got["germ_state"] > 0, present_w > 0)
    mm = present_w > 0
    # the oracle keeps the reference's double (a widened float, -888 or 0): compare as float bit patterns
    assert np.array_equal(bits(got["germ_val"])[mm], bits(want["germ_val"][pos_id].astype(np.float32))[mm])


@pytest.mark.parametrize("S,n_amp,depth,C,cut,seed", [
    (5, 30, 700, 0.002, 100, 11),
    (40, 20, 5000, 0.002, 100, 12),
    (33, 25, 2000, 0.001, 100, 13),
    (7, 12, 300, 0.005, 50, 14),
    (1, 10, 1500, 0.002, 100, 15),
    (100, 8, 2000, 0.0035, 150, 16),
    (20, 10, 160, 0.002, 100, 17),   # marginal coverage: the 0.338*N rule decides
])
@pytest.mark.parametrize("variant", [1, 0, 4, 7, 8])
def test_noise_matches_oracle(ctx, variant, S, n_amp, depth, C, cut, seed):
    _, slots, pos_id, U = synth.make_panel(n_amp, seed=seed)
    P = len(slots)
    counts, _ = synth.make_counts(S, P, depth=depth, seed=seed, pos_id=pos_id)
    nxt, head = ctx_twins(pos_id)
    ctx.set_noise_kernel(variant)
    try:
        got = ctx.estimate_thresholds(counts, C, cut, nxt, head)
    finally:
        ctx.set_noise_kernel(-1)
    want = oracle_noise(counts, pos_id, U, np.float32(C), cut)
    check_noise(got, want, pos_id)


def ctx_twins(pos_id):
    from amplisolve_b200 import twin_links
    return twin_links(pos_id)


@pytest.mark.parametrize("variant", [1, 4, 6, 0, 7, 8])
@pytest.mark.parametrize("seed", [71, 72, 73])
def test_noise_twin_pairs_with_different_rows(ctx, variant, seed):
    """Twin pairs reduced inside the streaming kernel (exact merge of two per-slot states), by the pair kernel (pairs
    cut by a CTA tile, variants without the in-tile path) and by the general kernel (three enumerations): the rows of
    the two slots differ and either slot can hold the pair's first qualifying Germ_Max record."""
    _, slots, pos_id, U = synth.make_panel(50, seed=seed, overlap_frac=0.8, amp_len=(20, 60))
    pos_id = pos_id.copy()
    pos_id[-3] = pos_id[5]          # a position enumerated three times
    P = len(slots)
    counts, _ = synth.make_counts(37, P, depth=1200, seed=seed, pos_id=pos_id, ragged_twins=True)
    nxt, head = ctx_twins(pos_id)
    ctx.set_noise_kernel(variant)
    try:
        got = ctx.estimate_thresholds(counts, 0.002, 100, nxt, head)
    finally:
        ctx.set_noise_kernel(-1)
    _, dense_id = np.unique(pos_id, return_inverse=True)
    check_noise(got, oracle_noise(counts, dense_id.astype(np.int32), dense_id.max() + 1, np.float32(0.002), 100),
                dense_id.astype(np.int32))


def test_noise_no_twins_ragged_and_empty(ctx):
    for P in (1, 31, 127, 129, 1000):
        pos_id = np.arange(P, dtype=np.int32)
        counts, _ = synth.make_counts(9, P, depth=900, seed=P)
        got = ctx.estimate_thresholds(counts, 0.002, 100)
        check_noise(got, oracle_noise(counts, pos_id, P, np.float32(0.002), 100), pos_id)
    # all rows absent: N = 0 -> 0 < 0 is false -> 0/0 -> NaN -> "-1_-1" (EE:1765-1770)
    counts = np.full((4, 2, 50, 4), 0xFFFFFFFF, dtype=np.uint32)
    got = ctx.estimate_thresholds(counts, 0.002, 100)
    assert np.isnan(got["thr"]).all() and (got["nrec"] == 0).all() and (got["germ_state"] == 0).all()
    # no normals at all
    got = ctx.estimate_thresholds(np.zeros((0, 2, 10, 4), np.uint32), 0.002, 100)
    assert np.isnan(got["thr"]).all()


def test_noise_counts_beyond_2_pow_24(ctx):
    """int -> float conversion is inexact above 2^24: the division-free filter must round like float()."""
    P = 400
    pos_id = np.arange(P, dtype=np.int32)
    counts, _ = synth.make_counts(6, P, depth=3000, seed=77, big_rate=0.5)
    got = ctx.estimate_thresholds(counts, 0.002, 100)
    check_noise(got, oracle_noise(counts, pos_id, P, np.float32(0.002), 100), pos_id)


def test_noise_filter_boundary_exhaustive(ctx):
    """Every (alt, depth) pair around the 5 % boundary for depths 100..4000: one slot per pair, one normal,
    so count is 0 or 1 and tells whether the record passed (EE:1613-1615)."""
    depths = np.arange(100, 4001)
    cases = []
    for D in depths:
        k0 = D // 20
        for k in (k0 - 1, k0, k0 + 1):
            if 0 <= k <= D:
                cases.append((D, k))
    cases = np.array(cases)
    P = len(cases)
    counts = np.zeros((1, 2, P, 4), dtype=np.uint32)
    counts[0, 0, :, 0] = cases[:, 0] - cases[:, 1]   # ref A
    counts[0, 0, :, 2] = cases[:, 1]                 # alt G on the forward strand
    counts[0, 1, :, 0] = 500                         # clean reverse strand
    pos_id = np.arange(P, dtype=np.int32)
    got = ctx.estimate_thresholds(counts, 0.002, 100)
    check_noise(got, oracle_noise(counts, pos_id, P, np.float32(0.002), 100), pos_id)
    assert 0 < got["count"][:, 2].sum() < P


def test_thresholds_caller_view_all_six_decimal_values(ctx):
    """The noise table crosses to the caller as "%f" text: every threshold is n/10^6 after the round trip.
    Exhaustive over the floats nearest to every n in 0..60000 plus random floats, against the oracle's
    sprintf/strtof."""
    import torch
    n = np.arange(0, 60001)
    base = (n / 1e6).astype(np.float32)
    rng = np.random.default_rng(5)
    vals = np.concatenate([base, np.nextafter(base, np.float32(1)), np.nextafter(base, np.float32(-1)),
                           ((n + 0.5) / 1e6).astype(np.float32),
                           rng.uniform(0, 1.0, 200000).astype(np.float32), np.array([np.nan], np.float32)])
    view = ctx.thresholds_caller_view_dev(torch.from_numpy(vals).cuda()).cpu().numpy()
    want = pyoracle.thr_as_caller_sees(np.where(np.isnan(vals), np.float32(0.01), vals))
    assert np.array_equal(bits(view), bits(want))


def test_kf_gammaq_grid(ctx):
    rng = np.random.default_rng(3)
    s = np.concatenate([np.arange(1, 200), rng.integers(200, 60000, 3000)]).astype(np.float64)
    ratio = rng.choice([0.01, 0.1, 0.5, 0.9, 0.99, 1.0, 1.000001, 1.01, 1.5, 3.0, 10.0], size=s.size)
    z = s * ratio
    got = ctx.kf_gammaq(s, z)
    want = np.array([pyoracle.kf_gammaq(a, b) for a, b in zip(s, z)])
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    # libdevice exp/log vs glibc: < 1 ulp each on an exponent of magnitude <= ~1e3 -> relative 1e-12
    assert np.allclose(got[fin], want[fin], rtol=2e-12, atol=1e-300)


def test_poisson_pvalues_grid(ctx):
    rng = np.random.default_rng(4)
    n = 20000
    rd = rng.integers(100, 50000, n).astype(np.int32)
    err = rng.choice(np.array([0.001, 0.002, 0.0035, 0.005, 0.01, 0.0, 0.000123], np.float32), n)
    lam = rd * np.where(err == 0, 0.0010008, err)
    k = np.maximum(0, np.rint(lam * rng.choice([0.0, 0.5, 1.0, 1.2, 1.5, 2.0, 3.0, 6.0], n))).astype(np.int32)
    p, q = ctx.mutation_rules_poisson_quality_score(k, rd, err)
    pw = np.array([pyoracle.poisson_p(a, b, c) for a, b, c in zip(k, rd, err)])
    qw = np.array([pyoracle.poisson_q(a, b, c) for a, b, c in zip(k, rd, err)])
    well = pw >= 1e-4
    assert np.all(np.abs(p[well] - pw[well]) <= P_REL_TOL * pw[well])
    assert np.all(np.abs(p[~well] - pw[~well]) <= P_ABS_QUANTA + P_REL_TOL * np.abs(pw[~well]))
    assert np.allclose(q, qw, rtol=1e-9, atol=1e-9)
    # err == -1 -> Q = -888 (VC:3844-3849)
    _, q2 = ctx.mutation_rules_poisson_quality_score([5], [1000], [-1.0])
    assert q2[0] == -888.0


def oracle_calls(counts, pos_id, U, ref_u, thr_u, cut):
    rows, row_off = pyoracle.dense_to_rows(synth.to_oracle_layout(counts), pos_id)
    return pyoracle.call_variants(rows, row_off, U, ref_u, thr_u, cut), rows, row_off


def check_calls(got, want, rows_slot):
    """want: oracle calls (sample, row, pos_id, alt); rows_slot[sample][row] = slot of that file row."""
    key_w = np.array([(c["sample"], rows_slot[c["sample"]][c["row"]], c["alt"]) for c in want], dtype=np.int64).reshape(-1, 3)
    key_g = np.stack([got["sample"], got["slot"], got["alt"]], axis=1).astype(np.int64).reshape(-1, 3)
    order = np.lexsort((key_w[:, 2], key_w[:, 1], key_w[:, 0]))
    key_w, want = key_w[order], want[order]
    assert key_g.shape == key_w.shape and np.array_equal(key_g, key_w), "call sets differ"
    assert np.array_equal(got["ref"], want["ref"].astype(np.int32))
    for side in ("fw", "bw"):
        pg, pw = got["p_" + side], want["p_" + side]
        well = pw >= 1e-4
        assert np.all(np.abs(pg[well] - pw[well]) <= P_REL_TOL * pw[well])
        assert np.all(np.abs(pg[~well] - pw[~well]) <= P_ABS_QUANTA + P_REL_TOL * pw[~well])
        assert np.allclose(got["q_" + side], want["q_" + side], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("variant", [11, 13, 3, 1, 0, 14, 20])
@pytest.mark.parametrize("T,n_amp,depth,cut,seed", [
    (3, 30, 1800, 100, 21),
    (24, 16, 5000, 100, 22),
    (150, 6, 2000, 100, 23),
    (8, 10, 50000, 100, 24),
    (5, 10, 400, 50, 25),
])
def test_calls_match_oracle(ctx, variant, T, n_amp, depth, cut, seed):
    _, slots, pos_id, U = synth.make_panel(n_amp, seed=seed)
    P = len(slots)
    normals, ref = synth.make_counts(12, P, depth=depth, seed=seed, pos_id=pos_id)
    tumours, _ = synth.make_counts(T, P, depth=depth, seed=seed + 1000, ref=ref, pos_id=pos_id, somatic_rate=0.01)
    ref = ref.copy()
    ref[::97] = 4   # 'N' / lower-case reference bases are never called (VC:3290-3293)
    # thresholds from the oracle's noise model, through the "%f" hand-over
    nz = oracle_noise(normals, pos_id, U, np.float32(0.002), 100)
    thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(nz["thr"]), np.float32(0.01), nz["thr"]))
    thr_u[::53, 1, :] = 0.0    # err == 0 -> 0.0010008 (VC:3852-3856)
    thr_u[::71, 2, 0] = -1.0   # err == -1 -> Q = -888, never a call (VC:3844-3849)
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    ref_slots = ref_u[pos_id]
    ctx.set_call_kernel(variant)
    try:
        got = ctx.call_variants(tumours, ref_slots, thr_u[pos_id], cut)
    finally:
        ctx.set_call_kernel(-1)
    want, rows, row_off = oracle_calls(tumours, pos_id, U, ref_u, thr_u, cut)
    present = tumours[:, 0, :, 0] != 0xFFFFFFFF
    rows_slot = [np.nonzero(present[s])[0] for s in range(T)]
    assert len(want) > 0
    check_calls(got, want, rows_slot)


@pytest.mark.parametrize("variant", [13, 14, 11, 20])
def test_integer_prescreen_boundaries(ctx, variant):
    """The integer pre-screen of the staged caller (umulhi(depth, floor(e * 2^32)) >= max(k, 2) drops the pair in the
    scan) must never drop a pair the exact screen m = rn(depth * e) >= k && m > 1 of VC:3728 keeps.  Records sit on and
    around the boundary k = depth * e for thresholds where the product is an exact integer (e = 2^-9, 2^-7), where it is
    not (0.002, 0.0010008 through err == 0), for m <= 1 (tiny e), for e = -1 and for e close to 1; the call set has to
    equal the straightforward kernel's (variant 0, no screens at all) and the oracle's."""
    rng = np.random.default_rng(77)
    es = np.array([2.0 ** -9, 2.0 ** -7, 0.002, 0.0, 0.01, 1e-6, 3e-4, -1.0, 0.05, 0.999], np.float32)
    P, T = 640, 40
    ref = rng.integers(0, 4, P).astype(np.uint8)
    thr = np.empty((P, 4, 2), np.float32)
    thr[:, :, 0] = es[rng.integers(0, len(es), (P, 4))]
    thr[:, :, 1] = np.where(rng.random((P, 4)) < 0.7, thr[:, :, 0], es[rng.integers(0, len(es), (P, 4))])
    same = rng.random(P) < 0.4      # one threshold for the whole slot: the pre-screen's bound (the smallest threshold of the
    thr[same] = es[rng.integers(0, len(es), same.sum())][:, None, None]   # strand) then coincides with every base's own
    tumours = np.zeros((T, 2, P, 4), np.uint32)
    depth = (rng.choice([512, 1024, 2048, 3000, 5000, 65536, 100000, 1 << 20], (T, 2, P)) * rng.integers(1, 4, (T, 2, P))).astype(np.int64)
    for st in (0, 1):
        for b in range(4):
            e = np.where(thr[:, b, st] == 0, np.float32(0.0010008), thr[:, b, st]).astype(np.float64)
            m = depth[:, st] * np.maximum(e, 0)[None, :]
            k = np.floor(m).astype(np.int64) + rng.integers(-2, 4, (T, P))        # on and around the boundary
            k = np.where(rng.random((T, P)) < 0.15, rng.integers(0, 4, (T, P)), k)  # and tiny counts (m <= 1 region)
            tumours[:, st, :, b] = np.clip(k, 0, depth[:, st] // 8)
    for st in (0, 1):   # the reference base takes the rest of the depth
        rest = depth[:, st] - tumours[:, st].sum(-1) + tumours[:, st, np.arange(P), ref]
        tumours[:, st, np.arange(P), ref] = np.maximum(rest, 0)
    pos_id = np.arange(P, dtype=np.int32)
    ctx.set_call_kernel(0)
    try:
        plain = ctx.call_variants(tumours, ref, thr, 100)
        ctx.set_call_kernel(variant)
        got = ctx.call_variants(tumours, ref, thr, 100)
    finally:
        ctx.set_call_kernel(-1)
    assert len(plain) > 500
    assert plain.tobytes() == got.tobytes()
    want, _, _ = oracle_calls(tumours, pos_id, P, ref, thr, 100)
    check_calls(got, want, [np.arange(P) for _ in range(T)])


@pytest.mark.parametrize("variant", [13, 1, 20])
def test_critical_mean_screen(ctx, variant):
    """The second exact screen of the caller (m >= AS_MCRIT[k-1] => the strand test cannot pass, k <= 64): records whose
    mean m = depth * e straddles the critical mean m*(k) of every k = 1..64 by relative offsets from 1e-6 to 1e-1, on both
    strands or on one (the other strand passing comfortably), against the straightforward kernel (no screens) and the
    oracle.  e = 2^-20 makes depth * e exact, so the offsets are what they say."""
    import sys
    sys.path.insert(0, str(ROOT / "scripts"))
    from critical_means import critical_means
    mc = np.array(critical_means()) / (1 + 1e-9)
    offs = np.array([-1e-1, -1e-2, -1e-4, -1e-6, 1e-6, 1e-4, 1e-2, 1e-1])
    ks = np.arange(1, 65)
    P, T = 64 * 2, len(offs)                      # slot = (k, both strands / one strand), sample = offset
    e = np.float32(2.0 ** -20)
    thr = np.full((P, 4, 2), e, np.float32)
    ref = np.zeros(P, np.uint8)                   # reference base A, alt C carries the test
    tumours = np.zeros((T, 2, P, 4), np.uint32)
    for si, off in enumerate(offs):
        depth = np.rint(mc * (1 + off) * 2.0 ** 20).astype(np.int64)      # m = depth * 2^-20
        for mode in (0, 1):
            slots = (ks - 1) * 2 + mode
            tumours[si, 0, slots, 1] = ks
            tumours[si, 0, slots, 0] = depth - ks
            if mode == 0:
                tumours[si, 1, slots, 1] = ks
                tumours[si, 1, slots, 0] = depth - ks
            else:                                  # reverse strand: same k at a tenth of the mean -> passes easily
                tumours[si, 1, slots, 1] = ks
                tumours[si, 1, slots, 0] = np.maximum(depth // 10, 200) - ks
    ctx.set_call_kernel(0)
    try:
        plain = ctx.call_variants(tumours, ref, thr, 100)
        ctx.set_call_kernel(variant)
        got = ctx.call_variants(tumours, ref, thr, 100)
    finally:
        ctx.set_call_kernel(-1)
    assert plain.tobytes() == got.tobytes()
    # below the critical mean the pair is a call, above it is not -- for every k and both layouts
    called = np.zeros((T, P), bool)
    called[got["sample"], got["slot"]] = True
    assert called[offs <= -1e-4].all() and not called[offs >= 1e-4].any()   # (the 1e-6 offsets are below the depth grid at small k)
    want, _, _ = oracle_calls(tumours, np.arange(P, dtype=np.int32), P, ref, thr, 100)
    check_calls(got, want, [np.arange(P) for _ in range(T)])


@pytest.mark.parametrize("tile", [128, 512, 1024])
def test_host_pipelines_across_tile_boundaries(ctx, tile):
    """The _host entry points tile over slots; twin groups that straddle a tile boundary take the gather path."""
    _, slots, pos_id, U = synth.make_panel(60, seed=41, overlap_frac=0.7)
    P = len(slots)
    normals, ref = synth.make_counts(11, P, depth=1500, seed=41, pos_id=pos_id)
    tumours, _ = synth.make_counts(9, P, depth=1500, seed=42, ref=ref, pos_id=pos_id, somatic_rate=0.02)
    nxt, head = ctx_twins(pos_id)
    crossing = sum(1 for p in range(P) if nxt[p] >= 0 and nxt[p] // tile != p // tile)
    assert crossing > 0 or tile == 1024   # (the 1024-slot case exercises multi-tile uploads even without a crossing pair)
    ctx.set_host_tile_slots(tile)
    try:
        got = ctx.estimate_thresholds(normals, 0.002, 100, nxt, head)
        want = oracle_noise(normals, pos_id, U, np.float32(0.002), 100)
        check_noise(got, want, pos_id)
        thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(want["thr"]), np.float32(0.01), want["thr"]))
        ref_u = np.zeros(U, np.uint8)
        ref_u[pos_id] = ref
        calls = ctx.call_variants(tumours, ref_u[pos_id], thr_u[pos_id], 100)
    finally:
        ctx.set_host_tile_slots(0)
    wcalls, _, _ = oracle_calls(tumours, pos_id, U, ref_u, thr_u, 100)
    present = tumours[:, 0, :, 0] != 0xFFFFFFFF
    check_calls(calls, wcalls, [np.nonzero(present[s])[0] for s in range(9)])


def test_wire16_host_entry_points_equal_the_32_bit_ones(ctx):
    """as_noise_estimate_host16 / as_call_variants_host16: same kernels behind a half-size PCIe format; records with a
    count beyond 16 bits travel in the side list of wide records."""
    from amplisolve_b200 import to_wire16
    _, slots, pos_id, U = synth.make_panel(30, seed=51, overlap_frac=0.5)
    P = len(slots)
    normals, ref = synth.make_counts(13, P, depth=4000, seed=51, pos_id=pos_id, big_rate=0.01)
    tumours, _ = synth.make_counts(10, P, depth=4000, seed=52, ref=ref, pos_id=pos_id, somatic_rate=0.02, big_rate=0.01)
    tumours[3, 1, 17, :] = [65534, 1, 0, 0]     # exactly the escape code: must be escaped, not mistaken for it
    normals[5, 0, 40, :] = [65533, 2, 0, 0]     # the largest count the narrow form carries
    nxt, head = ctx_twins(pos_id)
    n16, nw = to_wire16(normals)
    t16, tw = to_wire16(tumours)
    assert len(nw) > 5 and len(tw) > 5 and (n16[5, 0, 40] == [65533, 2, 0, 0]).all() and (t16[3, :, 17] == 0xFFFE).all()
    for tile in (0, 256):
        ctx.set_host_tile_slots(tile)
        try:
            wide = ctx.estimate_thresholds(normals, 0.002, 100, nxt, head, with_view=True)
            narrow = ctx.estimate_thresholds(n16, 0.002, 100, nxt, head, wide_records=nw, with_view=True)
            for k in wide:
                assert np.array_equal(wide[k].view(np.uint8), narrow[k].view(np.uint8)), k
            view = pyoracle.thr_as_caller_sees(np.where(np.isnan(wide["thr"]), np.float32(0.01), wide["thr"]))
            assert np.array_equal(bits(wide["thr_view"]), bits(view))      # the view filled by the host pipeline itself
            a = ctx.call_variants(tumours, ref, view, 100)
            b = ctx.call_variants(t16, ref, view, 100, wide_records=tw)
            assert len(a) > 0 and a.tobytes() == b.tobytes()
        finally:
            ctx.set_host_tile_slots(0)


def test_packed_host_entry_points_equal_the_32_bit_ones(ctx):
    """as_noise_estimate_host_packed / as_call_variants_host_packed: same kernels behind the 8-bytes-per-record PCIe
    format (16-bit major count + three 4-bit minors per strand word); records that do not fit -- variants, noisy
    positions, counts beyond 16 bits -- travel in the side list.  Escaped twin members that straddle an upload tile are
    read back through the host-side decoder."""
    from amplisolve_b200 import pack_counts
    _, slots, pos_id, U = synth.make_panel(30, seed=61, overlap_frac=0.5)
    P = len(slots)
    normals, ref = synth.make_counts(13, P, depth=3000, seed=61, pos_id=pos_id, big_rate=0.01)
    tumours, _ = synth.make_counts(10, P, depth=3000, seed=62, ref=ref, pos_id=pos_id, somatic_rate=0.02, big_rate=0.01)
    tumours[3, :, 17, :] = [[900, 0, 1, 2], [65535, 15, 15, 15]]      # the largest record the packed word carries
    tumours[4, :, 18, :] = [[3, 0, 800, 2], [15, 15, 65535, 15]]
    tumours[5, :, 19, :] = [[16, 0, 900, 0], [1, 0, 900, 0]]          # a minor count of 16: escaped
    tumours[6, :, 20, :] = [[0, 0, 0, 65536], [0, 0, 1, 900]]         # a major count beyond 16 bits: escaped
    normals[5, :, 40, :] = [[7, 7, 7, 7], [900, 1, 1, 1]]             # all equal: the major is base 0
    nxt, head = ctx_twins(pos_id)
    pn, nw = pack_counts(normals)
    pt, tw = pack_counts(tumours)
    assert pn.shape == (13, 2, P) and len(nw) > 5 and len(tw) > 5
    assert pt[3, 1, 17] < 0xFFFFFFFE and pt[4, 1, 18] < 0xFFFFFFFE and (pt[5, :, 19] == 0xFFFFFFFE).all() and (pt[6, :, 20] == 0xFFFFFFFE).all()
    for tile in (0, 256):
        ctx.set_host_tile_slots(tile)
        try:
            wide = ctx.estimate_thresholds(normals, 0.002, 100, nxt, head, with_view=True)
            packed = ctx.estimate_thresholds(pn, 0.002, 100, nxt, head, wide_records=nw, with_view=True)
            for k in wide:
                assert np.array_equal(wide[k].view(np.uint8), packed[k].view(np.uint8)), k
            a = ctx.call_variants(tumours, ref, wide["thr_view"], 100)
            b = ctx.call_variants(pt, ref, wide["thr_view"], 100, wide_records=tw)
            assert len(a) > 0 and a.tobytes() == b.tobytes()
        finally:
            ctx.set_host_tile_slots(0)
    # an all-absent panel and an empty one
    empty = np.full((3, 2, 130, 4), 0xFFFFFFFF, np.uint32)
    pe, we = pack_counts(empty)
    assert len(we) == 0 and (pe == 0xFFFFFFFF).all()
    out = ctx.estimate_thresholds(pe, 0.002, 100)
    assert (out["nrec"] == 0).all() and np.isnan(out["thr"]).all()


@pytest.mark.parametrize("n_c", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("seed,big_rate,ragged", [(101, 0.0, False), (102, 0.01, True)])
def test_noise_sweep_equals_one_pass_per_c_value(ctx, n_c, seed, big_rate, ragged):
    """as_noise_estimate_sweep_dev: the threshold table of every C value from shared passes over the normals is bit-identical
    to the plain noise model run with that value (twin pairs inside and across CTA tiles, longer chains, absent rows, counts
    beyond 2^24), and so are the outputs that do not depend on C."""
    import torch
    _, slots, pos_id, U = synth.make_panel(60, seed=seed, overlap_frac=0.6, amp_len=(20, 70))
    pos_id = pos_id.copy()
    pos_id[-3] = pos_id[5]          # a position enumerated three times (general twin kernel)
    P = len(slots)
    normals, _ = synth.make_counts(21, P, depth=3000, seed=seed, pos_id=pos_id, big_rate=big_rate, ragged_twins=ragged)
    nxt, head = ctx_twins(pos_id)
    d_n = torch.from_numpy(normals.view(np.int32)).cuda()
    d_nxt, d_head = torch.from_numpy(nxt).cuda(), torch.from_numpy(head).cuda()
    c_values = [0.001, 0.0015, 0.002, 0.003, 0.004, 0.005, 0.0075, 0.01][:n_c]
    thr = torch.full((n_c, P, 4, 2), -7.0, dtype=torch.float32, device="cuda")
    out = ctx.alloc_noise_outputs(P)
    ctx.estimate_thresholds_sweep_dev(d_n, c_values, 100, thr, out, d_nxt, d_head)
    torch.cuda.synchronize()
    thr = thr.cpu().numpy()
    for ci, c in enumerate(c_values):
        one = ctx.alloc_noise_outputs(P)
        ctx.estimate_thresholds_dev(d_n, c, 100, one, d_nxt, d_head)
        torch.cuda.synchronize()
        assert np.array_equal(bits(thr[ci]), bits(one["thr"].cpu().numpy())), c
        for k in ("germ_val", "germ_state", "count", "nrec"):
            assert np.array_equal(out[k].cpu().numpy().view(np.uint8), one[k].cpu().numpy().view(np.uint8)), (c, k)
    assert n_c == 1 or not np.array_equal(bits(thr[0]), bits(thr[-1]))


@pytest.mark.parametrize("mode", ["deferred", "deferred-overflow", "in-stage"])
def test_noise_floor_sweep_equals_one_pass_per_c_value(ctx, mode):
    """as_call_variants_sweep_dev (BASELINE configs[3]: C_value 0.001 ... 0.005): one pass over the tumour tensor for all
    threshold tables gives, per table, exactly the call set of the plain caller run with that table -- and that is the
    oracle's for that C_value.  Modes: the deferred scan -> resolve -> series pipeline (default), the same with candidate /
    survivor lists of a handful of entries (everything that does not fit is resolved on the spot), and the in-stage sweep
    kernel of round 1."""
    import torch
    from amplisolve_b200 import CALL_DTYPE
    _, slots, pos_id, U = synth.make_panel(20, seed=91)
    P = len(slots)
    normals, ref = synth.make_counts(14, P, depth=2500, seed=91, pos_id=pos_id)
    tumours, _ = synth.make_counts(37, P, depth=2500, seed=92, ref=ref, pos_id=pos_id, somatic_rate=0.02)
    nxt, head = ctx_twins(pos_id)
    c_values = [0.001, 0.002, 0.003, 0.004, 0.005]
    views = [ctx.estimate_thresholds(normals, c, 100, nxt, head, with_view=True)["thr_view"] for c in c_values]
    d_t = torch.from_numpy(tumours.view(np.int32)).cuda()
    d_ref = torch.from_numpy(ref).cuda()
    d_views = torch.from_numpy(np.stack(views)).cuda()
    cap = 1 << 15
    d_calls = torch.zeros(len(c_values) * cap * CALL_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(len(c_values), dtype=torch.int64, device="cuda")
    if mode == "in-stage":
        ctx.set_call_kernel(11)
    if mode == "deferred-overflow":
        ctx.set_option("deferred_capacity", 7)
    try:
        assert ctx.call_variants_sweep_dev(d_t, d_ref, d_views, 100, d_calls, d_n) == cap
        torch.cuda.synchronize()
    finally:
        ctx.set_call_kernel(-1)
        ctx.set_option("deferred_capacity", 0)
    n = d_n.cpu().numpy()
    lists = d_calls.cpu().numpy().view(CALL_DTYPE).reshape(len(c_values), cap)
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = ref
    sizes = []
    for ci, c in enumerate(c_values):
        got = np.sort(lists[ci, :n[ci]], order=["sample", "slot", "alt"])
        want = ctx.call_variants(tumours, ref, views[ci], 100)          # sorted by (sample, slot, alt)
        assert len(want) > 0 and got.tobytes() == want.tobytes(), c
        sizes.append(len(want))
    assert sizes[0] > sizes[-1]          # a higher noise floor calls less
    # and the oracle, for the two ends of the sweep
    for ci in (0, len(c_values) - 1):
        nz = oracle_noise(normals, pos_id, U, np.float32(c_values[ci]), 100)
        thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(nz["thr"]), np.float32(0.01), nz["thr"]))
        want, _, _ = oracle_calls(tumours, pos_id, U, ref_u, thr_u, 100)
        got = np.sort(lists[ci, :n[ci]], order=["sample", "slot", "alt"])
        present = tumours[:, 0, :, 0] != 0xFFFFFFFF
        check_calls(got, want, [np.nonzero(present[s])[0] for s in range(tumours.shape[0])])


def test_device_merge_of_gathered_calls_equals_numpy_sort(ctx):
    """shard.sort_calls_device (the NCCL path of gather_calls merges the ranks' lists on the device) == the numpy lexsort."""
    import torch
    from amplisolve_b200 import CALL_DTYPE
    from amplisolve_b200.api import sort_calls
    from amplisolve_b200.shard import sort_calls_device
    rng = np.random.default_rng(4)
    n = 50_000
    calls = np.zeros(n, CALL_DTYPE)
    keys = rng.choice(400 * 100_000 * 4, n, replace=False)
    calls["sample"], calls["slot"], calls["alt"] = keys // 400_000, (keys // 4) % 100_000, keys % 4
    calls["p_fw"], calls["q_bw"], calls["ref"] = rng.random(n), rng.random(n), rng.integers(0, 4, n)
    rows = torch.from_numpy(calls.view(np.uint8).reshape(n, CALL_DTYPE.itemsize)).cuda()
    assert sort_calls_device(rows).tobytes() == sort_calls(calls).tobytes()
    assert len(sort_calls_device(rows[:0])) == 0


def test_sort_calls_dev_offsets_and_orders_like_numpy(ctx):
    """as_sort_calls_dev (the last step of the multi-process gather): slot offset applied, reference row order."""
    import torch
    from amplisolve_b200 import CALL_DTYPE
    rng = np.random.default_rng(5)
    n = 70_000
    calls = np.zeros(n, dtype=CALL_DTYPE)
    keys = rng.choice(500 * 40_000 * 4, size=n, replace=False)
    calls["sample"], calls["slot"], calls["alt"] = keys // (40_000 * 4), (keys // 4) % 40_000, keys % 4
    calls["p_fw"] = rng.random(n)
    d_in = torch.from_numpy(calls.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.empty_like(d_in)
    ctx.sort_calls_dev(d_in, n, d_out, slot_offset=123_456)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(CALL_DTYPE)
    want = calls.copy()
    want["slot"] += 123_456
    want = want[np.lexsort((want["alt"], want["slot"], want["sample"]))]
    assert got.tobytes() == want.tobytes()


def test_device_fisher_equals_host_and_boost(ctx):
    """as_fisher_tests_host (one warp per 2x2 table) against the scalar host form as_fisher_test and the
    Boost.Math-pinned fixture: p within 1e-13 relative of the host's (same lgamma values, device exp, lane-ordered sum),
    identical FisherPvalue strings at the two precisions the reference prints, identical YES/NO flags."""
    from amplisolve_b200 import fisher_test
    from tests import golden_util as gu
    g = np.load(gu.GOLDEN / "fisher_boost.npz")
    tables = np.stack([g["fw"], g["bw"], g["alt_fw"], g["alt_bw"]], 1).astype(np.int32)
    extra = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [5, 5, 5, 5], [100, 100, 0, 0], [100, 0, 0, 100], [3, 2000, 3, 0],
                      [70000, 70000, 35000, 35001]], np.int32)
    tables = np.concatenate([tables, extra])
    got = ctx.fisher_tests(tables)
    host = np.array([fisher_test(*t) for t in tables.tolist()])
    normal = host > 1e-290
    assert np.all(np.abs(got - host)[normal] <= 1e-13 * host[normal])
    assert np.all(got[~normal] < 1e-290)
    for digits in (6, 4):
        assert [gu.fmt_g(x, digits) for x in got[normal]] == [gu.fmt_g(x, digits) for x in host[normal]]
    boost = g["p_boost"]
    nb = len(boost)
    N = tables[:nb].astype(np.int64).sum(1)
    # symmetric tables beyond Boost's prime-factorisation branch: the one unpinned corner (tests/test_oracle_golden.py)
    okb = (boost > 1e-300) & ~((np.arange(nb) >= int(g["n_before_ties"])) & (N > 104723))
    assert np.all(np.abs(got[:nb] - boost)[okb] <= 2e-9 * boost[okb])
    assert np.array_equal(got[:nb][N <= 170], boost[N <= 170])      # Boost's factorial-table branch: bit-identical
    for p_value in (np.float32(0.05), np.float32(0.01)):
        assert np.array_equal(got <= p_value, host <= p_value)
        assert np.array_equal((got[:nb] <= p_value)[okb], (boost <= p_value)[okb])
    assert ctx.fisher_tests(np.zeros((0, 4), np.int32)).shape == (0,)


def test_device_pipeline_matches_host_entry_points(ctx):
    """_dev entry points (inputs resident in HBM, torch tensors) == _host entry points, incl. slot ranges."""
    import torch
    from amplisolve_b200 import calls_from_device
    _, slots, pos_id, U = synth.make_panel(24, seed=31)
    P = len(slots)
    normals, ref = synth.make_counts(10, P, depth=2500, seed=31, pos_id=pos_id)
    tumours, _ = synth.make_counts(7, P, depth=2500, seed=32, ref=ref, pos_id=pos_id, somatic_rate=0.01)
    nxt, head = ctx_twins(pos_id)
    host = ctx.estimate_thresholds(normals, 0.002, 100, nxt, head)
    d_norm = torch.from_numpy(normals.view(np.int32)).cuda()
    out = ctx.alloc_noise_outputs(P)
    ctx.estimate_thresholds_dev(d_norm, 0.002, 100, out, torch.from_numpy(nxt).cuda(), torch.from_numpy(head).cuda())
    torch.cuda.synchronize()
    for k in host:
        a, b = out[k].cpu().numpy(), host[k]
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), k
    view = ctx.thresholds_caller_view_dev(out["thr"])
    host_view = pyoracle.thr_as_caller_sees(np.where(np.isnan(host["thr"]), np.float32(0.01), host["thr"]))
    assert np.array_equal(bits(view.cpu().numpy()), bits(host_view))
    want = ctx.call_variants(tumours, ref, host_view, 100)
    d_tum = torch.from_numpy(tumours.view(np.int32)).cuda()
    calls = torch.zeros(48 * 100000, dtype=torch.uint8, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    mid = P // 2
    for rng_ in ((0, mid), (mid, P)):
        ctx.call_variants_dev(d_tum, torch.from_numpy(ref).cuda(), view, 100, calls, n, slot_range=rng_)
    got = calls_from_device(calls, n)
    assert len(got) == len(want) > 0
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_synthetic_generator_round_trip(ctx):
    """The HBM generator (bench input, with duplicated positions) feeds both kernels; checked against the oracle."""
    import torch
    from amplisolve_b200 import calls_from_device
    P, S, T = 3000, 20, 16
    gen = dict(seed=20181, mean_depth=2000.0, absent_rate=0.02, twin_period=3, slot_offset=125 * 40)
    normals, ref = ctx.synth_counts_dev(S, P, **gen)
    tumours, _ = ctx.synth_counts_dev(T, P, somatic_rate=2e-4, sample_offset=1 << 20, want_ref=False, **gen)
    nxt, head = ctx.synth_twin_links_dev(P, seed=20181, slot_offset=125 * 40, twin_period=3)
    out = ctx.alloc_noise_outputs(P)
    ctx.estimate_thresholds_dev(normals, 0.002, 100, out, nxt, head)
    view = ctx.thresholds_caller_view_dev(out["thr"])
    calls = torch.zeros(48 * 200000, dtype=torch.uint8, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.call_variants_dev(tumours, ref, view, 100, calls, n)
    got_calls = calls_from_device(calls, n)
    h_norm = normals.cpu().numpy().view(np.uint32)
    h_tum = tumours.cpu().numpy().view(np.uint32)
    h_ref = ref.cpu().numpy()
    h_head, h_next = head.cpu().numpy(), nxt.cpu().numpy()
    # geometry: heads point backwards, links are mutual, twins carry identical rows and the same reference base
    twins = np.nonzero(h_head != np.arange(P))[0]
    assert 10 < len(twins) < 0.05 * P
    assert np.array_equal(h_next[h_head[twins]], twins) and (h_head[twins] < twins).all()
    assert np.array_equal(h_norm[:, :, twins], h_norm[:, :, h_head[twins]]) and np.array_equal(h_ref[twins], h_ref[h_head[twins]])
    uniq, pos_id = np.unique(h_head, return_inverse=True)
    pos_id = pos_id.astype(np.int32)
    U = len(uniq)
    want = oracle_noise(h_norm, pos_id, U, np.float32(0.002), 100)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    got["count"] = got["count"].view(np.uint32)
    got["nrec"] = got["nrec"].view(np.uint32)
    check_noise(got, want, pos_id)
    thr_u = pyoracle.thr_as_caller_sees(np.where(np.isnan(want["thr"]), np.float32(0.01), want["thr"]))
    ref_u = np.zeros(U, np.uint8)
    ref_u[pos_id] = h_ref
    wcalls, _, _ = oracle_calls(h_tum, pos_id, U, ref_u, thr_u, 100)
    present = h_tum[:, 0, :, 0] != 0xFFFFFFFF
    check_calls(got_calls, wcalls, [np.nonzero(present[s])[0] for s in range(T)])
    rd = h_norm.astype(np.int64).sum(axis=(1, 3))[h_norm[:, 0, :, 0] != 0xFFFFFFFF]
    assert 1200 < np.median(rd) < 3000


def test_call_list_capacity_and_empty_inputs(ctx):
    """AS_EOVERFLOW reports the true count (the wrapper retries with it); no tumours / no slots give no calls."""
    import ctypes as C
    from amplisolve_b200 import api
    _, slots, pos_id, U = synth.make_panel(10, seed=91)
    P = len(slots)
    normals, ref = synth.make_counts(8, P, depth=3000, seed=91, pos_id=pos_id)
    tumours, _ = synth.make_counts(6, P, depth=3000, seed=92, ref=ref, pos_id=pos_id, somatic_rate=0.05)
    nz = oracle_noise(normals, pos_id, U, np.float32(0.002), 100)
    thr = pyoracle.thr_as_caller_sees(np.where(np.isnan(nz["thr"]), np.float32(0.01), nz["thr"]))[pos_id]
    full = ctx.call_variants(tumours, ref, thr, 100)
    assert len(full) > 20
    small = ctx.call_variants(tumours, ref, thr, 100, cap=7)          # retried internally after AS_EOVERFLOW
    assert small.tobytes() == full.tobytes()
    calls = np.zeros(7, dtype=api.CALL_DTYPE)
    n = C.c_int64(0)
    rc = api.lib().as_call_variants_host(ctx._h, tumours.ctypes.data_as(C.c_void_p), 6, P, ref.ctypes.data_as(C.c_void_p),
                                         np.ascontiguousarray(thr).ctypes.data_as(C.c_void_p), 100,
                                         calls.ctypes.data_as(C.c_void_p), 7, C.byref(n))
    assert rc == -5 and n.value == len(full) and b"calls found" in api.lib().as_last_error()
    assert len(ctx.call_variants(np.zeros((0, 2, P, 4), np.uint32), ref, thr, 100)) == 0
    assert len(ctx.call_variants(np.zeros((3, 2, 0, 4), np.uint32), ref[:0], thr[:0], 100)) == 0
    with pytest.raises(api.AmpliSolveError, match="cutoff"):
        ctx.call_variants(tumours, ref, thr, 0)


def test_device_step_can_be_captured_in_a_cuda_graph(ctx):
    """The _dev entry points only enqueue work (kernels, memset nodes, a fork/join onto the side stream), so a whole
    step -- noise model, threshold hand-over, caller -- can be captured once and replayed: the way to run small panels,
    where seven launches cost more than the kernels.  Replays on fresh inputs must equal the eager path."""
    import torch
    from amplisolve_b200 import calls_from_device
    P, S, T = 1000, 12, 9
    gen = dict(mean_depth=1500.0, twin_period=2)
    normals, ref = ctx.synth_counts_dev(S, P, seed=1, **gen)
    tumours, _ = ctx.synth_counts_dev(T, P, seed=1, somatic_rate=0.01, sample_offset=1 << 20, want_ref=False, **gen)
    nxt, head = ctx.synth_twin_links_dev(P, seed=1, twin_period=2)
    out = ctx.alloc_noise_outputs(P)
    view = torch.empty_like(out["thr"])
    calls = torch.zeros(48 * 50000, dtype=torch.uint8, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step():
        ctx.estimate_thresholds_dev(normals, 0.002, 100, out, nxt, head)
        ctx.thresholds_caller_view_dev(out["thr"], view)
        n.zero_()
        ctx.call_variants_dev(tumours, ref, view, 100, calls, n)

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()                                   # warm-up outside capture: scratch allocations, kernel attributes
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step()
    for seed in (2, 3):
        fresh_n, fresh_ref = ctx.synth_counts_dev(S, P, seed=seed, **gen)
        fresh_t, _ = ctx.synth_counts_dev(T, P, seed=seed, somatic_rate=0.01, sample_offset=1 << 20, want_ref=False, **gen)
        normals.copy_(fresh_n); tumours.copy_(fresh_t); ref.copy_(fresh_ref)
        graph.replay()
        torch.cuda.synchronize()
        got_thr, got_calls = out["thr"].clone(), calls_from_device(calls, n)
        step()
        torch.cuda.synchronize()
        assert torch.equal(got_thr.view(torch.int32), out["thr"].view(torch.int32))
        want_calls = calls_from_device(calls, n)
        assert len(want_calls) > 0 and got_calls.tobytes() == want_calls.tobytes()
