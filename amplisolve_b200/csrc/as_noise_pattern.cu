// Noise model for up to 8 values of C in ONE pass over the normals (noise-floor sweep of BASELINE configs[3]), and the
// register-lean form of the plain noise kernel (NC = 1).
//
//   EE = source_codes/AmpliSolveErrorEstimation.cpp.  Same arithmetic as as_noise.cuh (filter + sums EE:1565-1631,
//   0.338*N rule and float divide EE:1742-1797, Germ_Max EE:1229-1271); what changes is WHERE the sums are kept.
//
// Per (base, strand, C) the model needs sum over the KEPT records of float(depth) * float(C) (EE:1617), sum of depth and the
// record count, where "kept" is decided per base (both strand AFs <= 5 %).  noise_staged_kernel holds all of them per base:
// 54 registers of state for one C, and 128 registers (3 CTAs/SM, 0.63 of the HBM roof) for the round-1 sweep kernel with
// four.  But the four keep flags of a record form a 4-bit PATTERN that is the same for almost every record of a slot --
// 0111 when the slot's reference base is A: the reference allele is far above 5 %, the three error alleles far below --
// and every sum is exact in fp64 (integers, or fp32 products of 24 significant bits), so it may be split
// freely.  The thread therefore keeps ONE set of sums for "records whose pattern equals the slot's pattern" -- 2 depth sums,
// 2 * NC product sums, one count -- and a record with another pattern (a germline carrier, a noisy low-depth row: a few per
// ten thousand) adds to per-base sums in LOCAL memory, off the hot path.  At the end
//     sum[base] = (base in pattern ? pattern sums : 0) + spill[base].
// State in registers: 29 words for NC = 1 (54 before), 45 for NC = 5; the per-record work that depends on C is one FMUL,
// one conversion and one DADD per (strand, C).  If the first record of a slot is the odd one, the pattern is re-elected as
// soon as the irregular records outnumber the regular ones (a matter of speed only: the sums are exact either way).
// Alt-read sums and Germ_Max stay per base (they differ per base by nature).  Twin pairs inside the CTA tile are merged
// after the sample loop through the idle ring memory, base by base.
#include "as_kernels.h"

#include "as_device.cuh"
#include "as_noise.cuh"
#include "as_pipeline.cuh"

namespace asdev {

#define AS_PAT_UNSET 0xFFFFFFFFu

template <int NC>
struct SweepCs {
    float c[NC];
};

template <int NC>
struct PatCore {  // what the rare path reads and rewrites: passed to it and returned BY VALUE, so that it stays in registers
    uint32_t pat;        // the slot's keep pattern (bit i = base i kept), AS_PAT_UNSET before the first covered record
    uint32_t cnt0, irr;  // records with that pattern / with another non-empty pattern since the pattern was elected
    uint32_t spilled;    // the local-memory sums are in use (zeroed on first use)
    double sd_fw, sd_bw;         // strand depth over the pattern's records
    double t_fw[NC], t_bw[NC];   // float(depth) * float(C) over the pattern's records
};
template <int NC>
struct PatState : PatCore<NC> {
    uint32_t nrec, big;
    uint32_t sb_fw[4], sb_bw[4]; // alt reads over kept records, per base (32-bit partial sums, folded every AS_FOLD_EVERY)
    uint32_t g_x[4], g_rd[4];    // Germ_Max, as in FastBase
};

template <int NC>
struct PatSpill {  // per base; local memory
    double sd_fw[4], sd_bw[4], sb_fw[4], sb_bw[4];
    double p_fw[4][NC], p_bw[4][NC];
    uint32_t count[4];
};

template <int NC>
struct PatSums {
    double sd_fw, sd_bw, t_fw[NC], t_bw[NC];
    uint32_t cnt;
};

template <int NC>
__device__ __noinline__ void spill_zero(PatSpill<NC>* sp) {
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        sp->sd_fw[b] = sp->sd_bw[b] = sp->sb_fw[b] = sp->sb_bw[b] = 0.0;
        sp->count[b] = 0u;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) sp->p_fw[b][c] = sp->p_bw[b][c] = 0.0;
    }
}

// add v to every base of `mask`
template <int NC>
__device__ __noinline__ void spill_add(PatSpill<NC>* sp, uint32_t mask, PatSums<NC> v) {
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        if (!((mask >> b) & 1u)) continue;
        sp->sd_fw[b] = __dadd_rn(sp->sd_fw[b], v.sd_fw);
        sp->sd_bw[b] = __dadd_rn(sp->sd_bw[b], v.sd_bw);
        sp->count[b] += v.cnt;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            sp->p_fw[b][c] = __dadd_rn(sp->p_fw[b][c], v.t_fw[c]);
            sp->p_bw[b][c] = __dadd_rn(sp->p_bw[b][c], v.t_bw[c]);
        }
    }
}

template <int NC>
__device__ __forceinline__ void pat_init(PatState<NC>& s) {
    s.nrec = s.big = 0u;
    s.pat = AS_PAT_UNSET;
    s.cnt0 = s.irr = s.spilled = 0u;
    s.sd_fw = s.sd_bw = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) s.t_fw[c] = s.t_bw[c] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s.sb_fw[i] = s.sb_bw[i] = 0u;
        s.g_x[i] = 1u;
        s.g_rd[i] = 0u;
    }
}

// 32-bit alt-read sums -> local memory (a base's kept alt reads are <= 0.05 * 2^24 per record: 4096 records fit 32 bits)
template <int NC>
__device__ __forceinline__ void pat_fold(PatState<NC>& s, PatSpill<NC>* sp) {
    if (!s.spilled) { spill_zero<NC>(sp); s.spilled = 1u; }
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const uint32_t f = i == 0 ? s.sb_fw[0] : i == 1 ? s.sb_fw[1] : i == 2 ? s.sb_fw[2] : s.sb_fw[3];
        const uint32_t b = i == 0 ? s.sb_bw[0] : i == 1 ? s.sb_bw[1] : i == 2 ? s.sb_bw[2] : s.sb_bw[3];
        sp->sb_fw[i] = __dadd_rn(sp->sb_fw[i], u32_to_double(f));
        sp->sb_bw[i] = __dadd_rn(sp->sb_bw[i], u32_to_double(b));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s.sb_fw[i] = s.sb_bw[i] = 0u;
}

// What one record contributes to the sums that depend on C (and the depth sums), as exact doubles.
template <int NC>
__device__ __forceinline__ void pat_products(PatSums<NC>& v, uint32_t FW, uint32_t BW, const SweepCs<NC>& cs) {
    v.sd_fw = u32_to_double(FW);
    v.sd_bw = u32_to_double(BW);
    const float Ff = __uint2float_rn(FW), Bf = __uint2float_rn(BW);
#pragma unroll
    for (int c = 0; c < NC; ++c) {  // EE:1617: fp32 product, then widened
        v.t_fw[c] = (double)__fmul_rn(Ff, cs.c[c]);
        v.t_bw[c] = (double)__fmul_rn(Bf, cs.c[c]);
    }
    v.cnt = 1u;
}

// One record, branch-free (the K records of a stage interleave).  Returns true when the record carries a non-empty keep
// pattern other than the slot's: its per-C sums are then NOT added here; the caller hands it to pat_irregular.
#define AS_LIM_NEVER ((int32_t)0x80000000) /* no count (< 2^31, or the -1 / -2 of an absent row) is <= this */
template <int NC>
__device__ __forceinline__ bool pat_accumulate(PatState<NC>& s, const uint4 fw, const uint4 bw, const SweepCs<NC>& cs,
                                               const uint32_t cut) {
    const bool present = (int32_t)fw.x >= 0;  // AS_ABSENT otherwise
    s.nrec += present ? 1u : 0u;
    const uint32_t FW = fw.x + fw.y + fw.z + fw.w;  // EE:1155-1176
    const uint32_t BW = bw.x + bw.y + bw.z + bw.w;
    const uint32_t RD = FW + BW;
    s.big |= present ? RD : 0u;
    const bool cov = present & (min(FW, BW) >= cut);  // EE:1615, EE:1251
    const int32_t lim_fw = cov ? (int32_t)af_limit(FW) : AS_LIM_NEVER;
    const int32_t lim_bw = cov ? (int32_t)af_limit(BW) : AS_LIM_NEVER;
    const int32_t lim_rd = cov ? (int32_t)af_limit(RD) : AS_LIM_NEVER;
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t bf = comp(fw, i), bb = comp(bw, i);
        const bool keep = ((int32_t)bf <= lim_fw) & ((int32_t)bb <= lim_bw);  // EE:1613-1615
        mask |= keep ? (1u << i) : 0u;
        if (keep) { s.sb_fw[i] += bf; s.sb_bw[i] += bb; }
        // Germ_Max (EE:1251-1271), as fast_base_update
        const uint32_t x = bf + bb;
        const bool qual = (int32_t)x <= lim_rd;
        const bool ge = (unsigned long long)x * s.g_rd[i] >= (unsigned long long)s.g_x[i] * RD;  // EE:1263: value <= AF
        const bool first = qual & (s.g_rd[i] == 0u);
        const bool upd = qual & ge;
        s.g_x[i] = first ? 0u : (upd ? x : s.g_x[i]);
        s.g_rd[i] = first ? 1u : (upd ? RD : s.g_rd[i]);
    }
    PatSums<NC> v;
    pat_products<NC>(v, FW, BW, cs);
    s.pat = (s.pat == AS_PAT_UNSET && mask != 0u) ? mask : s.pat;  // the first covered record elects the pattern
    const bool regular = mask == s.pat;  // false for mask == 0 once a pattern is elected; (0 == UNSET never)
    if (regular) {  // predicated adds
        s.cnt0 += 1u;
        s.sd_fw = __dadd_rn(s.sd_fw, v.sd_fw);
        s.sd_bw = __dadd_rn(s.sd_bw, v.sd_bw);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            s.t_fw[c] = __dadd_rn(s.t_fw[c], v.t_fw[c]);
            s.t_bw[c] = __dadd_rn(s.t_bw[c], v.t_bw[c]);
        }
    }
    return (mask != 0u) & !regular;
}

// The rare record with another pattern (re-read from the resident stage): per-base sums in local memory.
template <int NC>
__device__ __noinline__ PatCore<NC> pat_irregular(PatCore<NC> s, PatSpill<NC>* sp, const uint4 fw, const uint4 bw, SweepCs<NC> cs) {
    const uint32_t FW = fw.x + fw.y + fw.z + fw.w, BW = bw.x + bw.y + bw.z + bw.w;
    const int32_t lim_fw = (int32_t)af_limit(FW), lim_bw = (int32_t)af_limit(BW);  // the record is covered (mask != 0)
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        mask |= (((int32_t)comp(fw, i) <= lim_fw) & ((int32_t)comp(bw, i) <= lim_bw)) ? (1u << i) : 0u;
    if (!s.spilled) { spill_zero<NC>(sp); s.spilled = 1u; }
    PatSums<NC> v;
    pat_products<NC>(v, FW, BW, cs);
    if (mask == s.pat) {  // the pattern was re-elected by an earlier record of the same stage: regular after all
        s.cnt0 += 1;
        s.sd_fw = __dadd_rn(s.sd_fw, v.sd_fw); s.sd_bw = __dadd_rn(s.sd_bw, v.sd_bw);
#pragma unroll
        for (int c = 0; c < NC; ++c) { s.t_fw[c] = __dadd_rn(s.t_fw[c], v.t_fw[c]); s.t_bw[c] = __dadd_rn(s.t_bw[c], v.t_bw[c]); }
    } else if (s.cnt0 <= s.irr) {
        // the elected pattern is not the majority: its sums move to the spill and this record's pattern takes over
        PatSums<NC> old;
        old.sd_fw = s.sd_fw; old.sd_bw = s.sd_bw; old.cnt = s.cnt0;
#pragma unroll
        for (int c = 0; c < NC; ++c) { old.t_fw[c] = s.t_fw[c]; old.t_bw[c] = s.t_bw[c]; }
        spill_add<NC>(sp, s.pat, old);
        s.pat = mask;
        s.cnt0 = 1u; s.irr = 0u;
        s.sd_fw = v.sd_fw; s.sd_bw = v.sd_bw;
#pragma unroll
        for (int c = 0; c < NC; ++c) { s.t_fw[c] = v.t_fw[c]; s.t_bw[c] = v.t_bw[c]; }
    } else {
        s.irr += 1;
        spill_add<NC>(sp, mask, v);
    }
    return s;
}

// totals of one base: pattern sums (when the base is in the pattern) + spill
template <int NC>
struct BaseTotals {
    double sd_fw, sd_bw, sb_fw, sb_bw, p_fw[NC], p_bw[NC];
    uint32_t count;
};

template <int NC>
__device__ __forceinline__ void pat_base_totals(const PatState<NC>& s, const PatSpill<NC>* sp, int i, BaseTotals<NC>& o) {
    const bool in_pat = s.pat != AS_PAT_UNSET && ((s.pat >> i) & 1u);
    o.sd_fw = in_pat ? s.sd_fw : 0.0;
    o.sd_bw = in_pat ? s.sd_bw : 0.0;
    o.count = in_pat ? s.cnt0 : 0u;
    o.sb_fw = u32_to_double(s.sb_fw[i]);
    o.sb_bw = u32_to_double(s.sb_bw[i]);
#pragma unroll
    for (int c = 0; c < NC; ++c) { o.p_fw[c] = in_pat ? s.t_fw[c] : 0.0; o.p_bw[c] = in_pat ? s.t_bw[c] : 0.0; }
    if (s.spilled) {
        o.sd_fw = __dadd_rn(o.sd_fw, sp->sd_fw[i]); o.sd_bw = __dadd_rn(o.sd_bw, sp->sd_bw[i]);
        o.sb_fw = __dadd_rn(o.sb_fw, sp->sb_fw[i]); o.sb_bw = __dadd_rn(o.sb_bw, sp->sb_bw[i]);
        o.count += sp->count[i];
#pragma unroll
        for (int c = 0; c < NC; ++c) { o.p_fw[c] = __dadd_rn(o.p_fw[c], sp->p_fw[i][c]); o.p_bw[c] = __dadd_rn(o.p_bw[c], sp->p_bw[i][c]); }
    }
}

// what the second slot of an in-tile twin pair hands to the first, in rounds through the idle ring memory
struct PairGermXfer {
    uint32_t g_x[4], g_rd[4];
    PairFirst first;
    uint32_t nrec, big;
};
template <int NC>
struct PairBaseXfer {
    BaseTotals<NC> t;
};

__device__ __forceinline__ void named_sync() { asm volatile("bar.sync 1, %0;" ::"n"(AS_TILE_SLOTS) : "memory"); }
__device__ __forceinline__ bool named_any(bool pred) {  // OR over the 128 consumer threads (barrier 1)
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "bar.red.or.pred q, 1, %2, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)pred), "n"(AS_TILE_SLOTS)
        : "memory");
    return r != 0;
}

// role of a slot inside its CTA tile: 0 singleton, 1 first slot of a twin pair whose second slot is in this tile, 2 that
// second slot, 3 any other member of a twin group (left to noise_pair_kernel / noise_twin_kernel)
__device__ __forceinline__ int intile_distance(const int32_t* __restrict__ twin_next, const int32_t* __restrict__ twin_head,
                                               int64_t idx, int64_t gid, int tid, int n_slots) {
    const int32_t nx = twin_next[idx];
    if (nx < 0 || twin_head[idx] != (int32_t)gid) return 0;
    const int64_t d = (int64_t)nx - gid;
    if (d <= 0 || tid + d >= n_slots) return 0;
    if (twin_next[idx + d] >= 0) return 0;  // three or more enumerations: general kernel
    return (int)d;
}

// INTERLEAVE: the K records of a stage are unrolled (the compiler interleaves them: more instruction-level parallelism,
// more registers) or walked one by one (the four bases of a record are four independent chains already; fewer registers,
// more resident CTAs).
template <int NC, int K, int STAGES, int MINB, bool INTERLEAVE>
__global__ void __launch_bounds__(AS_CTA_THREADS, MINB)
noise_pattern_kernel(const uint4* __restrict__ counts, int S, int64_t P, int64_t p0, int64_t p1,
                     const int32_t* __restrict__ twin_next, const int32_t* __restrict__ twin_head, int64_t twin_base,
                     SweepCs<NC> cs, int n_c, uint32_t cut, float* __restrict__ thr, int64_t thr_stride,
                     float* __restrict__ germ_val, uint8_t* __restrict__ germ_state, uint32_t* __restrict__ count,
                     uint32_t* __restrict__ nrec) {
    static_assert(StageRing<K, STAGES>::kStageBytes * STAGES >= AS_TILE_SLOTS * (int)sizeof(PairBaseXfer<NC>), "hand-over block");
    static_assert(StageRing<K, STAGES>::kStageBytes * STAGES >= AS_TILE_SLOTS * (int)sizeof(PairGermXfer), "hand-over block");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES];
    StageRing<K, STAGES> ring;
    ring.init(smem_raw, bars);
    const int64_t tile0 = p0 + (int64_t)blockIdx.x * AS_TILE_SLOTS;
    const int n_slots = (int)min((int64_t)AS_TILE_SLOTS, p1 - tile0);
    const int tid = threadIdx.x;
    if (tid >= AS_TILE_SLOTS) {  // producer warp
        if (tid == AS_TILE_SLOTS) ring.produce(counts + tile0, 2 * P, P, 0, S, n_slots);
        return;
    }
    const int64_t p = tile0 + tid;
    bool active = tid < n_slots;
    int role = 0, twin_d = 0;
    if (active && twin_next != nullptr) {
        const int64_t gid = p + twin_base;
        if (twin_next[p] >= 0 || twin_head[p] != (int32_t)gid) role = 3;
        twin_d = intile_distance(twin_next, twin_head, p, gid, tid, n_slots);
        if (twin_d > 0) {
            role = 1;
        } else if (role == 3) {
            const int64_t back = gid - (int64_t)twin_head[p];
            if (back > 0 && back <= tid && intile_distance(twin_next, twin_head, p - back, gid - back, tid - (int)back, n_slots) == (int)back)
                role = 2;
        }
        if (role == 3) active = false;
    }

    PatState<NC> s;
    PatSpill<NC> spill;  // local memory: only irregular records and the 4096-sample fold touch it
    pat_init(s);
    int it = 0, since_fold = 0;
    for (int t = 0; t < S; t += K, ++it) {
        const uint4* st = ring.consumer_wait(it);
        const int k = min(K, S - t);
        if (active) {
            uint32_t irregular = 0;
            if (INTERLEAVE) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (j < k) {
                        const uint4 fw = st[(j * 2 + 0) * AS_TILE_SLOTS + tid];
                        const uint4 bw = st[(j * 2 + 1) * AS_TILE_SLOTS + tid];
                        irregular |= pat_accumulate<NC>(s, fw, bw, cs, cut) ? (1u << j) : 0u;
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < k; ++j) {
                    const uint4 fw = st[(j * 2 + 0) * AS_TILE_SLOTS + tid];
                    const uint4 bw = st[(j * 2 + 1) * AS_TILE_SLOTS + tid];
                    irregular |= pat_accumulate<NC>(s, fw, bw, cs, cut) ? (1u << j) : 0u;
                }
            }
            while (irregular) {  // a few per ten thousand records
                const int j = __ffs(irregular) - 1;
                irregular &= irregular - 1;
                static_cast<PatCore<NC>&>(s) = pat_irregular<NC>(static_cast<const PatCore<NC>&>(s), &spill, st[(j * 2 + 0) * AS_TILE_SLOTS + tid],
                                                                  st[(j * 2 + 1) * AS_TILE_SLOTS + tid], cs);
            }
        }
        ring.consumer_release(it);
        since_fold += K;
        if (since_fold >= AS_FOLD_EVERY) { if (active) pat_fold<NC>(s, &spill); since_fold = 0; }
    }

    // ---- twin pairs inside the tile: the second slot's thread hands its state to the first slot's (pair_merge semantics)
    bool pairs = false;
    if (twin_next != nullptr) pairs = named_any(role == 1 || role == 2);  // also: every consumer warp is done with the ring
    if (pairs) {
        PairGermXfer* gx = reinterpret_cast<PairGermXfer*>(smem_raw);
        PairFirst mine;
        if (role == 1 || role == 2) {
            FastAcc f;  // pair_find_first only looks at g_rd
#pragma unroll
            for (int i = 0; i < 4; ++i) f.b[i].g_rd = s.g_rd[i];
            pair_find_first(mine, f, counts + p, S, P, cs.c[0], cut);
        }
        if (role == 2) {
            PairGermXfer& o = gx[tid];
#pragma unroll
            for (int i = 0; i < 4; ++i) { o.g_x[i] = s.g_x[i]; o.g_rd[i] = s.g_rd[i]; }
            o.first = mine; o.nrec = s.nrec; o.big = s.big;
        }
        named_sync();
        if (role == 1) {
            const PairGermXfer& o = gx[tid + twin_d];
            s.nrec += o.nrec;
            s.big |= o.big;
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // Germ_Max part of pair_merge (as_noise.cuh)
                const int ca = s.g_rd[i] == 0u ? 0 : (s.g_rd[i] == 1u ? 1 : 2), cb = o.g_rd[i] == 0u ? 0 : (o.g_rd[i] == 1u ? 1 : 2);
                if (ca == 0 && cb == 0) continue;
                const bool a_is_first = ca > 0 && (cb == 0 || mine.s[i] <= o.first.s[i]);
                uint32_t bx = 0, brd = 1;
                bool have = false;
                if (ca == 2) rational_max(bx, brd, have, s.g_x[i], s.g_rd[i]);
                if (cb == 2) rational_max(bx, brd, have, o.g_x[i], o.g_rd[i]);
                if (ca > 0 && cb > 0) {
                    if (a_is_first) rational_max(bx, brd, have, o.first.x[i], o.first.rd[i]);
                    else rational_max(bx, brd, have, mine.x[i], mine.rd[i]);
                }
                s.g_x[i] = have ? bx : 0u;
                s.g_rd[i] = have ? brd : 1u;
            }
        }
        named_sync();
    }
    const int tw = role == 1 ? twin_d : 0;
    const bool store = active && role != 2;
    const bool redo = s.big >= (1u << 24);  // a depth of 2^24 or more: int -> float is inexact, general code below

    uint32_t cnt_out[4];
    float g_out[4];
    uint32_t gs_out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        BaseTotals<NC> tot;
        pat_base_totals<NC>(s, &spill, i, tot);
        if (pairs) {
            PairBaseXfer<NC>* bx = reinterpret_cast<PairBaseXfer<NC>*>(smem_raw);
            if (role == 2) bx[tid].t = tot;
            named_sync();
            if (role == 1) {
                const BaseTotals<NC>& o = bx[tid + twin_d].t;
                tot.sd_fw = __dadd_rn(tot.sd_fw, o.sd_fw); tot.sd_bw = __dadd_rn(tot.sd_bw, o.sd_bw);
                tot.sb_fw = __dadd_rn(tot.sb_fw, o.sb_fw); tot.sb_bw = __dadd_rn(tot.sb_bw, o.sb_bw);
                tot.count += o.count;
#pragma unroll
                for (int c = 0; c < NC; ++c) { tot.p_fw[c] = __dadd_rn(tot.p_fw[c], o.p_fw[c]); tot.p_bw[c] = __dadd_rn(tot.p_bw[c], o.p_bw[c]); }
            }
            named_sync();
        }
        cnt_out[i] = tot.count;
        if (store && !redo) {
            NoiseBase nb;
            nb.s_b_fw = nb.s_b_bw = 0ull;  // folded into s_p below
            nb.s_d_fw = (unsigned long long)__double2ll_rn(tot.sd_fw);
            nb.s_d_bw = (unsigned long long)__double2ll_rn(tot.sd_bw);
            nb.count = tot.count;
            nb.g_n = s.g_rd[i] == 0u ? 0u : (s.g_rd[i] == 1u ? 1u : 2u);
            nb.g_x = s.g_x[i]; nb.g_rd = s.g_rd[i] == 0u ? 1u : s.g_rd[i];
            nb.g_first_x = 0; nb.g_first_rd = 1;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c < n_c) {
                    nb.s_p_fw = __dadd_rn(tot.sb_fw, tot.p_fw[c]);
                    nb.s_p_bw = __dadd_rn(tot.sb_bw, tot.p_bw[c]);
                    float q_fw, q_bw, g;
                    uint32_t st;
                    noise_final_base(nb, s.nrec, i, q_fw, q_bw, g, st);
                    *reinterpret_cast<float2*>(thr + c * thr_stride + p * 8 + 2 * i) = make_float2(q_fw, q_bw);
                    if (tw > 0) *reinterpret_cast<float2*>(thr + c * thr_stride + (p + tw) * 8 + 2 * i) = make_float2(q_fw, q_bw);
                    if (c == 0) { g_out[i] = g; gs_out |= st << (8 * i); }
                }
            }
        }
    }
    if (!store) return;
    if (!redo) {
#pragma unroll 1
        for (int r = 0; r <= (tw > 0 ? 1 : 0); ++r) {
            const int64_t q = p + r * tw;
            *reinterpret_cast<float4*>(germ_val + q * 4) = make_float4(g_out[0], g_out[1], g_out[2], g_out[3]);
            *reinterpret_cast<uint32_t*>(germ_state + q * 4) = gs_out;
            *reinterpret_cast<uint4*>(count + q * 4) = make_uint4(cnt_out[0], cnt_out[1], cnt_out[2], cnt_out[3]);
            nrec[q] = s.nrec;
        }
        return;
    }
    // general code, value by value (never seen in practice; keeps the result exact for every uint32 input)
    const uint4* q = counts + p;
    for (int c = 0; c < n_c; ++c) {
        float cv = cs.c[0];
#pragma unroll
        for (int cc = 1; cc < NC; ++cc) cv = c == cc ? cs.c[cc] : cv;
        NoiseAcc acc;
        noise_init(acc);
#pragma unroll 1
        for (int smp = 0; smp < S; ++smp) {
#pragma unroll 1
            for (int r = 0; r <= (tw > 0 ? 1 : 0); ++r) {
                const uint4 fw = ld_stream(q + (int64_t)smp * 2 * P + r * tw);
                const uint4 bw = ld_stream(q + (int64_t)smp * 2 * P + P + r * tw);
                noise_accumulate<false>(acc, fw, bw, cv, cut);
            }
        }
        float t[8], g[4];
        uint32_t cn[4], gs = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t st;
            noise_final_base(acc.b[i], acc.nrec, i, t[2 * i], t[2 * i + 1], g[i], st);
            cn[i] = acc.b[i].count;
            gs |= st << (8 * i);
        }
        for (int r = 0; r <= (tw > 0 ? 1 : 0); ++r) {
            const int64_t slot = p + r * tw;
            float4* t4 = reinterpret_cast<float4*>(thr + c * thr_stride + slot * 8);
            t4[0] = make_float4(t[0], t[1], t[2], t[3]);
            t4[1] = make_float4(t[4], t[5], t[6], t[7]);
            if (c == 0) {
                *reinterpret_cast<float4*>(germ_val + slot * 4) = make_float4(g[0], g[1], g[2], g[3]);
                *reinterpret_cast<uint32_t*>(germ_state + slot * 4) = gs;
                *reinterpret_cast<uint4*>(count + slot * 4) = make_uint4(cn[0], cn[1], cn[2], cn[3]);
                nrec[slot] = acc.nrec;
            }
        }
    }
}

}  // namespace asdev

using namespace asdev;

template <int NC, int K, int STAGES, int MINB, bool INTERLEAVE = true>
static cudaError_t launch_pattern(const uint32_t* d_counts, int S, int64_t P, int64_t p0, int64_t p1, const int32_t* tn,
                                  const int32_t* th, int64_t twin_base, const float* c_values, int n_c, uint32_t cut, float* thr,
                                  int64_t thr_stride, float* gv, uint8_t* gs, uint32_t* cnt, uint32_t* nrec, cudaStream_t st) {
    static bool configured[AS_MAX_DEVICES] = {};
    const int smem = StageRing<K, STAGES>::kStageBytes * STAGES;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= AS_MAX_DEVICES || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(noise_pattern_kernel<NC, K, STAGES, MINB, INTERLEAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < AS_MAX_DEVICES) configured[dev] = true;
    }
    SweepCs<NC> cs;
    for (int c = 0; c < NC; ++c) cs.c[c] = c_values[c < n_c ? c : n_c - 1];
    noise_pattern_kernel<NC, K, STAGES, MINB, INTERLEAVE><<<(unsigned)((p1 - p0 + AS_TILE_SLOTS - 1) / AS_TILE_SLOTS), AS_CTA_THREADS, smem, st>>>(
        reinterpret_cast<const uint4*>(d_counts), S, P, p0, p1, tn, th, twin_base, cs, n_c, cut, thr, thr_stride, gv, gs, cnt, nrec);
    return cudaGetLastError();
}

// every output of the noise model for n_c (1..8) values of C in one pass: table c at d_thr + c * thr_stride floats.
// Slots of twin groups that are not a pair inside one CTA tile are left to as_launch_noise_twin_groups (per value).
// geom: ring geometry / CTAs per SM, chosen by measurement (profiles/r02_noise_pattern_geometries.log).
cudaError_t as_launch_noise_pattern(int geom, const uint32_t* d_counts, int S, int64_t P, int64_t p0, int64_t p1,
                                    const int32_t* d_twin_next, const int32_t* d_twin_head, int64_t twin_base,
                                    const float* c_values, int n_c, uint32_t cut, float* d_thr, int64_t thr_stride,
                                    float* d_germ_val, uint8_t* d_germ_state, uint32_t* d_count, uint32_t* d_nrec, cudaStream_t st) {
    if (p1 <= p0 || n_c <= 0) return cudaSuccess;
#define AS_PAT_ARGS d_counts, S, P, p0, p1, d_twin_next, d_twin_head, twin_base, c_values, n_c, cut, d_thr, thr_stride, d_germ_val, d_germ_state, d_count, d_nrec, st
    // CTAs per SM follow the register need of the per-C sums (ptxas: profiles/r02_ptxas_registers.txt); no variant spills in
    // its sample loop.  Measured on the c3 shard (profiles/r02_noise_pattern_geometries.log): NC = 1 is slower than
    // noise_staged_kernel (1.35 vs 1.10 ms: the ALU pipe, not the register file, bounds one value) and is kept as a
    // cross-check (noise variants 7, 8); NC = 5 at 3 CTAs/SM 1.90 ms against 2.93 ms at 4 CTAs/SM with spills.
    if (n_c == 1) {
        if (geom == 1) return launch_pattern<1, 4, 3, 4, false>(AS_PAT_ARGS);
        return launch_pattern<1, 4, 3, 4>(AS_PAT_ARGS);
    }
    if (n_c <= 3) return launch_pattern<3, 4, 3, 4>(AS_PAT_ARGS);
    if (n_c <= 5) {
        if (geom == 1) return launch_pattern<5, 4, 3, 3, false>(AS_PAT_ARGS);
        return launch_pattern<5, 4, 3, 3>(AS_PAT_ARGS);
    }
    if (n_c <= 8) return launch_pattern<8, 4, 3, 2>(AS_PAT_ARGS);
#undef AS_PAT_ARGS
    return cudaErrorInvalidValue;
}
