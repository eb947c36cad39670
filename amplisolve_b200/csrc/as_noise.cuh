// Noise-model arithmetic (device side) -- the per-record filter/accumulate and the per-position
// finalise of AmpliSolve's estimateThresholds and Germ_Max, restated for dense count tensors.
//
//   EE = source_codes/AmpliSolveErrorEstimation.cpp of the reference.
//   record filter + sums      EE:1565-1631 (A) == 1816-1878 (C) == 2062-2125 (G) == 2309-2372 (T)
//   0.338*N rule, float divide EE:1742-1797
//   Germ_Max                  EE:1229-1232, EE:1251-1271 (A, floor -888), EE:1309-1329 (C/G/T, floor 0)
//
// Division-free exact forms (each is an identity, not an approximation -- DESIGN.md "noise kernel"):
//   (double)(float(b)/float(D)) <= 0.05   <=>   b' <= (26843545 * D') >> 29
//       0.05 lies between the floats 0x3D4CCCCC and 0x3D4CCCCD; their midpoint is 26843545/2^29 and the
//       tie rounds to the even lower float, so RN(b'/D') <= 0x3D4CCCCC <=> b'/D' <= 26843545/2^29.
//       b', D' are the operands after int->float conversion: equal to b, D below 2^24, and rounded to
//       24 significant bits above (rare path, handled by round24()).
//   running max of float(X)/float(RD): IEEE rounding is monotone, so the max quotient is the quotient
//       of the max rational; rationals are compared by 64-bit cross multiplication and divided once.
//   sum_nt = sum(b) + sum(float(D)*C): every partial sum of the reference is exact in fp64 for
//       cut >= ~10 and C >= ~1e-4, so integer sums of b and D plus one fp64 sum of the
//       fp32 products, added once at the end, give the same double in any order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "as_device.cuh"

namespace asdev {

#define AS_AF_LIMIT_NUM 26843545ull /* 0.05 as the float-rounding boundary, times 2^29 */

__device__ __forceinline__ uint32_t round24(uint32_t v) { /* the integer value of float(v) */
    return __float2uint_rn(__uint2float_rn(v));
}
__device__ __forceinline__ uint32_t af_limit(uint32_t depth_as_float) {
    return (uint32_t)((AS_AF_LIMIT_NUM * (unsigned long long)depth_as_float) >> 29);
}

// the same for depths below 2^29 (fast paths: a depth of 2^24 or more raises `big` and the slot is redone anyway): the high
// word of (8 * D) * 26843545, one multiply on the FMA pipe instead of a wide multiply and a 64-bit shift on the ALU pipe
__device__ __forceinline__ uint32_t af_limit_fast(uint32_t depth) { return __umulhi(depth * 8u, (uint32_t)AS_AF_LIMIT_NUM); }

struct NoiseBase {
    unsigned long long s_b_fw, s_b_bw;  // sum of alt reads over kept records (EE:1617, EE:1619)
    unsigned long long s_d_fw, s_d_bw;  // sum of strand depth over kept records (EE:1618, EE:1620)
    double s_p_fw, s_p_bw;              // sum of float(depth)*float(C) over kept records
    uint32_t count;                     // EE:1626
    // Germ_Max: qualifying records seen, the first one (needed only when segments are merged) and the
    // best rational among the others
    uint32_t g_n, g_first_x, g_first_rd, g_x, g_rd;
};

struct NoiseAcc {
    NoiseBase b[4];
    uint32_t nrec;  // records of the position = Value_Hash.count() (EE:1742)
};

__device__ __forceinline__ void noise_init(NoiseAcc& a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a.b[i].s_b_fw = a.b[i].s_b_bw = a.b[i].s_d_fw = a.b[i].s_d_bw = 0ull;
        a.b[i].s_p_fw = a.b[i].s_p_bw = 0.0;
        a.b[i].count = 0;
        a.b[i].g_n = 0;
        a.b[i].g_first_x = 0; a.b[i].g_first_rd = 1;
        a.b[i].g_x = 0; a.b[i].g_rd = 1;
    }
    a.nrec = 0;
}

// One (sample, slot) record.  TRACK_FIRST keeps the first qualifying Germ_Max record so that partial
// states of consecutive sample segments can be merged (noise_merge); the single-pass kernel does not
// need it because the reference drops that record anyway (EE:1258-1262).
template <bool TRACK_FIRST>
__device__ __forceinline__ void noise_accumulate(NoiseAcc& a, const uint4 fw, const uint4 bw, const float C,
                                                 const uint32_t cut) {
    if ((int32_t)fw.x < 0) return;  // AS_ABSENT: no ASEQ row for this (sample, slot)
    a.nrec += 1;
    const uint32_t FW = fw.x + fw.y + fw.z + fw.w;  // EE:1155-1176
    const uint32_t BW = bw.x + bw.y + bw.z + bw.w;
    const uint32_t RD = FW + BW;
    if (!(FW >= cut && BW >= cut)) return;  // EE:1615 (thresholds) and EE:1251 (Germ_Max) both need it

    // operands as the reference's float conversions see them
    uint4 tfw = fw, tbw = bw;
    uint4 tx = make_uint4(fw.x + bw.x, fw.y + bw.y, fw.z + bw.z, fw.w + bw.w);  // totals A,C,G,T (EE:1229-1232)
    uint32_t tFW = FW, tBW = BW, tRD = RD;
    if (RD >= (1u << 24)) {  // > 16.7 M reads on one position: int->float is no longer exact
        tfw = make_uint4(round24(fw.x), round24(fw.y), round24(fw.z), round24(fw.w));
        tbw = make_uint4(round24(bw.x), round24(bw.y), round24(bw.z), round24(bw.w));
        tx = make_uint4(round24(tx.x), round24(tx.y), round24(tx.z), round24(tx.w));
        tFW = round24(FW); tBW = round24(BW); tRD = round24(RD);
    }
    const uint32_t lim_fw = af_limit(tFW), lim_bw = af_limit(tBW), lim_rd = af_limit(tRD);
    const double p_fw = (double)__fmul_rn(__uint2float_rn(FW), C);  // EE:1617: fp32 product, then widened
    const double p_bw = (double)__fmul_rn(__uint2float_rn(BW), C);

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        NoiseBase& s = a.b[i];
        const uint32_t bf = comp(fw, i), bb = comp(bw, i);
        if (comp(tfw, i) <= lim_fw && comp(tbw, i) <= lim_bw) {  // EE:1613-1615
            s.s_b_fw += bf; s.s_b_bw += bb;
            s.s_d_fw += FW; s.s_d_bw += BW;
            s.s_p_fw = __dadd_rn(s.s_p_fw, p_fw);
            s.s_p_bw = __dadd_rn(s.s_p_bw, p_bw);
            s.count += 1;
        }
        const uint32_t x = comp(tx, i);
        if (x <= lim_rd) {  // EE:1251: (double)(float(X)/float(RD)) <= 0.05
            if (s.g_n == 0) {
                if (TRACK_FIRST) { s.g_first_x = x; s.g_first_rd = tRD; }
            } else if ((unsigned long long)x * s.g_rd >= (unsigned long long)s.g_x * tRD) {  // EE:1263: value <= AF
                s.g_x = x; s.g_rd = tRD;
            }
            s.g_n += 1;
        }
    }
}

// ---- fast path of the single-pass kernel -----------------------------------------------------------
// Same arithmetic as noise_accumulate<false> for records whose depth is below 2^24 (int -> float exact),
// arranged for the issue budget: the depth sums ride the otherwise idle fp64 pipe (exact: integers below
// 2^53), the alt-read sums are 32-bit (b <= 0.05 * 2^24, folded into the fp64 sum every AS_FOLD_EVERY
// samples), the coverage gate is folded into the compare chains.  A record with depth >= 2^24 only
// raises `big`; the caller then redoes the slot with the general code.
#define AS_FOLD_EVERY 4096

struct FastBase {
    uint32_t s_b_fw, s_b_bw;  // 32-bit partial sums of alt reads
    double s_d_fw, s_d_bw;    // sum of strand depth (exact in fp64)
    double s_p_fw, s_p_bw;    // sum of float(depth)*float(C), plus the folded alt-read sums
    uint32_t count;
    // Germ_Max without a counter: g_rd == 0 no qualifying record yet (g_x = 1 makes every comparison fail),
    // g_rd == 1 exactly one (it is dropped, EE:1258-1262; real depths are >= 2 * cut), else the best rational
    uint32_t g_x, g_rd;
};
struct FastAcc {
    FastBase b[4];
    uint32_t nrec, big;
};

__device__ __forceinline__ void fast_init(FastAcc& a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a.b[i].s_b_fw = a.b[i].s_b_bw = 0u;
        a.b[i].s_d_fw = a.b[i].s_d_bw = a.b[i].s_p_fw = a.b[i].s_p_bw = 0.0;
        a.b[i].count = 0u;
        a.b[i].g_x = 1u;
        a.b[i].g_rd = 0u;
    }
    a.nrec = 0; a.big = 0;
}

__device__ __forceinline__ double u32_to_double(uint32_t v) {  // exact, one DADD instead of a conversion-unit op
    return __dsub_rn(__hiloint2double(0x43300000, (int)v), 4503599627370496.0);
}

__device__ __forceinline__ void fast_fold(FastAcc& a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a.b[i].s_p_fw = __dadd_rn(a.b[i].s_p_fw, u32_to_double(a.b[i].s_b_fw));
        a.b[i].s_p_bw = __dadd_rn(a.b[i].s_p_bw, u32_to_double(a.b[i].s_b_bw));
        a.b[i].s_b_fw = a.b[i].s_b_bw = 0u;
    }
}

// what one record contributes to every base of its slot
struct FastRecord {
    int32_t lim_fw, lim_bw, lim_rd;  // 5 % limits (af_limit), or -1 when the coverage gate fails: every signed test fails
    uint32_t RD;
    double d_fw, d_bw, p_fw, p_bw;   // strand depths and float(depth)*float(C), as exact doubles
};

__device__ __forceinline__ void fast_record(FastRecord& r, const uint4 fw, const uint4 bw, const float C,
                                            const uint32_t cut) {
    const uint32_t FW = fw.x + fw.y + fw.z + fw.w;
    const uint32_t BW = bw.x + bw.y + bw.z + bw.w;
    r.RD = FW + BW;
    const bool cov = min(FW, BW) >= cut;  // counts are < 2^31, so the signed compares below are safe
    r.lim_fw = cov ? (int32_t)af_limit_fast(FW) : -1;
    r.lim_bw = cov ? (int32_t)af_limit_fast(BW) : -1;
    r.lim_rd = cov ? (int32_t)af_limit_fast(r.RD) : -1;
    r.d_fw = u32_to_double(FW);
    r.d_bw = u32_to_double(BW);
    r.p_fw = (double)__fmul_rn(__uint2float_rn(FW), C);
    r.p_bw = (double)__fmul_rn(__uint2float_rn(BW), C);
}

__device__ __forceinline__ void fast_base_update(FastBase& s, const FastRecord& r, const uint32_t bf, const uint32_t bb) {
    if (((int32_t)bf <= r.lim_fw) & ((int32_t)bb <= r.lim_bw)) {
        s.s_b_fw += bf; s.s_b_bw += bb;
        s.s_d_fw = __dadd_rn(s.s_d_fw, r.d_fw); s.s_d_bw = __dadd_rn(s.s_d_bw, r.d_bw);
        s.s_p_fw = __dadd_rn(s.s_p_fw, r.p_fw); s.s_p_bw = __dadd_rn(s.s_p_bw, r.p_bw);
        s.count += 1;
    }
    const uint32_t x = bf + bb;
    const bool qual = (int32_t)x <= r.lim_rd;
    // bitwise &: both products are always formed so that the update is predicated, not branched
    // EE:1263 (value <= AF) as the sign of x * g_rd - g_x * RD: two wide multiply-adds and ONE compare of the high word
    // (all four numbers are below 2^24 where it counts, so the difference fits 64 signed bits with room to spare)
    int32_t diff_hi;
    asm("{\n"
        ".reg .u64 p, d;\n"
        ".reg .s32 ng, lo;\n"
        "mul.wide.u32 p, %1, %2;\n"
        "neg.s32 ng, %3;\n"
        "mad.wide.s32 d, ng, %4, p;\n"
        "mov.b64 {lo, %0}, d;\n"
        "}\n"
        : "=r"(diff_hi)
        : "r"(x), "r"(s.g_rd), "r"(s.g_x), "r"(r.RD));
    const bool ge = diff_hi >= 0;
    const bool first = qual & (s.g_rd == 0u);
    const bool upd = qual & ge;
    s.g_x = first ? 0u : (upd ? x : s.g_x);
    s.g_rd = first ? 1u : (upd ? r.RD : s.g_rd);
}

__device__ __forceinline__ void fast_accumulate(FastAcc& a, const uint4 fw, const uint4 bw, const float C,
                                                const uint32_t cut) {
    if ((int32_t)fw.x < 0) return;  // AS_ABSENT
    a.nrec += 1;
    FastRecord r;
    fast_record(r, fw, bw, C, cut);
    a.big |= r.RD;
#pragma unroll
    for (int i = 0; i < 4; ++i) fast_base_update(a.b[i], r, comp(fw, i), comp(bw, i));
}

__device__ __forceinline__ void fast_base_to_general(const FastBase& s, NoiseBase& d) {
    d.s_b_fw = s.s_b_fw; d.s_b_bw = s.s_b_bw;
    d.s_d_fw = (unsigned long long)__double2ll_rn(s.s_d_fw); d.s_d_bw = (unsigned long long)__double2ll_rn(s.s_d_bw);
    d.s_p_fw = s.s_p_fw; d.s_p_bw = s.s_p_bw;
    d.count = s.count;
    d.g_n = s.g_rd == 0u ? 0u : (s.g_rd == 1u ? 1u : 2u);
    d.g_x = s.g_x; d.g_rd = s.g_rd == 0u ? 1u : s.g_rd;
    d.g_first_x = 0; d.g_first_rd = 1;
}

__device__ __forceinline__ void fast_to_general(const FastAcc& f, NoiseAcc& a) {
    a.nrec = f.nrec;
#pragma unroll
    for (int i = 0; i < 4; ++i) fast_base_to_general(f.b[i], a.b[i]);
}

// ---- exact merge of the two slots of a twin pair (fast-path states) ----------------------------------------
// Each thread reduced the rows of its own slot as if it were a singleton, so each dropped ITS first qualifying
// Germ_Max record.  In file order the rows interleave -- (sample 0, slot a), (sample 0, slot b), (sample 1, slot a) ...
// -- and the reference drops only the overall first qualifying record (EE:1258-1262).  PairFirst recovers, per base,
// the sample index and the rational of a thread's own first qualifying record (a short re-scan from sample 0 that
// stops as soon as every base that has one is found).
struct PairFirst {
    uint32_t s[4], x[4], rd[4];
};
struct PairXfer {
    FastAcc f;
    PairFirst first;
};

__device__ __forceinline__ void pair_find_first(PairFirst& o, const FastAcc& f, const uint4* __restrict__ q, int S, int64_t P,
                                                float C, uint32_t cut) {
    uint32_t need = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.s[i] = 0xFFFFFFFFu; o.x[i] = 0; o.rd[i] = 1;
        if (f.b[i].g_rd != 0u) need |= 1u << i;
    }
#pragma unroll 1
    for (int s = 0; s < S && need; ++s) {
        uint4 fw, bw;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(fw.x), "=r"(fw.y), "=r"(fw.z), "=r"(fw.w) : "l"(q + (int64_t)s * 2 * P));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(bw.x), "=r"(bw.y), "=r"(bw.z), "=r"(bw.w) : "l"(q + (int64_t)s * 2 * P + P));
        if ((int32_t)fw.x < 0) continue;
        FastRecord r;
        fast_record(r, fw, bw, C, cut);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = comp(fw, i) + comp(bw, i);
            if ((need >> i & 1u) && (int32_t)x <= r.lim_rd) {
                o.s[i] = (uint32_t)s; o.x[i] = x; o.rd[i] = r.RD;
                need &= ~(1u << i);
            }
        }
    }
}

__device__ __forceinline__ void rational_max(uint32_t& bx, uint32_t& brd, bool& have, uint32_t x, uint32_t rd) {
    if (!have || (unsigned long long)x * brd >= (unsigned long long)bx * rd) { bx = x; brd = rd; }
    have = true;
}

// A: state of the first slot's thread (rows precede B's within a sample), B: the second slot's.  Result in A.
__device__ __forceinline__ void pair_merge(FastAcc& A, const PairFirst& fa, const FastAcc& B, const PairFirst& fb) {
    A.nrec += B.nrec;
    A.big |= B.big;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        FastBase& a = A.b[i];
        const FastBase& b = B.b[i];
        a.s_b_fw += b.s_b_fw; a.s_b_bw += b.s_b_bw;
        a.s_d_fw = __dadd_rn(a.s_d_fw, b.s_d_fw); a.s_d_bw = __dadd_rn(a.s_d_bw, b.s_d_bw);
        a.s_p_fw = __dadd_rn(a.s_p_fw, b.s_p_fw); a.s_p_bw = __dadd_rn(a.s_p_bw, b.s_p_bw);
        a.count += b.count;
        const int ca = a.g_rd == 0u ? 0 : (a.g_rd == 1u ? 1 : 2), cb = b.g_rd == 0u ? 0 : (b.g_rd == 1u ? 1 : 2);
        if (ca == 0 && cb == 0) continue;  // still absent: (g_x, g_rd) = (1, 0)
        const bool a_is_first = ca > 0 && (cb == 0 || fa.s[i] <= fb.s[i]);
        uint32_t bx = 0, brd = 1;
        bool have = false;
        if (ca == 2) rational_max(bx, brd, have, a.g_x, a.g_rd);
        if (cb == 2) rational_max(bx, brd, have, b.g_x, b.g_rd);
        if (ca > 0 && cb > 0) {
            if (a_is_first) rational_max(bx, brd, have, fb.x[i], fb.rd[i]);
            else rational_max(bx, brd, have, fa.x[i], fa.rd[i]);
        }
        a.g_x = have ? bx : 0u;    // (0, 1): exactly one qualifying record in the pair
        a.g_rd = have ? brd : 1u;
    }
}

// Merge the state R of a LATER record segment into L (earlier).  Associative.
__device__ __forceinline__ void noise_merge(NoiseAcc& L, const NoiseAcc& R) {
    L.nrec += R.nrec;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        NoiseBase& l = L.b[i];
        const NoiseBase& r = R.b[i];
        l.s_b_fw += r.s_b_fw; l.s_b_bw += r.s_b_bw;
        l.s_d_fw += r.s_d_fw; l.s_d_bw += r.s_d_bw;
        l.s_p_fw = __dadd_rn(l.s_p_fw, r.s_p_fw);
        l.s_p_bw = __dadd_rn(l.s_p_bw, r.s_p_bw);
        l.count += r.count;
        if (r.g_n == 0) continue;
        if (l.g_n == 0) {
            l.g_n = r.g_n; l.g_first_x = r.g_first_x; l.g_first_rd = r.g_first_rd; l.g_x = r.g_x; l.g_rd = r.g_rd;
            continue;
        }
        // R's first record is an ordinary later record from L's point of view
        if ((unsigned long long)r.g_first_x * l.g_rd >= (unsigned long long)l.g_x * r.g_first_rd) {
            l.g_x = r.g_first_x; l.g_rd = r.g_first_rd;
        }
        if (r.g_n >= 2 && (unsigned long long)r.g_x * l.g_rd >= (unsigned long long)l.g_x * r.g_rd) {
            l.g_x = r.g_x; l.g_rd = r.g_rd;
        }
        l.g_n += r.g_n;
    }
}

// Per-(position, base) finalise (EE:1742-1797, EE:1258-1270): the two thresholds (NaN = "-1_-1"), the Germ_Max value
// and its state (0 absent, 1 one record = the floor, 2 value).
__device__ __forceinline__ void noise_final_base(const NoiseBase& s, const uint32_t n_records, const int i, float& q_fw,
                                                 float& q_bw, float& g, uint32_t& st) {
    const double n_rule = __dmul_rn(0.338, (double)n_records);  // EE:1742
    if ((double)s.count < n_rule) {
        q_fw = q_bw = __int_as_float(0x7fc00000);  // "-1_-1"
    } else {
        const double nt_fw = __dadd_rn((double)s.s_b_fw, s.s_p_fw);
        const double nt_bw = __dadd_rn((double)s.s_b_bw, s.s_p_bw);
        q_fw = __fdiv_rn(__double2float_rn(nt_fw), __double2float_rn((double)s.s_d_fw));  // EE:1762
        q_bw = __fdiv_rn(__double2float_rn(nt_bw), __double2float_rn((double)s.s_d_bw));  // EE:1763
        if (isnan(q_fw) || isnan(q_bw)) q_fw = q_bw = __int_as_float(0x7fc00000);         // EE:1765-1770
    }
    st = s.g_n == 0 ? 0u : (s.g_n == 1 ? 1u : 2u);
    g = st == 2 ? __fdiv_rn(__uint2float_rn(s.g_x), __uint2float_rn(s.g_rd))  // EE:1229-1232
                : (i == 0 ? -888.0f : 0.0f);                                     // EE:1260, EE:1318
}

__device__ __forceinline__ void noise_store_raw(int64_t slot, const float (&t)[8], const float (&g)[4], const uint32_t gs,
                                                const uint32_t (&c)[4], const uint32_t n_records, float* __restrict__ thr,
                                                float* __restrict__ germ_val, uint8_t* __restrict__ germ_state,
                                                uint32_t* __restrict__ count, uint32_t* __restrict__ nrec) {
    float4* t4 = reinterpret_cast<float4*>(thr + slot * 8);
    t4[0] = make_float4(t[0], t[1], t[2], t[3]);
    t4[1] = make_float4(t[4], t[5], t[6], t[7]);
    *reinterpret_cast<float4*>(germ_val + slot * 4) = make_float4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<uint32_t*>(germ_state + slot * 4) = gs;
    *reinterpret_cast<uint4*>(count + slot * 4) = make_uint4(c[0], c[1], c[2], c[3]);
    nrec[slot] = n_records;
}

// Per-position finalise and store for one slot.
__device__ __forceinline__ void noise_store(const NoiseAcc& a, int64_t slot, float* __restrict__ thr,
                                            float* __restrict__ germ_val, uint8_t* __restrict__ germ_state,
                                            uint32_t* __restrict__ count, uint32_t* __restrict__ nrec) {
    float t[8], g[4];
    uint32_t c[4];
    uint32_t gs = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t st;
        noise_final_base(a.b[i], a.nrec, i, t[2 * i], t[2 * i + 1], g[i], st);
        c[i] = a.b[i].count;
        gs |= st << (8 * i);
    }
    noise_store_raw(slot, t, g, gs, c, a.nrec, thr, germ_val, germ_state, count, nrec);
}

}  // namespace asdev
