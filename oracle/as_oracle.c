/* TEST INFRASTRUCTURE ONLY -- see as_oracle.h.  CPU restatement of AmpliSolve's hot path, used
 * solely as the parity checker and as the "port" CPU baseline.  Plain C11, scalar, fp-contract
 * off (see oracle/Makefile).  EE/VC citations refer to the files named in as_oracle.h. */
#define _GNU_SOURCE
#include "as_oracle.h"
#include "factorials.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* noise half                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* the four numbers the reference stores per (row, base) as "bfw_FW_bbw_BW" text (EE:1241-1245) */
typedef struct {
    int32_t b_fw, FW, b_bw, BW;
} noise_rec;

static void split_row(const aso_row* r, int32_t tot[4], int32_t rs[4], int32_t fw[4], int32_t* FW, int32_t* BW) {
    tot[0] = r->A;  tot[1] = r->C;  tot[2] = r->G;  tot[3] = r->T;
    rs[0] = r->Ars; rs[1] = r->Crs; rs[2] = r->Grs; rs[3] = r->Trs;
    *FW = 0; *BW = 0;
    for (int b = 0; b < 4; ++b) {           /* EE:1155-1176, VC:760-770 */
        fw[b] = tot[b] - rs[b];
        *FW += fw[b];
        *BW += rs[b];
    }
}

void aso_noise_estimate(const aso_row* rows, const int64_t* row_off, int S, int32_t U, float C, int cut, float* thr,
                        double* germ_val, uint8_t* germ_present, int32_t* count, int32_t* nrec) {
    const int64_t R = row_off[S];
    /* the multimap of EE:222 as a CSR keyed by pos_id, records kept in insertion order */
    int64_t* start = (int64_t*)calloc((size_t)U + 1, sizeof(int64_t));
    for (int64_t i = 0; i < R; ++i)
        if (rows[i].pos_id >= 0) start[rows[i].pos_id + 1]++;
    for (int32_t u = 0; u < U; ++u) start[u + 1] += start[u];
    noise_rec* recs = (noise_rec*)malloc(sizeof(noise_rec) * 4 * (size_t)(start[U] ? start[U] : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)U + 1));
    memcpy(fill, start, sizeof(int64_t) * ((size_t)U + 1));
    memset(germ_present, 0, (size_t)U * 4);
    for (int64_t i = 0; i < (int64_t)U * 4; ++i) germ_val[i] = 0.0;

    /* storeGermlineStatistics: files in the given order, rows in file order (EE:1081-1467) */
    for (int s = 0; s < S; ++s) {
        for (int64_t i = row_off[s]; i < row_off[s + 1]; ++i) {
            const aso_row* r = &rows[i];
            if (r->pos_id < 0) continue; /* keys outside the panel are stored but never read back */
            int32_t tot[4], rs[4], fw[4], FW, BW;
            split_row(r, tot, rs, fw, &FW, &BW);
            const int64_t slot = fill[r->pos_id]++;
            for (int b = 0; b < 4; ++b) {
                noise_rec* nr = &recs[slot * 4 + b];
                nr->b_fw = fw[b]; nr->FW = FW; nr->b_bw = rs[b]; nr->BW = BW;
                /* Germ_Max (EE:1229-1232, EE:1251-1271 for A; C/G/T copies start from 0 not -888) */
                float af_total = (float)tot[b] / (float)r->RD;
                if ((double)af_total <= 0.05 && FW >= cut && BW >= cut) {
                    const int64_t g = (int64_t)r->pos_id * 4 + b;
                    if (!germ_present[g]) {
                        germ_present[g] = 1;
                        germ_val[g] = (b == 0) ? -888.0 : 0.0; /* this record's AF is dropped */
                    } else if (germ_val[g] <= (double)af_total) {
                        germ_val[g] = (double)af_total;
                    }
                }
            }
        }
    }

    /* estimateThresholds (EE:1536-2544), one (position, base) at a time */
    for (int32_t u = 0; u < U; ++u) {
        const int64_t n = start[u + 1] - start[u];
        nrec[u] = (int32_t)n;
        for (int b = 0; b < 4; ++b) {
            double sum_fw_nt = 0, sum_fw_RD = 0, sum_bw_nt = 0, sum_bw_RD = 0;
            int cnt = 0;
            for (int64_t j = start[u]; j < start[u + 1]; ++j) {
                const noise_rec* nr = &recs[j * 4 + b];
                double AF_limit = 0.05;
                double AF_fw = (float)nr->b_fw / (float)nr->FW; /* fp32 divide, widened (EE:1613) */
                double AF_bw = (float)nr->b_bw / (float)nr->BW;
                if (AF_fw <= AF_limit && AF_bw <= AF_limit && nr->FW >= cut && nr->BW >= cut) {
                    sum_fw_nt = sum_fw_nt + nr->b_fw + ((float)nr->FW * (float)C); /* EE:1617 */
                    sum_fw_RD = sum_fw_RD + nr->FW;
                    sum_bw_nt = sum_bw_nt + nr->b_bw + ((float)nr->BW * (float)C);
                    sum_bw_RD = sum_bw_RD + nr->BW;
                    cnt++;
                }
            }
            float* t = &thr[((int64_t)u * 4 + b) * 2];
            count[(int64_t)u * 4 + b] = cnt;
            if ((double)cnt < 0.338 * (double)n) { /* EE:1742 */
                t[0] = t[1] = NAN;
            } else {
                float q_fw = (float)sum_fw_nt / (float)sum_fw_RD; /* EE:1762-1763 */
                float q_bw = (float)sum_bw_nt / (float)sum_bw_RD;
                if (isnan(q_fw) || isnan(q_bw)) {
                    t[0] = t[1] = NAN;
                } else {
                    t[0] = q_fw;
                    t[1] = q_bw;
                }
            }
        }
    }
    free(fill);
    free(recs);
    free(start);
}

int aso_format_thr_cell(float thr_fw, float thr_bw, int base_is_ref, char* out) {
    if (base_is_ref) return sprintf(out, "-2_-2");                       /* EE:2668-2673 */
    if (isnan(thr_fw) || isnan(thr_bw)) return sprintf(out, "0.01_0.01"); /* "-1_-1" -> EE:2680-2684 */
    return sprintf(out, "%f_%f", (double)thr_fw, (double)thr_bw);        /* EE:1787 */
}

int aso_format_germ_cell(double v, int present, char* out) {
    if (!present) return sprintf(out, "-"); /* EE:2807-2849 */
    return sprintf(out, "%g", v);           /* ostream << double, default precision 6 */
}

float aso_thr_as_caller_sees(float thr) {
    char buf[64];
    snprintf(buf, sizeof buf, "%f", (double)thr);
    return strtof(buf, NULL); /* std::stof, VC:889-890 */
}

/* ------------------------------------------------------------------------------------------ */
/* caller half: kfunc restatement                                                               */
/* ------------------------------------------------------------------------------------------ */

#define GAMMA_EPS 1e-14 /* VC:149 */
#define GAMMA_TINY 1e-290 /* VC:150 */

double aso_kf_lgamma(double z) {
    double x = 0;
    x += 0.1659470187408462e-06 / (z + 7);
    x += 0.9934937113930748e-05 / (z + 6);
    x -= 0.1385710331296526 / (z + 5);
    x += 12.50734324009056 / (z + 4);
    x -= 176.6150291498386 / (z + 3);
    x += 771.3234287757674 / (z + 2);
    x -= 1259.139216722289 / (z + 1);
    x += 676.5203681218835 / z;
    x += 0.9999999999995183;
    return log(x) - 5.58106146679532777 - z + (z - 0.5) * log(z + 6.5);
}

/* lower regularised gamma by its power series, at most 99 terms (VC:3785-3794) */
static double lower_series(double s, double z) {
    double sum = 1.0, x = 1.0;
    for (int k = 1; k < 100; ++k) {
        x *= z / (s + k);
        sum += x;
        if (x / sum < GAMMA_EPS) break;
    }
    return exp(s * log(z) - z - aso_kf_lgamma(s + 1.) + log(sum));
}

/* upper regularised gamma by modified Lentz, at most 99 steps (VC:3733-3752) */
static double upper_cf(double s, double z) {
    double f = 1. + z - s, C = f, D = 0.;
    for (int j = 1; j < 100; ++j) {
        double a = j * (s - j), b = (j << 1) + 1 + z - s, d;
        D = b + a * D;
        if (D < GAMMA_TINY) D = GAMMA_TINY;
        C = b + a / C;
        if (C < GAMMA_TINY) C = GAMMA_TINY;
        D = 1. / D;
        d = C * D;
        f *= d;
        if (fabs(d - 1.) < GAMMA_EPS) break;
    }
    return exp(s * log(z) - z - aso_kf_lgamma(s) - log(f));
}

double aso_kf_gammaq(double s, double z) { /* VC:3726-3729 */
    return (z <= 1. || z < s) ? 1. - lower_series(s, z) : upper_cf(s, z);
}

double aso_poisson_p(int k, int rd, float err) {
    if (err == 0) err = 0.0010008; /* VC:3852-3856: double literal narrowed to float */
    if (k == 0) return 1.0;
    double m = (double)rd * err; /* VC:3864: double * float */
    return 1 - aso_kf_gammaq(k, m);
}

long double aso_poisson_q_ld(int k, int rd, float err) {
    if (err == -1) return -888; /* VC:3844-3849 */
    long double pvalue = aso_poisson_p(k, rd, err);
    long double p_limit = 0.0000000001; /* VC:3838 */
    if (pvalue < p_limit) return -10 * log10l(p_limit);
    if (pvalue == 1) return 0;
    return -10 * log10l(pvalue);
}

double aso_poisson_q(int k, int rd, float err) { return (double)aso_poisson_q_ld(k, rd, err); }

/* The long double Q the reference derives from a double p (VC:3868-3882), and its comparison with 5 (VC:898). */
static long double q_from_p_ld(double p) {
    long double pvalue = p;
    long double p_limit = 0.0000000001;
    if (pvalue < p_limit) return -10 * log10l(p_limit);
    if (pvalue == 1) return 0;
    return -10 * log10l(pvalue);
}
int aso_q_at_least(double p, int threshold) { return q_from_p_ld(p) >= threshold; }
double aso_q_from_p(double p) { return (double)q_from_p_ld(p); }

/* two-sided Fisher exact test as VC:3797-3814, over the stand-in hypergeometric pdf */
/* operands in a canonical order (smaller of k, n - k first): C(n,k) == C(n,n-k) bitwise, so terms tied mathematically stay tied */
static double log_choose(double n, double k) {
    const double lo = k < n - k ? k : n - k, hi = k < n - k ? n - k : k;
    return (lgamma(n + 1.0) - lgamma(lo + 1.0)) - lgamma(hi + 1.0);
}
/* N <= 170: Boost.Math's factorial-table method (hypergeometric_pdf_factorial_imp), bit-identical to SciPy's compiled-in Boost */
static double hyper_pdf_factorial(unsigned r, unsigned n, unsigned N, unsigned k) {
    double result = ASO_FACTORIAL[n];
    const double num[3] = {ASO_FACTORIAL[r], ASO_FACTORIAL[N - n], ASO_FACTORIAL[N - r]};
    const double den[5] = {ASO_FACTORIAL[N], ASO_FACTORIAL[k], ASO_FACTORIAL[n - k], ASO_FACTORIAL[r - k], ASO_FACTORIAL[N - n - r + k]};
    int i = 0, j = 0;
    while (i < 3 || j < 5) {
        while (j < 5 && (result >= 1 || i >= 3)) result /= den[j++];
        while (i < 3 && (result <= 1 || j >= 5)) result *= num[i++];
    }
    return result > 1 ? 1.0 : result;
}
static double hyper_pdf(unsigned r, unsigned n, unsigned N, unsigned k) {
    if (N <= 170) return hyper_pdf_factorial(r, n, N, k);
    return exp(log_choose(r, k) + log_choose((double)N - r, (double)n - k) - log_choose(N, n));
}
double aso_fisher(int a, int b, int c, int d) {
    unsigned N = a + b + c + d, r = a + c, n = c + d;
    unsigned hi = r < n ? r : n;
    int lo_i = (int)(r + n - N);
    unsigned lo = lo_i > 0 ? (unsigned)lo_i : 0u;
    double cutoff = hyper_pdf(r, n, N, (unsigned)c), acc = 0.0;
    for (int k = (int)lo; k < (int)hi + 1; ++k) {
        double p = hyper_pdf(r, n, N, (unsigned)k);
        if (p <= cutoff) acc += p;
    }
    return acc;
}

int64_t aso_call_variants(const aso_row* rows, const int64_t* row_off, int T, int32_t U, const uint8_t* ref,
                          const float* thr, int cut, aso_call* out, int64_t cap) {
    (void)U;
    int64_t ncall = 0;
    for (int s = 0; s < T; ++s) {
        for (int64_t i = row_off[s]; i < row_off[s + 1]; ++i) {
            const aso_row* r = &rows[i];
            if (r->pos_id < 0) continue;   /* "mistake..." VC:847-852 */
            const int rb = ref[r->pos_id];
            if (rb > 3) continue;          /* VC:3290-3293 */
            int32_t tot[4], rs[4], fw[4], FW, BW;
            split_row(r, tot, rs, fw, &FW, &BW);
            const int32_t rd_rev = rs[3] + rs[2] + rs[1] + rs[0]; /* VC:819 */
            for (int alt = 0; alt < 4; ++alt) { /* increasing base order = the order of VC:869-3288 */
                if (alt == rb) continue;
                const float e_fw = thr[((int64_t)r->pos_id * 4 + alt) * 2 + 0];
                const float e_bw = thr[((int64_t)r->pos_id * 4 + alt) * 2 + 1];
                long double Q_fw = aso_poisson_q_ld(tot[alt] - rs[alt], r->RD - rd_rev, e_fw); /* VC:895 */
                long double Q_bw = aso_poisson_q_ld(rs[alt], rd_rev, e_bw);                    /* VC:896 */
                if (FW >= cut && BW >= cut && Q_fw >= 5 && Q_bw >= 5) {                        /* VC:898 */
                    if (ncall < cap) {
                        aso_call* c = &out[ncall];
                        memset(c, 0, sizeof *c);
                        c->sample = s;
                        c->row = (int32_t)(i - row_off[s]);
                        c->pos_id = r->pos_id;
                        c->ref = (int8_t)rb;
                        c->alt = (int8_t)alt;
                        c->k_fw = fw[alt]; c->k_bw = rs[alt]; c->FW = FW; c->BW = BW;
                        c->p_fw = aso_poisson_p(fw[alt], r->RD - rd_rev, e_fw);
                        c->p_bw = aso_poisson_p(rs[alt], rd_rev, e_bw);
                        c->q_fw = (double)Q_fw;
                        c->q_bw = (double)Q_bw;
                        c->fisher_p = aso_fisher(r->RD - rd_rev, rd_rev, fw[alt], rs[alt]); /* VC:902 */
                    }
                    ncall++;
                }
            }
        }
    }
    return ncall;
}

int aso_homopolymer(const char* down, const char* up, char alt) { /* VC:3615-3718 */
    int n[4] = {0, 0, 0, 0};
    const char* letters = "ACGT";
    for (int b = 0; b < 4; ++b)
        if (alt == letters[b]) n[b] = 1;
    const char* strs[2] = {down, up};
    for (int t = 0; t < 2; ++t)
        for (const char* p = strs[t]; *p; ++p)
            for (int b = 0; b < 4; ++b)
                if (*p == letters[b]) n[b]++;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (n[i] + n[j] > 18) return 1;
    return 0;
}
