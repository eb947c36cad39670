/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of AmpliSolve's hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (amplisolve_b200/) never does.
 *
 * PARITY PINNED: validated in tests/ against the real reference compiled from /root/reference
 * into oracle/_ref (Toy_data end-to-end, function-level grids for kf_gammaq / Q score) and
 * against the golden fixtures under tests/golden/ generated from that reference.
 * aso_fisher follows the Boost stand-in header (Boost 1.61 is a missing blob of the reference); it is pinned to
 * printed precision against SciPy's compiled-in Boost.Math pdf (tests/golden/fisher_boost.npz).
 *
 * Deliberately written the way the REFERENCE works -- one text-keyed record at a time, float
 * divides, no early-outs -- and on a different data representation (file rows keyed by a
 * position id) than the CUDA path (dense slot tensors), so that it checks the layout logic too.
 *
 * Citations: EE = /root/reference/source_codes/AmpliSolveErrorEstimation.cpp,
 *            VC = /root/reference/source_codes/AmpliSolveVariantCalling.cpp.
 */
#ifndef AS_ORACLE_H
#define AS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A file row as the reference parses it (EE:1149, VC:752): the nine integers after the six
 * text columns: A C G T RD Ars Crs Grs Trs.  pos_id = index of the row's "chrom_pos" key among
 * the unique panel positions, or -1 when the key is not in the panel. */
typedef struct {
    int32_t pos_id;
    int32_t A, C, G, T, RD, Ars, Crs, Grs, Trs;
} aso_row;

/* ---- noise half ------------------------------------------------------------------------- */
/* rows of sample s = rows[row_off[s] .. row_off[s+1]) in FILE order; samples already in the
 * reference's hash-iteration order (EE:1081).  Outputs per unique position u and base b:
 *   thr[u][b][0..1]  fw/bw threshold as float; NaN encodes the "-1_-1" text (EE:1742-1770)
 *   germ_val/germ_present[u][b]  Germ_Max value (double holding a widened float, -888 or 0) and
 *                    whether the key exists (EE:1251-1271 and copies)
 *   count[u][b]      records that passed the filter (EE:1626)
 *   nrec[u]          records of the key = Value_Hash.count() (EE:1742)                        */
void aso_noise_estimate(const aso_row* rows, const int64_t* row_off, int S, int32_t U, float C, int cut, float* thr,
                        double* germ_val, uint8_t* germ_present, int32_t* count, int32_t* nrec);

/* Text cells of the noise table (EE:2668-2690, EE:2807-2849).  Return bytes written. */
int aso_format_thr_cell(float thr_fw, float thr_bw, int base_is_ref, char* out);
int aso_format_germ_cell(double v, int present, char* out);
/* The "%f" text round trip of a threshold as the caller sees it (EE:1787 -> VC:889-890). */
float aso_thr_as_caller_sees(float thr);

/* ---- caller half ------------------------------------------------------------------------ */
double aso_kf_lgamma(double z);               /* VC:3817-3830 */
double aso_kf_gammaq(double s, double z);     /* VC:3726-3729 with VC:3733-3752, VC:3785-3794 */
double aso_poisson_p(int k, int rd, float err);            /* the double p of VC:3858-3866 (1 when k==0) */
long double aso_poisson_q_ld(int k, int rd, float err);    /* VC:3834-3884 */
double aso_poisson_q(int k, int rd, float err);            /* (double) of the above */
double aso_fisher(int a, int b, int c, int d);             /* VC:3797-3814 over the Boost stand-in pdf */
int aso_q_at_least(double p, int threshold);               /* (long double)Q(p) >= threshold, VC:3868-3882 + VC:898 / VC:1023 */
double aso_q_from_p(double p);                             /* (double) of that long double Q */

typedef struct {
    int32_t sample;   /* index into the sample list as given */
    int32_t row;      /* row index within that sample's file */
    int32_t pos_id;
    int8_t ref;       /* 0..3 = A,C,G,T */
    int8_t alt;       /* 0..3 */
    int8_t pad[2];
    int32_t k_fw, k_bw, FW, BW;
    double p_fw, p_bw;   /* raw double p-values of the two strand tests */
    double q_fw, q_bw;   /* (double) of the long double Q scores */
    double fisher_p;
} aso_call;

/* Every tumour row x alt allele, in the reference's order (sample, row, alt order of
 * VC:869-3288).  ref[u] in 0..3, anything else = not callable (VC:3290).  thr[u][b][2] are the
 * floats the caller parsed from the noise table.  Returns the number of calls; writes at most
 * cap of them. */
int64_t aso_call_variants(const aso_row* rows, const int64_t* row_off, int T, int32_t U, const uint8_t* ref,
                          const float* thr, int cut, aso_call* out, int64_t cap);

/* VC:3615-3718 and the padding rules of VC:3307-3611 (neighbour bases: 0..3, 4 = other letter, 255 = absent). */
int aso_homopolymer(const char* down, const char* up, char alt);

#ifdef __cplusplus
}
#endif
#endif
