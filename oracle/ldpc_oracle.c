/*
 * ldpc_oracle.c -- TEST INFRASTRUCTURE ONLY (see ldpc_oracle.h for the usage rules).
 * CPU restatement of the reference's fixed-point decode path.  Parity status: PINNED
 * (tests/test_oracle_vs_reference.py, tests/golden/).
 */
#include "ldpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- pairwise check operator -------------------------------------------------------
 * Reference: ArrayLDPC_Decoder.cpp:677-694.  sgn(t) = t > 0 ? 1 : -1
 * (ArrayLDPCMacro.h:222-224), Constant = int(5/8 * 2^FRAC_WIDTH) = 10 (:175),
 * WIDTH_MASK = 0xff (:29).  No saturation anywhere (quirk Q1); the 8-bit mask wraps
 * (quirk Q2). */
int oracle_sxor(int x, int y)
{
    const int constant = 10, mask = 0xff;
    int v1 = x < 0 ? -x : x;
    int v2 = y < 0 ? -y : y;
    int sum = (v1 + v2) & mask;
    int diff = (v1 > v2 ? v1 - v2 : v2 - v1) & mask;
    int part1 = constant - (sum >> 2);
    int part2 = constant - (diff >> 2);
    int sign = ((x > 0) == (y > 0)) ? 1 : -1; /* sgn(x)*sgn(y) with sgn(0) = -1 */
    if (part1 < 0) part1 = 0;
    if (part2 < 0) part2 = 0;
    return sign * ((v1 < v2 ? v1 : v2) + part1 - part2);
}

/* ---- syndrome on hard decisions ------------------------------------------------------
 * Reference: ArrayLDPC_Decoder.cpp:296-333 (checkPost_fp_general); the array variant
 * :375-420 visits the same edges through ROM::CirShift.  All n decisions are written
 * before the fail-fast scan starts (:305-308). */
int oracle_check_post(const oracle_code *code, const int *post, int *bits)
{
    for (int v = 0; v < code->n; ++v) bits[v] = post[v] > 0 ? 0 : 1;
    for (int c = 0; c < code->m; ++c) {
        const int *row = code->clist + (long)c * code->dc_max;
        int parity = 0;
        for (int k = 0; k < code->cdeg[c]; ++k) parity ^= bits[row[k]];
        if (parity) return 1;
    }
    return 0;
}

/* ---- one flooding iteration ----------------------------------------------------------
 * Check phase  : ArrayLDPC_Decoder.cpp:66-118  (forward/backward sxor recursion, in place)
 * Variable phase: ArrayLDPC_Decoder.cpp:121-156 (slot of v in check c == how many smaller
 *                 variables of c were already visited: the addr_count trick, :137,:153)
 * edge[k*m + c] is EdgeRAM[k].BRAM_fp[c]. */
static void check_phase(const oracle_code *code, int *edge, int *fwd, int *bwd)
{
    const int m = code->m;
    for (int c = 0; c < m; ++c) {
        const int d = code->cdeg[c];
        int *e = edge + c; /* stride m between slots */
        /* the reference also evaluates the two dead scan tails fwd[d-1], bwd[0] (Q4) */
        fwd[0] = e[0];
        bwd[d - 1] = e[(long)(d - 1) * m];
        for (int k = 1; k < d; ++k) {
            fwd[k] = oracle_sxor(fwd[k - 1], e[(long)k * m]);
            bwd[d - 1 - k] = oracle_sxor(bwd[d - k], e[(long)(d - 1 - k) * m]);
        }
        e[0] = bwd[1];
        e[(long)(d - 1) * m] = fwd[d - 2];
        for (int k = 1; k < d - 1; ++k) e[(long)k * m] = oracle_sxor(fwd[k - 1], bwd[k + 1]);
    }
}

static void variable_phase(const oracle_code *code, const int *llr, int *edge, int *post,
                           int *slot_count, int *c2v)
{
    const int m = code->m;
    memset(slot_count, 0, sizeof(int) * (size_t)m);
    for (int v = 0; v < code->n; ++v) {
        const int *col = code->vlist + (long)v * code->dv_max;
        const int d = code->vdeg[v];
        int acc = 0;
        for (int j = 0; j < d; ++j) {
            c2v[j] = edge[(long)slot_count[col[j]] * m + col[j]];
            acc += c2v[j];
        }
        acc += llr[v];
        post[v] = acc;
        for (int j = 0; j < d; ++j) {
            edge[(long)slot_count[col[j]] * m + col[j]] = acc - c2v[j];
            slot_count[col[j]]++;
        }
    }
}

/* Reference: ArrayLDPC_Decoder.cpp:18-171. */
int oracle_decode_general_fp(const oracle_code *code, const int *llr, int max_iter,
                             int *bits, int *post, int *edge)
{
    const int m = code->m;
    int *scratch = (int *)malloc(sizeof(int) * (size_t)(2 * code->dc_max + m + code->dv_max));
    int *fwd = scratch, *bwd = fwd + code->dc_max, *slot_count = bwd + code->dc_max,
        *c2v = slot_count + m;
    int iter = 0;

    /* :45-61  v2c(0) = channel value of the slot's variable */
    for (int c = 0; c < m; ++c)
        for (int k = 0; k < code->cdeg[c]; ++k)
            edge[(long)k * m + c] = llr[code->clist[(long)c * code->dc_max + k]];

    while (iter < max_iter) { /* :63 */
        check_phase(code, edge, fwd, bwd);
        variable_phase(code, llr, edge, post, slot_count, c2v);
        ++iter;                                            /* :157 */
        if (!oracle_check_post(code, post, bits)) break;   /* :164-167 */
    }
    free(scratch);
    return iter;
}

/* Reference: ArrayLDPC_Decoder.cpp:422-639 after setState(PCV) (PerfTest.cpp:121,180,298).
 * hardDecision (:270-294) overwrites DecodedCodeword and, when the channel word already
 * satisfies H, decode_fixpoint returns 0 before touching Posteriori_fp / EdgeRAM. */
int oracle_decode_fixpoint(const oracle_code *code, const int *llr, int max_iter,
                           int *bits, int *post, int *edge)
{
    if (!oracle_check_post(code, llr, bits)) return 0;
    return oracle_decode_general_fp(code, llr, max_iter, bits, post, edge);
}

long oracle_decode_many(const oracle_code *code, const int *llr, long frames, int max_iter,
                        int use_precheck, int *iters)
{
    int *bits = (int *)malloc(sizeof(int) * (size_t)code->n);
    int *post = (int *)malloc(sizeof(int) * (size_t)code->n);
    int *edge = (int *)calloc((size_t)code->dc_max * code->m, sizeof(int));
    long total = 0;
    for (long f = 0; f < frames; ++f) {
        const int *x = llr + f * code->n;
        int it = use_precheck ? oracle_decode_fixpoint(code, x, max_iter, bits, post, edge)
                              : oracle_decode_general_fp(code, x, max_iter, bits, post, edge);
        if (iters) iters[f] = it;
        total += it;
    }
    free(bits); free(post); free(edge);
    return total;
}

/* ---- BER bookkeeping -----------------------------------------------------------------
 * Reference: ArrayLDPC_Decoder.cpp:178-197 (setInfoBit), :707-722 (calculateBER). */
void oracle_set_info_bit(const char *in, int in_len, int k, int *true_info)
{
    int count = 0;
    for (int i = 0; i < in_len - 1; ++i)
        for (int j = 0; j < 8; ++j) true_info[count++] = (in[i] >> j) & 1;
    for (int j = 0; j < k % 8; ++j) true_info[count++] = (in[in_len - 1] >> j) & 1;
}

int oracle_calculate_ber(const int *bits, const int *info_index, const int *true_info, int k)
{
    int errors = 0;
    for (int i = 0; i < k; ++i) errors += bits[info_index[i]] != true_info[i];
    return errors;
}

/* ---- channel ---------------------------------------------------------------------------
 * Reference: rngs.cpp:52-69 (Random), rvgs.cpp:152-181 (Normal), PerfTest.cpp:108-120. */
double oracle_random(long *seed)
{
    const long modulus = 2147483647L, multiplier = 48271L;
    const long q = modulus / multiplier, r = modulus % multiplier;
    long t = multiplier * (*seed % q) - r * (*seed / q);
    *seed = t > 0 ? t : t + modulus;
    return (double)*seed / modulus;
}

double oracle_normal(long *seed, double mean, double sd)
{
    static const double p[5] = {0.322232431088, 1.0, 0.342242088547, 0.204231210245e-1,
                                0.453642210148e-4};
    static const double q[5] = {0.099348462606, 0.588581570495, 0.531103462366,
                                0.103537752850, 0.385607006340e-2};
    double u = oracle_random(seed);
    double t = u < 0.5 ? sqrt(-2.0 * log(u)) : sqrt(-2.0 * log(1.0 - u));
    double num = p[0] + t * (p[1] + t * (p[2] + t * (p[3] + t * p[4])));
    double den = q[0] + t * (q[1] + t * (q[2] + t * (q[3] + t * q[4])));
    double z = u < 0.5 ? (num / den) - t : t - (num / den);
    return mean + sd * z;
}

void oracle_channel_frame(long *seed, const int *codeword, int n, double snr, double sigma,
                          int frac_width, int *llr_fp)
{
    for (int i = 0; i < n; ++i) {
        int c = codeword ? codeword[i] : 0;
        double llr = 2 * snr * (1 - 2 * c + oracle_normal(seed, 0, sigma));
        llr_fp[i] = (int)(llr * (1 << frac_width)); /* truncation toward zero, Q11 */
    }
}

/* ---- encoder ---------------------------------------------------------------------------
 * Reference: ArrayLDPC_Encoder.cpp:160-225.  Info bits are unpacked LSB first, the last
 * byte supplying (n - rows) % 8 bits (:179-183); info bits land on the unflagged columns
 * in file order (:56-69, :184-187); parity i = XOR of the unflagged members of row i and
 * lands on the i-th flagged column (:199-210). */
void oracle_encode(const oracle_gen *g, const char *in, int in_len, int *codeword)
{
    const int k = g->n - g->rows;
    int *info = (int *)calloc((size_t)k + 8, sizeof(int));
    int count = 0, ii = 0, pi = 0;
    for (int i = 0; i < in_len - 1; ++i)
        for (int j = 0; j < 8; ++j) info[count++] = (in[i] >> j) & 1;
    for (int j = 0; j < k % 8; ++j) info[count++] = (in[in_len - 1] >> j) & 1;

    for (int v = 0; v < g->n; ++v)
        if (!g->flag[v]) codeword[v] = info[ii++];
    for (int v = 0; v < g->n; ++v) {
        if (!g->flag[v]) continue;
        const int *row = g->mlist + (long)pi * g->stride;
        int parity = 0;
        for (int j = 0; j < g->deg[pi]; ++j)
            if (!g->flag[row[j]]) parity ^= codeword[row[j]];
        codeword[v] = parity;
        ++pi;
    }
    free(info);
}

/* ------------------------------------------------------------------------------------------
 * floating-point decoder (dead code in the reference's drivers, kept as the fixed-point loss yardstick)
 * ------------------------------------------------------------------------------------------ */
#define ORACLE_MAX_DEG 64
static int sgn_d(double x) { return x > 0 ? 1 : -1; } /* ArrayLDPCMacro.h:218-221 */

/* ArrayLDPC_Decoder.cpp:724-732 */
double oracle_sxor_f64(double x, double y)
{
    double v1 = fabs(x), v2 = fabs(y);
    double sum_abs = v1 + v2;
    double diff_abs = fabs(v1 - v2);
    double mn = v2 < v1 ? v2 : v1; /* std::min(v1, v2) */
    return sgn_d(x) * sgn_d(y) * (mn + log(1 + exp(-sum_abs)) - log(1 + exp(-diff_abs)));
}

/* ArrayLDPC_Decoder.cpp:735-933 */
int oracle_decode_general_f64(const oracle_code *code, const double *llr, int max_iter,
                              int *bits, double *post, double *edge)
{
    const int n = code->n, m = code->m;
    double fwd[ORACLE_MAX_DEG], bwd[ORACLE_MAX_DEG], mv[ORACLE_MAX_DEG];
    int *addr_count = (int *)malloc(sizeof(int) * (size_t)m);
    int it = 0, c, k, v, j;
    /* :757-774 */
    for (c = 0; c < m; c++)
        for (k = 0; k < code->cdeg[c]; k++) edge[(size_t)k * m + c] = llr[code->clist[(size_t)c * code->dc_max + k]];
    while (it < max_iter) {
        /* check phase :779-833 */
        for (c = 0; c < m; c++) {
            const int d = code->cdeg[c];
            for (k = 0; k < d; k++) mv[k] = edge[(size_t)k * m + c];
            fwd[0] = mv[0];
            bwd[d - 1] = mv[d - 1];
            for (k = 1; k < d; k++) {
                fwd[k] = oracle_sxor_f64(fwd[k - 1], mv[k]);
                bwd[d - k - 1] = oracle_sxor_f64(bwd[d - k], mv[d - 1 - k]);
            }
            edge[c] = bwd[1];
            edge[(size_t)(d - 1) * m + c] = fwd[d - 2];
            for (k = 1; k < d - 1; k++) edge[(size_t)k * m + c] = oracle_sxor_f64(fwd[k - 1], bwd[k + 1]);
        }
        /* variable phase :835-872 */
        memset(addr_count, 0, sizeof(int) * (size_t)m);
        for (v = 0; v < n; v++) {
            const int d = code->vdeg[v];
            double accum = 0, mc[ORACLE_MAX_DEG];
            for (j = 0; j < d; j++) {
                const int chk = code->vlist[(size_t)v * code->dv_max + j];
                mc[j] = edge[(size_t)addr_count[chk] * m + chk];
                accum = accum + mc[j];
            }
            accum = accum + llr[v];
            post[v] = accum;
            for (j = 0; j < d; j++) {
                const int chk = code->vlist[(size_t)v * code->dv_max + j];
                edge[(size_t)addr_count[chk] * m + chk] = accum - mc[j];
                addr_count[chk]++;
            }
        }
        it++;
        /* checkPost :335-372 */
        {
            int fail = 0;
            for (v = 0; v < n; v++) bits[v] = post[v] > 0 ? 0 : 1;
            for (c = 0; c < m && !fail; c++) {
                int sum = 0;
                for (k = 0; k < code->cdeg[c]; k++) sum ^= bits[code->clist[(size_t)c * code->dc_max + k]];
                if (sum) fail = 1;
            }
            if (!fail) break;
        }
    }
    free(addr_count);
    return it;
}
