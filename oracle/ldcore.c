/* ldcore.c -- CPU restatement of the `tomahawk calc` arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity oracle for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it. The product library (libtwkb.so) neither
 * links nor calls anything in oracle/.
 *
 * Parity status: PINNED. tests/test_oracle.py diffs this file
 * against the reference's own `calc` binary (oracle/_ref/tomahawk_calc, built
 * from /root/reference by oracle/build_ref.sh) on shared synthetic .twk inputs,
 * against the reference's kt_fisher_exact (oracle/_ref/libref_fisher.so), and
 * against the five worked phased rows of docs/tutorial.md:608-612; the outputs
 * of those reference runs are committed under tests/golden/ so the pin also
 * holds where /root/reference is absent.
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths relative to the reference tree). Floating point is IEEE double in the
 * reference's operation order; build with -ffp-contract=off (no FMA), which is
 * what the reference's -msse4.2 x86-64 build executes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LD_MIN_ALLELES 5        /* lib/ld/ld_engine.h:36 */
#define LD_ROUNDING_SLACK 1e-5  /* lib/ld/ld_engine.h:37 */
#define LD_LOW_AC 5             /* lib/ld/ld_engine.h:33 */
#define LD_BAD_HWE 1e-4         /* lib/ld/ld_engine.h:34 */
#define LD_LONG_RANGE 500e3     /* lib/ld/ld_engine.h:35 */
#define LD_RECORD_BYTES 106     /* include/core.h:758-759 */

/* per-variant metadata: the subset of twk1_t (include/core.h:291-295) calc reads */
typedef struct {
    uint32_t rid, pos, ac, an;
    double hwe;
    uint8_t gt_missing, gt_phase, pad[6];
} ld_variant;

/* mirrors the fields of twk_ld_settings (include/core.h:909-924) calc reads */
typedef struct {
    double minP, minR2, maxR2, minDprime, maxDprime;
    int32_t force_phased, forced_unphased; /* -p / -u; neither = auto mode */
    int32_t window, l_window;              /* -w */
    int32_t emulate_quirks;                /* 1: reproduce Q1/Q3 of SURVEY.md App. C */
    int32_t block_size;                    /* .twk block length (500, lib/importer.h:36) */
    int32_t skip_min_cell_rule;            /* test-only: bypass the "< 5" rule (:1174-1186) so the
                                              tutorial rows, which predate it, can be replayed */
    int32_t bitmaps;                       /* -p -m -M: CalculatePhasedBitmap(Window) (:2351-2524) */
} ld_params;

/* ---------------------------------------------------------------- Fisher exact
 * lib/fisher_math.cpp:183-267 (htslib kfunc). */
static double log_binom(int n, int k) { /* :183-187 */
    if (k == 0 || n == k) return 0;
    return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1);
}

static double hyper_pmf(int n11, int n1_, int n_1, int n) { /* :195-198 */
    return exp(log_binom(n1_, n11) + log_binom(n - n1_, n_1 - n11) - log_binom(n, n_1));
}

typedef struct {
    int n11, n1_, n_1, n;
    double p;
} hyper_state;

/* :206-229 -- incremental pmf; recomputed from scratch whenever n11 % 11 == 0,
 * the n22 cell is empty, or the step is not +-1. */
static double hyper_step(int n11, int n1_, int n_1, int n, hyper_state* st) {
    if (n1_ || n_1 || n) {
        st->n11 = n11; st->n1_ = n1_; st->n_1 = n_1; st->n = n;
    } else {
        if (n11 % 11 && n11 + st->n - st->n1_ - st->n_1) {
            if (n11 == st->n11 + 1) {
                st->p *= (double)(st->n1_ - st->n11) / n11 * (st->n_1 - st->n11) / (n11 + st->n - st->n1_ - st->n_1);
                st->n11 = n11;
                return st->p;
            }
            if (n11 == st->n11 - 1) {
                st->p *= (double)st->n11 / (st->n1_ - n11) * (st->n11 + st->n - st->n1_ - st->n_1) / (st->n_1 - n11);
                st->n11 = n11;
                return st->p;
            }
        }
        st->n11 = n11;
    }
    st->p = hyper_pmf(st->n11, st->n1_, st->n_1, st->n);
    return st->p;
}

/* :231-267 -- returns the two-sided P (the only output calc uses). */
double ldcore_fisher(int n11, int n12, int n21, int n22, double* left_out, double* right_out) {
    int i, j, max, min;
    double p, q, left, right, two;
    hyper_state st;
    int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
    max = (n_1 < n1_) ? n_1 : n1_;
    min = n1_ + n_1 - n;
    if (min < 0) min = 0;
    if (left_out) *left_out = 1.;
    if (right_out) *right_out = 1.;
    if (min == max) return 1.;
    q = hyper_step(n11, n1_, n_1, n, &st);
    p = hyper_step(min, 0, 0, 0, &st);
    for (left = 0., i = min + 1; p < 0.99999999 * q && i <= max; ++i) {
        left += p;
        p = hyper_step(i, 0, 0, 0, &st);
    }
    --i;
    if (p < 1.00000001 * q) left += p; else --i;
    p = hyper_step(max, 0, 0, 0, &st);
    for (right = 0., j = max - 1; p < 0.99999999 * q && j >= 0; --j) {
        right += p;
        p = hyper_step(j, 0, 0, 0, &st);
    }
    ++j;
    if (p < 1.00000001 * q) right += p; else ++j;
    two = left + right;
    if (two > 1.) two = 1.;
    if (abs(i - n11) < abs(j - n11)) right = 1. - left + q; else left = 1.0 - right + q;
    if (left_out) *left_out = left;
    if (right_out) *right_out = right;
    return two;
}

/* ------------------------------------------------------------- output record
 * lib/core.cpp:470-490: u16 flags, u32 ridA, ridB, u32 posA<<2, u32 posB<<2,
 * f64 cnt[4], D, Dprime, R, R2, P, ChiSqFisher, ChiSqModel. */
typedef struct {
    uint16_t flags;
    double cnt[4], D, Dprime, R, R2, P, chi_fisher, chi_model;
} ld_stats;

static void put_record(uint8_t* dst, const ld_stats* s, const ld_variant* a, const ld_variant* b) {
    uint32_t pa = a->pos << 2, pb = b->pos << 2; /* Amiss/Aphased bits are never set by calc */
    memcpy(dst + 0, &s->flags, 2);
    memcpy(dst + 2, &a->rid, 4);
    memcpy(dst + 6, &b->rid, 4);
    memcpy(dst + 10, &pa, 4);
    memcpy(dst + 14, &pb, 4);
    memcpy(dst + 18, s->cnt, 32);
    memcpy(dst + 50, &s->D, 8);
    memcpy(dst + 58, &s->Dprime, 8);
    memcpy(dst + 66, &s->R, 8);
    memcpy(dst + 74, &s->R2, 8);
    memcpy(dst + 82, &s->P, 8);
    memcpy(dst + 90, &s->chi_fisher, 8);
    memcpy(dst + 98, &s->chi_model, 8);
}

/* flag bits shared by both maths: lib/ld/ld_engine.cpp:1244-1255, 1674-1684 */
static uint16_t variant_flags(const ld_variant* a, const ld_variant* b) {
    uint16_t f = 0;
    int same = a->rid == b->rid;
    int32_t diff = (int32_t)a->pos - (int32_t)b->pos;
    if (same) f |= 1u << 1;
    if (abs(diff) > LD_LONG_RANGE && same) f |= 1u << 2;
    if (a->an) f |= 1u << 8;
    if (b->an) f |= 1u << 9;
    if (a->ac < LD_LOW_AC) f |= 1u << 10;
    if (b->ac < LD_LOW_AC) f |= 1u << 11;
    if (a->hwe < LD_BAD_HWE) f |= 1u << 12;
    if (b->hwe < LD_BAD_HWE) f |= 1u << 13;
    return f;
}

/* ------------------------------------------------------------------ phased math
 * lib/ld/ld_engine.cpp:1162-1259. Cells in the reference's slot order:
 * c0 = [REFREF], c1 = [ALTREF] (slot 1), c4 = [REFALT] (slot 4), c5 = [ALTALT].
 * Returns 1 and fills *s when the pair passes every filter. */
int ldcore_phased_stats(uint64_t c0, uint64_t c1, uint64_t c4, uint64_t c5, const ld_params* prm,
                        const ld_variant* a, const ld_variant* b, ld_stats* s) {
    uint64_t T = c0 + c4 + c1 + c5; /* :1164 */
    if (T < LD_MIN_ALLELES) return 0;
    if (!prm->skip_min_cell_rule) {
        if (c0 < c5) { /* :1174-1186 */
            if (c4 + c1 + c0 < 5) return 0;
        } else {
            if (c5 + c4 + c1 < 5) return 0;
        }
    }
    double pA = (double)c0 / T, qA = (double)c1 / T, pB = (double)c4 / T, qB = (double)c5 / T; /* :1189-1192 */
    if (pA * qB - qA * pB == 0) return 0;
    const double g0 = ((double)c0 + c4) / T; /* :1197-1200 */
    const double g1 = ((double)c1 + c5) / T;
    const double h0 = ((double)c0 + c1) / T;
    const double h1 = ((double)c4 + c5) / T;
    s->D = pA * qB - qA * pB;
    s->R2 = s->D * s->D / (g0 * g1 * h0 * h1);
    if (s->R2 < prm->minR2 || s->R2 > prm->maxR2) return 0;
    double dmax;
    if (s->D >= 0) dmax = g0 * h1 < h0 * g1 ? g0 * h1 : h0 * g1; /* :1210-1211 */
    else dmax = g0 * g1 < h0 * h1 ? -g0 * g1 : -h0 * h1;
    s->Dprime = s->D / dmax;
    if (s->Dprime < prm->minDprime || s->Dprime > prm->maxDprime) return 0;
    double both = ldcore_fisher((int)c0, (int)c4, (int)c1, (int)c5, 0, 0); /* :1222-1226 */
    if (both > prm->minP) return 0;
    s->P = both;
    s->R = sqrt(s->R2);
    s->cnt[0] = (double)c0; s->cnt[1] = (double)c1; s->cnt[2] = (double)c4; s->cnt[3] = (double)c5; /* :1239-1242 */
    s->flags = variant_flags(a, b) | 1u; /* bit0: phased math */
    if (c0 < 1 || c4 < 1 || c1 < 1 || c5 < 1) s->flags |= 1u << 3;
    if (s->R2 > 0.99) s->flags |= 1u << 4;
    s->chi_model = 0;
    s->chi_fisher = T * s->R2; /* :1259 */
    return 1;
}

/* ---------------------------------------------------------------- unphased math
 * 3x3 genotype table t[gA][gB], g in {0: 0/0, 1: het, 2: 1/1}; T = sum.
 * lib/ld/ld_engine.cpp:1562-1588 */
static double chisq_unphased(const uint64_t t[3][3], uint64_t T, double target, double p, double q) {
    const double f12 = p - target;
    const double f21 = q - target;
    const double f22 = 1 - (target + f12 + f21);
    const double e1111 = T * pow(target, 2);
    const double e1112 = 2 * T * target * f12;
    const double e1122 = T * pow(f12, 2);
    const double e1211 = 2 * T * target * f21;
    const double e1212 = 2 * T * f12 * f21 + 2 * T * target * f22;
    const double e1222 = 2 * T * f12 * f22;
    const double e2211 = T * pow(f21, 2);
    const double e2212 = 2 * T * f21 * f22;
    const double e2222 = T * pow(f22, 2);
    const double x1111 = e1111 > 0 ? pow((double)t[0][0] - e1111, 2) / e1111 : 0;
    const double x1112 = e1112 > 0 ? pow((double)t[0][1] - e1112, 2) / e1112 : 0;
    const double x1122 = e1122 > 0 ? pow((double)t[0][2] - e1122, 2) / e1122 : 0;
    const double x1211 = e1211 > 0 ? pow((double)t[1][0] - e1211, 2) / e1211 : 0;
    const double x1212 = e1212 > 0 ? pow((double)t[1][1] - e1212, 2) / e1212 : 0;
    const double x1222 = e1222 > 0 ? pow((double)t[1][2] - e1222, 2) / e1222 : 0;
    const double x2211 = e2211 > 0 ? pow((double)t[2][0] - e2211, 2) / e2211 : 0;
    const double x2212 = e2212 > 0 ? pow((double)t[2][1] - e2212, 2) / e2212 : 0;
    const double x2222 = e2222 > 0 ? pow((double)t[2][2] - e2222, 2) / e2222 : 0;
    return x1111 + x1112 + x1122 + x1211 + x1212 + x1222 + x2211 + x2212 + x2222;
}

/* lib/ld/ld_engine.cpp:1590-1684 */
static int choose_f11(uint64_t T, double target, double p, double q, uint16_t flags, const ld_params* prm,
                      const ld_variant* a, const ld_variant* b, ld_stats* s) {
    double f11 = target, f12 = p - f11, f21 = q - f11;
    double f22 = 1 - (f11 + f12 + f21);
    double D = (f11 * f22) - (f12 * f21);
    s->D = D;
    s->R2 = (D * D) / (p * (1 - p) * q * (1 - q));
    if (s->R2 < prm->minR2 || s->R2 > prm->maxR2) return 0;
    s->R = sqrt(s->R2);
    /* :1624-1627 -- slot REFALT(2) <- f12, slot ALTREF(1) <- f21 */
    s->cnt[0] = f11 * 2 * T;
    s->cnt[2] = f12 * 2 * T;
    s->cnt[1] = f21 * 2 * T;
    s->cnt[3] = f22 * 2 * T;
    if (s->cnt[0] < s->cnt[3]) { /* :1631-1643 */
        if (s->cnt[2] + s->cnt[1] + s->cnt[0] < 5) return 0;
    } else {
        if (s->cnt[3] + s->cnt[2] + s->cnt[1] < 5) return 0;
    }
    double dmax; /* :1645-1648 */
    if (s->D >= 0) dmax = p * (1.0 - q) < q * (1.0 - p) ? p * (1.0 - q) : q * (1.0 - p);
    else dmax = p * q < (1 - p) * (1 - q) ? -p * q : -(1 - p) * (1 - q);
    s->Dprime = s->D / dmax;
    if (s->Dprime < prm->minDprime || s->Dprime > prm->maxDprime) return 0;
    s->P = ldcore_fisher((int)round(s->cnt[0]), (int)round(s->cnt[2]), (int)round(s->cnt[1]), (int)round(s->cnt[3]), 0, 0);
    if (s->P > prm->minP) return 0;
    s->chi_model = 0; /* :1670 */
    s->chi_fisher = (s->cnt[0] + s->cnt[2] + s->cnt[1] + s->cnt[3]) * s->R2;
    s->flags = flags | variant_flags(a, b); /* bit0 NOT set */
    if (s->cnt[0] < 1 || s->cnt[2] < 1 || s->cnt[1] < 1 || s->cnt[3] < 1) s->flags |= 1u << 3;
    if (s->R2 > 0.99) s->flags |= 1u << 4;
    return 1;
}

static int in_range(double x, double lo, double hi) { return x >= lo - LD_ROUNDING_SLACK && x <= hi + LD_ROUNDING_SLACK; }

/* lib/ld/ld_engine.cpp:1312-1560 */
int ldcore_unphased_stats(const uint64_t t[3][3], const ld_params* prm, const ld_variant* a, const ld_variant* b,
                          ld_stats* s) {
    uint64_t T = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T += t[i][j];
    if (T < LD_MIN_ALLELES) return 0;
    const uint64_t hets = t[1][1];
    if (hets == 0) { /* :1334-1348 -- no phase uncertainty: haplotype counts + phased math */
        uint64_t c0 = 2 * t[0][0] + t[0][1] + t[1][0];
        uint64_t c4 = 2 * t[0][2] + t[0][1] + t[1][2];
        uint64_t c1 = 2 * t[2][0] + t[1][0] + t[2][1];
        uint64_t c5 = 2 * t[2][2] + t[2][1] + t[1][2];
        return ldcore_phased_stats(c0, c1, c4, c5, prm, a, b, s);
    }
    /* :1363-1375 */
    const double P = ((t[0][0] + t[0][1] + t[0][2]) * 2.0 + (t[1][0] + t[1][1] + t[1][2])) / (2.0 * T);
    const double Q = ((t[0][0] + t[1][0] + t[2][0]) * 2.0 + (t[0][1] + t[1][1] + t[2][1])) / (2.0 * T);
    const double n11 = (2.0 * t[0][0] + t[0][1] + t[1][0]);
    const double minhap = n11 / (2.0 * T);
    const double maxhap = (n11 + hets) / (2.0 * T);
    const double dee = -n11 * P * Q;
    const double c = -n11 * (1.0 - 2.0 * P - 2.0 * Q) - hets * (1.0 - P - Q) + (2.0 * T * P * Q);
    const double bb = 2.0 * T * (1.0 - 2.0 * P - 2.0 * Q) - 2.0 * n11 - hets;
    const double aa = 4.0 * T;
    /* :1388-1392 */
    const double xN = -bb / (3.0 * aa);
    const double d2 = (pow(bb, 2) - 3.0 * aa * c) / (9 * pow(aa, 2));
    const double yN = aa * pow(xN, 3) + bb * pow(xN, 2) + c * xN + dee;
    const double yN2 = pow(yN, 2);
    const double h2 = 4 * pow(aa, 2) * pow(d2, 3);
    const double diff = yN2 - h2;
    uint16_t flags = 0;
    if (diff < 0) { /* :1438-1496 three real roots */
        double h = pow(h2, 0.5);
        double theta = ((acos(-yN / h)) / 3.0);
        double delta = pow(d2, 0.5);
        double alpha = xN + 2.0 * delta * cos(theta);
        double beta = xN + 2.0 * delta * cos(2.0 * M_PI / 3.0 + theta);
        double gamma = xN + 2.0 * delta * cos(4.0 * M_PI / 3.0 + theta);
        int possible = 0;
        double best = 1.7976931348623157e308, chosen = alpha;
        if (in_range(alpha, minhap, maxhap)) { ++possible; best = chisq_unphased(t, T, alpha, P, Q); }
        if (in_range(beta, minhap, maxhap)) {
            ++possible;
            double x = chisq_unphased(t, T, beta, P, Q);
            if (x < best) { chosen = beta; best = x; }
        }
        if (in_range(gamma, minhap, maxhap)) {
            ++possible;
            double x = chisq_unphased(t, T, gamma, P, Q);
            if (x < best) { chosen = gamma; best = x; }
        }
        if (possible == 0) return 0;
        if (possible > 1) flags |= 1u << 5;
        return choose_f11(T, chosen, P, Q, flags, prm, a, b, s);
    } else if (diff > 0) { /* :1498-1519 one real root (Cardano) */
        double n1, n2;
        if ((1.0 / (2.0 * aa) * (-yN + pow((yN2 - h2), 0.5))) < 0) n1 = -pow(-(1.0 / (2.0 * aa) * (-yN + pow((yN2 - h2), 0.5))), 1.0 / 3.0);
        else n1 = pow((1.0 / (2.0 * aa) * (-yN + pow((yN2 - h2), 0.5))), 1.0 / 3.0);
        if ((1.0 / (2.0 * aa) * (-yN - pow((yN2 - h2), 0.5))) < 0) n2 = -pow(-(1.0 / (2.0 * aa) * (-yN - pow((yN2 - h2), 0.5))), 1.0 / 3.0);
        else n2 = pow((1.0 / (2.0 * aa) * (-yN - pow((yN2 - h2), 0.5))), 1.0 / 3.0);
        double alpha = xN + n1 + n2;
        if (!in_range(alpha, minhap, maxhap)) return 0;
        return choose_f11(T, alpha, P, Q, flags, prm, a, b, s);
    } else { /* :1521-1558 repeated root */
        const double delta = pow((yN / 2.0 * aa), (1.0 / 3.0));
        const double alpha = xN + delta;
        const double gamma = xN - 2.0 * delta;
        if (isnan(alpha) || isnan(gamma)) return 0;
        int possible = 0;
        double best = 1.7976931348623157e308, chosen = alpha;
        if (in_range(alpha, minhap, maxhap)) { ++possible; best = chisq_unphased(t, T, alpha, P, Q); }
        if (in_range(gamma, minhap, maxhap)) {
            ++possible;
            double x = chisq_unphased(t, T, gamma, P, Q);
            if (x < best) { chosen = gamma; best = x; }
        }
        if (possible == 0) return 0;
        return choose_f11(T, chosen, P, Q, flags, prm, a, b, s);
    }
}

/* ------------------------------------------------------------------- counting */
static inline uint64_t pc(uint64_t x) { return (uint64_t)__builtin_popcountll(x); }

/* phased, no missing: lib/ld/ld_engine.cpp:230-246 / :668-685.
 * n11 = popcount(A & B); the other cells follow from ac and 2N. */
void ldcore_count_phased_nomiss(const uint64_t* A, const uint64_t* B, uint32_t words, uint32_t n_samples,
                                uint32_t acA, uint32_t acB, uint64_t out[4] /* c0,c1,c4,c5 */) {
    uint64_t n11 = 0;
    for (uint32_t k = 0; k < words; ++k) n11 += pc(A[k] & B[k]);
    out[3] = n11;
    out[1] = acA - n11;                                    /* slot ALTREF(1): A alt, B ref */
    out[2] = acB - n11;                                    /* slot REFALT(4): A ref, B alt */
    out[0] = 2 * (uint64_t)n_samples - ((acA + acB) - n11);
}

/* phased with a missing mask on either side: lib/ld/ld_engine.cpp:513-634.
 * quirks=0: the four masked haplotype counts. quirks=1 additionally reproduces
 * the scalar-tail defect Q1 (:594-609): for the words past the last full
 * 128-bit register, REFREF accumulates popc(A ref & B alt), the two mixed cells
 * swap slots, and (pad/2) is subtracted from REFREF, all in uint64. */
void ldcore_count_phased_masked(const uint64_t* A, const uint64_t* mA, const uint64_t* B, const uint64_t* mB,
                                uint32_t n_samples, int quirks, uint64_t out[4]) {
    const uint32_t byte_width = (2 * n_samples + 63) / 64;          /* ld_engine.cpp:58 */
    const uint32_t aligned_end = (2 * n_samples / 128) * 2;         /* :59-60 */
    uint64_t rr = 0, ar = 0, ra = 0, aa = 0; /* ar: A alt/B ref (slot 1); ra: A ref/B alt (slot 4) */
    const uint32_t body_end = quirks ? aligned_end : byte_width;
    for (uint32_t k = 0; k < body_end; ++k) {
        uint64_t v = ~((mA ? mA[k] : 0) | (mB ? mB[k] : 0));
        aa += pc(A[k] & B[k] & v);
        rr += pc(~A[k] & ~B[k] & v);
        ar += pc(A[k] & ~B[k] & v);
        ra += pc(~A[k] & B[k] & v);
    }
    if (!quirks) {
        rr -= (uint64_t)byte_width * 64 - 2 * (uint64_t)n_samples; /* padding bits read as ref/ref */
    } else {
        for (uint32_t k = aligned_end; k < byte_width; ++k) {
            uint64_t v = ~((mA ? mA[k] : 0) | (mB ? mB[k] : 0));
            uint64_t t_ra = pc(~A[k] & B[k] & v), t_ar = pc(A[k] & ~B[k] & v);
            rr += t_ra;   /* :600 */
            ar += t_ra;   /* :601,607 -> slot 1 */
            ra += t_ar;   /* :602,606 -> slot 4 */
            aa += pc(A[k] & B[k] & v);
        }
        rr -= ((uint64_t)byte_width * 64 - 2 * (uint64_t)n_samples) / 2; /* :61,609 */
    }
    out[0] = rr; out[1] = ar; out[2] = ra; out[3] = aa;
}

/* unphased 3x3 genotype table over samples present in both variants:
 * lib/ld/ld_engine.cpp:709-866 (masked) / :868-1009 / :1093-1160 (run-length);
 * all three give this table. Bits 2s, 2s+1 = sample s. */
void ldcore_count_unphased(const uint64_t* A, const uint64_t* mA, const uint64_t* B, const uint64_t* mB,
                           uint32_t n_samples, uint64_t t[3][3]) {
    const uint64_t LO = 0x5555555555555555ull;
    const uint32_t words = (2 * n_samples + 63) / 64;
    memset(t, 0, 9 * sizeof(uint64_t));
    for (uint32_t k = 0; k < words; ++k) {
        uint64_t live = LO;
        if (k == words - 1 && (2 * n_samples) % 64) live &= (1ull << ((2 * n_samples) % 64)) - 1;
        uint64_t ma = mA ? mA[k] : 0, mb = mB ? mB[k] : 0;
        uint64_t v = live & ~((ma | (ma >> 1) | mb | (mb >> 1)) & LO);
        uint64_t a0 = A[k] & LO, a1 = (A[k] >> 1) & LO, b0 = B[k] & LO, b1 = (B[k] >> 1) & LO;
        uint64_t gA[3], gB[3];
        gA[1] = (a0 ^ a1) & v; gA[2] = a0 & a1 & v; gA[0] = v & ~(a0 | a1);
        gB[1] = (b0 ^ b1) & v; gB[2] = b0 & b1 & v; gB[0] = v & ~(b0 | b1);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[i][j] += pc(gA[i] & gB[j]);
    }
}

/* One pair through the comparator the reference would pick (twk_ld_slave::{Phased,Unphased,Calculate*}:
 * -u -> 3x3 table; -p -> 2x2; neither -> unphased iff either variant has missing alleles, :2775) and the math. */
static int pair_stats(const uint64_t* data, const uint64_t* mask, size_t stride, uint32_t n_samples, uint32_t words,
                      uint32_t thresh_miss_p, const ld_variant* meta, uint32_t i, uint32_t j, const ld_params* prm, ld_stats* s) {
    const ld_variant *a = &meta[i], *b = &meta[j];
    const uint64_t *A = data + (size_t)i * stride, *B = data + (size_t)j * stride;
    const uint64_t* mA = (mask && a->gt_missing) ? mask + (size_t)i * stride : 0;
    const uint64_t* mB = (mask && b->gt_missing) ? mask + (size_t)j * stride : 0;
    int unphased = prm->forced_unphased;
    if (!prm->force_phased && !prm->forced_unphased) unphased = (a->an || b->an); /* :2775 */
    if (unphased) {
        uint64_t t[3][3];
        ldcore_count_unphased(A, mA, B, mB, n_samples, t);
        return ldcore_unphased_stats(t, prm, a, b, s);
    }
    uint64_t c[4];
    if (!a->gt_missing && !b->gt_missing) {
        ldcore_count_phased_nomiss(A, B, words, n_samples, a->ac, b->ac, c);
    } else if (prm->emulate_quirks && (a->ac + b->ac < thresh_miss_p || (prm->force_phased && prm->bitmaps))) {
        /* run-length comparator (:1011-1091): same counts, mixed cells in the
         * opposite slots (Q3) */
        ldcore_count_phased_masked(A, mA, B, mB, n_samples, 0, c);
        uint64_t tmp = c[1]; c[1] = c[2]; c[2] = tmp;
    } else {
        ldcore_count_phased_masked(A, mA, B, mB, n_samples, prm->emulate_quirks, c);
    }
    return ldcore_phased_stats(c[0], c[1], c[2], c[3], prm, a, b, s);
}

/* ------------------------------------------------------------------ enumeration
 * Pair loop of twk_ld_slave::{Phased,Unphased,Calculate*} over the block-pair
 * grid of twk_ld_dynamic_balancer (lib/ld/ld_engine.cpp:1898-2838,
 * lib/ld/ld_balancing.h:176-233), n_chunks = 1. Emits forward records only
 * (the reference also writes each record with (rid,pos) swapped, :1290-1298).
 * Returns the number of records, or -1 if out_cap was too small. */
int64_t ldcore_calc(const uint64_t* data, const uint64_t* mask, size_t stride, uint32_t n_samples, uint32_t n_variants,
                    const ld_variant* meta, const ld_params* prm, uint8_t* out, int64_t out_cap,
                    uint64_t* pairs_visited) {
    const uint32_t bs = prm->block_size > 0 ? (uint32_t)prm->block_size : 500;
    /* block starts: <= bs variants, one contig per block (lib/importer.cpp:196-236) */
    uint32_t* bstart = (uint32_t*)malloc(sizeof(uint32_t) * (n_variants + 2));
    uint32_t nb = 0;
    for (uint32_t v = 0; v < n_variants;) {
        bstart[nb++] = v;
        uint32_t e = v + 1;
        while (e < n_variants && e - v < bs && meta[e].rid == meta[v].rid) ++e;
        v = e;
    }
    bstart[nb] = n_variants;
    const uint32_t words = (2 * n_samples + 63) / 64;
    const uint32_t thresh_miss_p = (uint32_t)(0.0047 * n_samples + 5.2913); /* ld_engine.cpp:1910 */
    int64_t n_out = 0;
    uint64_t visited = 0;
    int overflow = 0;
    for (uint32_t bi = 0; bi < nb && !overflow; ++bi) {
        for (uint32_t bj = bi; bj < nb && !overflow; ++bj) {
            if (prm->window && bi != bj) { /* ld_balancing.h:189-196: prune the rest of the row */
                if (meta[bstart[bj]].pos - meta[bstart[bi + 1] - 1].pos > (uint32_t)prm->l_window) break;
            }
            const uint32_t i0 = bstart[bi], i1 = bstart[bi + 1], j0 = bstart[bj], j1 = bstart[bj + 1];
            int aborted = 0;
            for (uint32_t i = i0; i < i1 && !aborted; ++i) {
                for (uint32_t j = (bi == bj ? i + 1 : j0); j < j1; ++j) {
                    const ld_variant *a = &meta[i], *b = &meta[j];
                    /* Per-pair window rule. -p / -u (CalculatePhasedWindow :2553-2560, CalculateUnphasedWindow
                     * :2658-2664): the first out-of-window pair abandons the whole block pair. -p -m -M
                     * (CalculatePhasedBitmapWindow :2466-2471, :2490-2495) skips a pair only when the contigs
                     * DIFFER and the (wrapping) position difference exceeds the window. Auto mode
                     * (twk_ld_slave::Calculate :2737-2838) has no per-pair rule at all: only the balancer's
                     * row prune above applies. */
                    if (prm->window && prm->emulate_quirks && prm->force_phased && prm->bitmaps) {
                        if (a->rid != b->rid && (b->pos - a->pos) > (uint32_t)prm->l_window) continue;
                    } else if (prm->window && (!prm->emulate_quirks || prm->force_phased || prm->forced_unphased)) {
                        if (a->rid == b->rid && (b->pos - a->pos) > (uint32_t)prm->l_window) {
                            aborted = 1;
                            break;
                        }
                    }
                    if (a->ac + b->ac <= 2) continue; /* :1918 */
                    ld_stats s;
                    const int pass = pair_stats(data, mask, stride, n_samples, words, thresh_miss_p, meta, i, j, prm, &s);
                    if (pass) {
                        if (n_out >= out_cap) { overflow = 1; aborted = 1; break; }
                        put_record(out + (size_t)n_out * LD_RECORD_BYTES, &s, a, b);
                        ++n_out;
                    }
                }
            }
            if (!aborted) { /* progress->n_var, :1933, :2015; skipped for aborted window tiles (:2607) */
                uint64_t ni = i1 - i0, nj = j1 - j0;
                visited += (bi == bj) ? (ni * ni - ni) / 2 : ni * nj;
            }
        }
    }
    free(bstart);
    if (pairs_visited) *pairs_visited = visited;
    return overflow ? -1 : n_out;
}


/* `scalc` = twk_ld_slave::CalculateSingle (lib/ld/ld_engine.cpp:2226-2332) over the blocks of
 * twk_ld_impl::LoadTargetSingle (lib/ld/ld.cpp:123-255): the first n_targets variants are the target site(s); every
 * pair (target, later target) and (target, other) goes through the comparator of AUTO mode whatever -p / -u say
 * (twk_ld_slave::Start, :1826-1830) and WITHOUT the ac_i + ac_j <= 2 skip (commented out at :2265, :2290). The target
 * is variant A of every record. */
int64_t ldcore_calc_single(const uint64_t* data, const uint64_t* mask, size_t stride, uint32_t n_samples, uint32_t n_variants,
                           const ld_variant* meta, const ld_params* prm_in, uint32_t n_targets, uint8_t* out, int64_t out_cap,
                           uint64_t* pairs_visited) {
    ld_params prm = *prm_in;
    prm.force_phased = prm.forced_unphased = 0;
    prm.bitmaps = 0;
    const uint32_t words = (2 * n_samples + 63) / 64;
    const uint32_t thresh_miss_p = (uint32_t)(0.0047 * n_samples + 5.2913);
    int64_t n_out = 0;
    for (uint32_t i = 0; i < n_targets && i < n_variants; ++i)
        for (uint32_t j = i + 1; j < n_variants; ++j) {
            ld_stats s;
            if (!pair_stats(data, mask, stride, n_samples, words, thresh_miss_p, meta, i, j, &prm, &s)) continue;
            if (n_out >= out_cap) return -1;
            put_record(out + (size_t)n_out * LD_RECORD_BYTES, &s, &meta[i], &meta[j]);
            ++n_out;
        }
    if (pairs_visited) {
        const uint64_t t = n_targets, m = n_variants;
        *pairs_visited = (t * t - t) / 2 + t * (m - t);
    }
    return n_out;
}

int ldcore_record_bytes(void) { return LD_RECORD_BYTES; }
