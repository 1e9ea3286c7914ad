// stats.cuh -- the statistics stage of `tomahawk calc` on the device: D, D', R,
// R2, Fisher's exact P, chi-squared, flags, filters (in the reference's order)
// and the packed 106-byte record. One thread per surviving pair, pairs re-ordered inside a block so
// that the lanes of a warp walk Fisher tails of similar length.
//
// Reference: twk_ld_engine::PhasedMath lib/ld/ld_engine.cpp:1162-1310,
// UnphasedMath :1312-1560, ChiSquaredUnphasedTable :1562-1588,
// ChooseF11Calculate :1590-1740, kt_fisher_exact lib/fisher_math.cpp:183-267.
//
// All phased arithmetic uses the round-to-nearest intrinsics (__dmul_rn, ...)
// so no multiply-add is ever contracted: the reference runs on x86-64 SSE2
// doubles without FMA, and with the same operation order the device results
// are bit-identical. Fisher's P is NOT computed in the reference's operation order
// (see fisher_tail): its log-factorials come from a host-built table of glibc
// lgamma(n+1) (bit-identical to the reference's lgamma calls), the closed-form
// points use CUDA's exp() (<= 1 ulp from glibc's) and the recurrence between them is
// evaluated division-free; P agrees with the reference to ~1e-14 relative (tests: 1e-9).
#pragma once
#include "common.cuh"

namespace twkb {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ------------------------------------------------------------------ Fisher
struct LgTable {
    const double* lg;  // lg[n] = lgamma(n + 1), n in [0, len)
    uint32_t len;
    __device__ __forceinline__ double at(int n) const {
        if (n >= 0 && (uint32_t)n < len) return __ldg(lg + n);
        return lgamma((double)n + 1.0);
    }
};

// fisher_math.cpp:183-187
__device__ __forceinline__ double log_binom(const LgTable& t, int n, int k) {
    if (k == 0 || n == k) return 0.0;
    return dsub(dsub(t.at(n), t.at(k)), t.at(n - k));
}
// :195-198
__device__ __noinline__ double hyper_pmf(const LgTable& t, int n11, int n1_, int n_1, int n) {
    return exp(dsub(dadd(log_binom(t, n1_, n11), log_binom(t, n - n1_, n_1 - n11)), log_binom(t, n, n_1)));
}

// One tail of kt_fisher_exact's walk (fisher_math.cpp:240-256), DIR = +1 from the lower end of the support,
// DIR = -1 from the upper end: starting with p = pmf(cur), add p to `sum` and move one table further while
// p < qlo and the next table exists; on return p is the first probability that stopped the walk.
//
// The reference moves by the recurrence of hypergeo_acc (:211-227)
//     p(k+1) = p(k) * ((n1. - k) / (k + 1) * (n.1 - k)) / (k + 1 + d),      d = n - n1. - n.1
// (mirrored downwards) and re-evaluates the closed form at every k that is a multiple of 11 (or where the
// fourth cell is 0). Between two such points there are at most 10 recurrence steps. Here such a run is done
// without a single division inside: with n_j, d_j the integer numerator / denominator of step j (products of
// two counts, exact in a double for n < 2^26),
//     A_j = n_1 ... n_j,  B_j = d_1 ... d_j,  p_j = p A_j / B_j,  p_0 + ... + p_(j-1) = p W_j / B_(j-1),
//     W_(j+1) = W_j d_j + A_j   (W_0 = 0, d_0 = 1)
// so a step costs 2 integer-valued products, 3 multiplications and one FMA; the stop rule p_j < qlo is
// tested as p A_j < qlo B_j; the run ends with two divisions. Every p_j agrees with the reference's to a few
// ulp (its own three roundings per step are replaced by ~3 per step of the same size), far inside the 1e-9
// the tests hold P to, and the 10 steps are straight-line code: the previous one-step-at-a-time loop with its
// two correctly rounded divisions per step left the kernel stalled on instruction fetch (ncu, C1: "no
// instruction" 8.3 warps per issue, fp64 pipe 9 % busy; profiles/round2_ncu_stats_c1_before.csv).
// A run that starts from a DENORMAL probability, in the reference's own operation order (hypergeo_acc,
// fisher_math.cpp:213-225): there every product is rounded to the denormal grid, and when the ratios are large
// (far tails of strongly associated pairs) that rounding is carried up into terms that matter at the 1e-9
// level of a P ~ 1e-288. Rare, so it lives out of line to keep the hot loop small.
template <int DIR>
__device__ __noinline__ int fisher_run_denormal(int cur, double& p, int steps, const double qlo, const int n1_, const int n_1, const int n,
                                               double& sum) {
    const int d = n - n1_ - n_1;
    int m = 0;
    for (; m < steps && p < qlo; ++m) {
        sum = dadd(sum, p);
        const int c = cur + DIR * m, x = c + DIR;
        const double f = DIR > 0 ? ddiv(dmul(ddiv((double)(n1_ - c), (double)x), (double)(n_1 - c)), (double)(x + d))
                                 : ddiv(dmul(ddiv((double)c, (double)(n1_ - x)), (double)(c + d)), (double)(n_1 - x));
        p = dmul(p, f);
    }
    return m;
}

template <int DIR>
__device__ __forceinline__ void fisher_tail(const LgTable& t, int& cur, double& p, const int lim, const double qlo, const int n1_,
                                            const int n_1, const int n, double& sum) {
    const int d = n - n1_ - n_1;
    for (;;) {
        const int nxt = cur + DIR;
        if (!(p < qlo) || (DIR > 0 ? nxt > lim : nxt < lim)) return;
        // recurrence steps before the next closed-form point (a multiple of 11, or the table whose fourth cell is 0)
        int steps;
        if (DIR > 0) {
            const int r = nxt % 11;
            steps = r ? 11 - r : 0;
            steps = min(steps, lim - nxt + 1);
        } else {
            steps = nxt % 11;
            const int z = nxt + d;
            if (z >= 0 && z < steps) steps = z;
            steps = min(steps, nxt - lim + 1);
        }
        if (steps > 0 && p > 0.0 && p < 2.2250738585072014e-308) {
            const int m = fisher_run_denormal<DIR>(cur, p, steps, qlo, n1_, n_1, n, sum);
            cur += DIR * m;
            if (m < steps) return;
            continue;
        }
        if (steps > 0) {
            double A = 1.0, B = 1.0, Bprev = 1.0, W = 0.0, dlast = 1.0;
            int m = 0;
            // factors of the step from table c: up   n = (n1. - c)(n.1 - c),  den = (c + 1)(c + 1 + d)
            //                                   down n = c (c + d),           den = (n1. - c + 1)(n.1 - c + 1)
            double fa = DIR > 0 ? (double)(n1_ - cur) : (double)cur;
            double fb = DIR > 0 ? (double)(n_1 - cur) : (double)(cur + d);
            double fc = DIR > 0 ? (double)(cur + 1) : (double)(n1_ - cur + 1);
            double fd = DIR > 0 ? (double)(cur + 1 + d) : (double)(n_1 - cur + 1);
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                if (k < steps && p * A < qlo * B) {  // p_k < qlo: p_k joins the sum and the walk moves on
                    const double nk = fa * fb, dk = fc * fd;
                    W = fma(W, dlast, A);
                    A *= nk;
                    Bprev = B;
                    B *= dk;
                    dlast = dk;
                    fa -= 1.0; fb -= 1.0; fc += 1.0; fd += 1.0;
                    ++m;
                }
            }
            if (m > 0) {
                sum += p * (W / Bprev);
                p = p * (A / B);
                cur += DIR * m;
            }
            if (m < steps) return;  // stopped inside the run: p >= qlo
            continue;               // re-test the stop rule, then the closed-form point
        }
        // closed-form point (fisher_math.cpp:211, :226-228)
        sum += p;
        p = hyper_pmf(t, nxt, n1_, n_1, n);
        cur = nxt;
    }
}

// :231-267, two-sided P only.
//
// The reference walks the whole support from both ends, one recurrence step per table, until the
// probability reaches that of the observed table (q): ~min(n1_, n_1) steps, almost all of them
// over terms that are hundreds of orders of magnitude below q. Its recurrence is re-seeded from
// the closed form at every n11 that is a multiple of 11 (:211), so a walk may START at any such
// point. Here each tail starts at the multiple of 11 closest to the mode whose probability is still
// below FISHER_SKIP * q (found by bisection on the closed form; the pmf is monotone on either side of
// the mode): the skipped terms sum to < 2^20 * 1e-18 * q = 1e-12 * q even at 1 M haplotypes, three
// orders of magnitude inside the 1e-9 agreement the tests hold P to (P >= ~q), and the walk shrinks
// from the support to ~18 standard deviations.
constexpr double FISHER_SKIP = 1e-18;
constexpr int FISHER_SKIP_MIN_RANGE = 48;  // shorter supports are walked whole

__device__ double fisher_two_sided(const LgTable& t, int n11, int n12, int n21, int n22) {
    const int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
    const int max = (n_1 < n1_) ? n_1 : n1_;
    int min = n1_ + n_1 - n;
    if (min < 0) min = 0;
    if (min == max) return 1.0;
    const double q = hyper_pmf(t, n11, n1_, n_1, n);
    const double qlo = dmul(0.99999999, q), qhi = dmul(1.00000001, q);
    const bool may_skip = max - min > FISHER_SKIP_MIN_RANGE;
    const double cut = FISHER_SKIP * q;
    const int mode = (int)(((long long)(n1_ + 1) * (long long)(n_1 + 1)) / ((long long)n + 2));
    // ---- left tail: from min upwards
    double p = hyper_pmf(t, min, n1_, n_1, n);
    int cur = min;
    if (may_skip && p < cut) {
        // largest multiple of 11 in (min, mode] whose probability is below the cut
        int lo = min / 11 + 1, hi = mode / 11, best = -1;
        double best_p = 0.0;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const double pm = hyper_pmf(t, 11 * mid, n1_, n_1, n);
            if (pm < cut) { best = mid; best_p = pm; lo = mid + 1; }
            else hi = mid - 1;
        }
        if (best >= 0) { cur = 11 * best; p = best_p; }
    }
    double left = 0.0;
    fisher_tail<1>(t, cur, p, max, qlo, n1_, n_1, n, left);
    if (p < qhi) left += p;
    // ---- right tail: from max downwards
    p = hyper_pmf(t, max, n1_, n_1, n);
    cur = max;
    if (may_skip && p < cut) {
        // smallest multiple of 11 in (mode, max) whose probability is below the cut
        int lo = mode / 11 + 1, hi = (max - 1) / 11, best = -1;
        double best_p = 0.0;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const double pm = hyper_pmf(t, 11 * mid, n1_, n_1, n);
            if (pm < cut) { best = mid; best_p = pm; hi = mid - 1; }
            else lo = mid + 1;
        }
        if (best >= 0) { cur = 11 * best; p = best_p; }
    }
    double right = 0.0;
    fisher_tail<-1>(t, cur, p, 0, qlo, n1_, n_1, n, right);
    if (p < qhi) right += p;
    double two = left + right;
    if (two > 1.0) two = 1.0;
    return two;
}

// ---------------------------------------------------------------- the record
struct PairStats {
    uint32_t flags;
    double cnt[4], D, Dprime, R, R2, P, chi_fisher, chi_model;
};

__device__ __forceinline__ void st_u16(uint8_t* p, uint32_t v) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)v; }
__device__ __forceinline__ void st_u32_2(uint8_t* p, uint32_t v) {
    st_u16(p, v & 0xffffu);
    st_u16(p + 2, v >> 16);
}
__device__ __forceinline__ void st_f64_2(uint8_t* p, double d) {
    unsigned long long v = (unsigned long long)__double_as_longlong(d);
    st_u16(p, (uint32_t)(v & 0xffffu));
    st_u16(p + 2, (uint32_t)((v >> 16) & 0xffffu));
    st_u16(p + 4, (uint32_t)((v >> 32) & 0xffffu));
    st_u16(p + 6, (uint32_t)(v >> 48));
}
// lib/core.cpp:470-490. Records are only 2-byte aligned (106 = 2*53).
__device__ __forceinline__ void write_record(uint8_t* dst, const PairStats& s, const DevVariant& a, const DevVariant& b) {
    st_u16(dst, s.flags);
    st_u32_2(dst + 2, a.rid);
    st_u32_2(dst + 6, b.rid);
    st_u32_2(dst + 10, a.pos << 2);
    st_u32_2(dst + 14, b.pos << 2);
    st_f64_2(dst + 18, s.cnt[0]);
    st_f64_2(dst + 26, s.cnt[1]);
    st_f64_2(dst + 34, s.cnt[2]);
    st_f64_2(dst + 42, s.cnt[3]);
    st_f64_2(dst + 50, s.D);
    st_f64_2(dst + 58, s.Dprime);
    st_f64_2(dst + 66, s.R);
    st_f64_2(dst + 74, s.R2);
    st_f64_2(dst + 82, s.P);
    st_f64_2(dst + 90, s.chi_fisher);
    st_f64_2(dst + 98, s.chi_model);
}

// ld_engine.cpp:1244-1255 / :1674-1684
__device__ __forceinline__ uint32_t variant_flags(const DevVariant& a, const DevVariant& b) {
    uint32_t f = 0;
    const bool same = a.rid == b.rid;
    int diff = (int)a.pos - (int)b.pos;
    if (diff < 0) diff = -diff;
    if (same) f |= 1u << 1;
    if ((double)diff > 500e3 && same) f |= 1u << 2;
    if (a.flags & VF_HAS_MISSING) f |= 1u << 8;
    if (b.flags & VF_HAS_MISSING) f |= 1u << 9;
    if (a.ac < 5) f |= 1u << 10;
    if (b.ac < 5) f |= 1u << 11;
    if (a.flags & VF_BAD_HWE) f |= 1u << 12;
    if (b.flags & VF_BAD_HWE) f |= 1u << 13;
    return f;
}

// ----------------------------------------------------------------- phased math
// ld_engine.cpp:1162-1259. c0=REFREF, c1=slot 1, c4=slot 4, c5=ALTALT.
__device__ bool phased_stats(unsigned long long c0, unsigned long long c1, unsigned long long c4,
                             unsigned long long c5, const DevParams& prm, const LgTable& lg, const DevVariant& a,
                             const DevVariant& b, PairStats& s) {
    const unsigned long long T = c0 + c4 + c1 + c5;
    if (T < 5) return false;
    if (c0 < c5) {
        if (c4 + c1 + c0 < 5) return false;
    } else {
        if (c5 + c4 + c1 < 5) return false;
    }
    const double Td = (double)T;
    const double d0 = (double)c0, d1 = (double)c1, d4 = (double)c4, d5 = (double)c5;
    const double pA = ddiv(d0, Td), qA = ddiv(d1, Td), pB = ddiv(d4, Td), qB = ddiv(d5, Td);
    const double D = dsub(dmul(pA, qB), dmul(qA, pB));
    if (D == 0.0) return false;
    const double g0 = ddiv(dadd(d0, d4), Td);
    const double g1 = ddiv(dadd(d1, d5), Td);
    const double h0 = ddiv(dadd(d0, d1), Td);
    const double h1 = ddiv(dadd(d4, d5), Td);
    s.D = D;
    s.R2 = ddiv(dmul(D, D), dmul(dmul(dmul(g0, g1), h0), h1));
    if (s.R2 < prm.minR2 || s.R2 > prm.maxR2) return false;
    double dmax;
    if (D >= 0) {
        const double x = dmul(g0, h1), y = dmul(h0, g1);
        dmax = x < y ? x : y;
    } else {
        const double x = dmul(g0, g1), y = dmul(h0, h1);
        dmax = x < y ? -x : -y;
    }
    s.Dprime = ddiv(D, dmax);
    if (s.Dprime < prm.minDprime || s.Dprime > prm.maxDprime) return false;
    const double both = fisher_two_sided(lg, (int)c0, (int)c4, (int)c1, (int)c5);
    if (both > prm.minP) return false;
    s.P = both;
    s.R = __dsqrt_rn(s.R2);
    s.cnt[0] = d0; s.cnt[1] = d1; s.cnt[2] = d4; s.cnt[3] = d5;
    s.flags = variant_flags(a, b) | 1u;
    if (c0 < 1 || c4 < 1 || c1 < 1 || c5 < 1) s.flags |= 1u << 3;
    if (s.R2 > 0.99) s.flags |= 1u << 4;
    s.chi_model = 0.0;
    s.chi_fisher = dmul(Td, s.R2);
    return true;
}

// --------------------------------------------------------------- unphased math
// Transcendentals (pow with non-integer exponent, acos, cos) come from the CUDA
// math library and differ from glibc's by a few ulp: statistics of pairs that go
// through the cubic agree with the reference to ~1e-13 relative, not bit-exactly.
__device__ __forceinline__ double sq(double x) { return dmul(x, x); }  // gcc folds pow(x,2) to x*x

// ld_engine.cpp:1562-1588
__device__ double chisq_unphased(const uint32_t* t, double Td, double target, double p, double q) {
    const double f12 = dsub(p, target);
    const double f21 = dsub(q, target);
    const double f22 = dsub(1.0, dadd(dadd(target, f12), f21));
    const double T2 = dmul(2.0, Td);
    const double e[9] = {dmul(Td, sq(target)),
                         dmul(dmul(T2, target), f12),
                         dmul(Td, sq(f12)),
                         dmul(dmul(T2, target), f21),
                         dadd(dmul(dmul(T2, f12), f21), dmul(dmul(T2, target), f22)),
                         dmul(dmul(T2, f12), f22),
                         dmul(Td, sq(f21)),
                         dmul(dmul(T2, f21), f22),
                         dmul(Td, sq(f22))};
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double x = e[k] > 0 ? ddiv(sq(dsub((double)t[k], e[k])), e[k]) : 0.0;
        sum = dadd(sum, x);
    }
    return sum;
}

// ld_engine.cpp:1590-1684
__device__ bool choose_f11(double Td, double target, double p, double q, uint32_t flags, const DevParams& prm,
                           const LgTable& lg, const DevVariant& a, const DevVariant& b, PairStats& s) {
    const double f11 = target, f12 = dsub(p, f11), f21 = dsub(q, f11);
    const double f22 = dsub(1.0, dadd(dadd(f11, f12), f21));
    const double D = dsub(dmul(f11, f22), dmul(f12, f21));
    const double omp = dsub(1.0, p), omq = dsub(1.0, q);
    s.D = D;
    s.R2 = ddiv(dmul(D, D), dmul(dmul(dmul(p, omp), q), omq));
    if (s.R2 < prm.minR2 || s.R2 > prm.maxR2) return false;
    s.R = __dsqrt_rn(s.R2);
    s.cnt[0] = dmul(dmul(f11, 2.0), Td);
    s.cnt[2] = dmul(dmul(f12, 2.0), Td);
    s.cnt[1] = dmul(dmul(f21, 2.0), Td);
    s.cnt[3] = dmul(dmul(f22, 2.0), Td);
    if (s.cnt[0] < s.cnt[3]) {
        if (dadd(dadd(s.cnt[2], s.cnt[1]), s.cnt[0]) < 5) return false;
    } else {
        if (dadd(dadd(s.cnt[3], s.cnt[2]), s.cnt[1]) < 5) return false;
    }
    double dmax;
    if (D >= 0) {
        const double x = dmul(p, omq), y = dmul(q, omp);
        dmax = x < y ? x : y;
    } else {
        const double x = dmul(p, q), y = dmul(omp, omq);
        dmax = x < y ? -x : -y;
    }
    s.Dprime = ddiv(D, dmax);
    if (s.Dprime < prm.minDprime || s.Dprime > prm.maxDprime) return false;
    s.P = fisher_two_sided(lg, (int)round(s.cnt[0]), (int)round(s.cnt[2]), (int)round(s.cnt[1]), (int)round(s.cnt[3]));
    if (s.P > prm.minP) return false;
    s.chi_model = 0.0;
    s.chi_fisher = dmul(dadd(dadd(dadd(s.cnt[0], s.cnt[2]), s.cnt[1]), s.cnt[3]), s.R2);
    s.flags = flags | variant_flags(a, b);
    if (s.cnt[0] < 1 || s.cnt[2] < 1 || s.cnt[1] < 1 || s.cnt[3] < 1) s.flags |= 1u << 3;
    if (s.R2 > 0.99) s.flags |= 1u << 4;
    return true;
}

__device__ __forceinline__ bool hap_in_range(double x, double lo, double hi) {
    return x >= dsub(lo, 1e-5) && x <= dadd(hi, 1e-5);
}

// ld_engine.cpp:1312-1560. t = 3x3 genotype table, row-major t[gA*3+gB].
__device__ bool unphased_stats(const uint32_t* t, const DevParams& prm, const LgTable& lg, const DevVariant& a,
                               const DevVariant& b, PairStats& s) {
    unsigned long long T = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) T += t[k];
    if (T < 5) return false;
    const unsigned long long hets = t[4];
    if (hets == 0) {
        const unsigned long long c0 = 2ull * t[0] + t[1] + t[3];
        const unsigned long long c4 = 2ull * t[2] + t[1] + t[5];
        const unsigned long long c1 = 2ull * t[6] + t[3] + t[7];
        const unsigned long long c5 = 2ull * t[8] + t[7] + t[5];
        return phased_stats(c0, c1, c4, c5, prm, lg, a, b, s);
    }
    const double Td = (double)T, T2 = dmul(2.0, Td), hd = (double)hets;
    const double P = ddiv(dadd(dmul((double)(t[0] + t[1] + t[2]), 2.0), (double)(t[3] + t[4] + t[5])), T2);
    const double Q = ddiv(dadd(dmul((double)(t[0] + t[3] + t[6]), 2.0), (double)(t[1] + t[4] + t[7])), T2);
    const double n11 = (double)(2ull * t[0] + t[1] + t[3]);
    const double minhap = ddiv(n11, T2);
    const double maxhap = ddiv(dadd(n11, hd), T2);
    const double G = dsub(dsub(1.0, dmul(2.0, P)), dmul(2.0, Q));
    const double dee = dmul(dmul(-n11, P), Q);
    const double c = dadd(dsub(dmul(-n11, G), dmul(hd, dsub(dsub(1.0, P), Q))), dmul(dmul(T2, P), Q));
    const double bb = dsub(dsub(dmul(T2, G), dmul(2.0, n11)), hd);
    const double aa = dmul(4.0, Td);
    const double xN = ddiv(-bb, dmul(3.0, aa));
    const double d2 = ddiv(dsub(sq(bb), dmul(dmul(3.0, aa), c)), dmul(9.0, sq(aa)));
    const double yN = dadd(dadd(dadd(dmul(aa, pow(xN, 3.0)), dmul(bb, sq(xN))), dmul(c, xN)), dee);
    const double yN2 = sq(yN);
    const double h2 = dmul(dmul(4.0, sq(aa)), pow(d2, 3.0));
    const double diff = dsub(yN2, h2);
    uint32_t flags = 0;
    if (diff < 0) {
        const double h = sqrt(h2);
        const double theta = ddiv(acos(ddiv(-yN, h)), 3.0);
        const double delta = sqrt(d2);
        const double two_delta = dmul(2.0, delta);
        const double alpha = dadd(xN, dmul(two_delta, cos(theta)));
        const double beta = dadd(xN, dmul(two_delta, cos(dadd(ddiv(dmul(2.0, 3.14159265358979323846), 3.0), theta))));
        const double gamma = dadd(xN, dmul(two_delta, cos(dadd(ddiv(dmul(4.0, 3.14159265358979323846), 3.0), theta))));
        int possible = 0;
        double best = 1.7976931348623157e308, chosen = alpha;
        if (hap_in_range(alpha, minhap, maxhap)) { ++possible; best = chisq_unphased(t, Td, alpha, P, Q); }
        if (hap_in_range(beta, minhap, maxhap)) {
            ++possible;
            const double x = chisq_unphased(t, Td, beta, P, Q);
            if (x < best) { chosen = beta; best = x; }
        }
        if (hap_in_range(gamma, minhap, maxhap)) {
            ++possible;
            const double x = chisq_unphased(t, Td, gamma, P, Q);
            if (x < best) { chosen = gamma; best = x; }
        }
        if (possible == 0) return false;
        if (possible > 1) flags |= 1u << 5;
        return choose_f11(Td, chosen, P, Q, flags, prm, lg, a, b, s);
    } else if (diff > 0) {
        const double root = sqrt(dsub(yN2, h2));
        const double k = ddiv(1.0, dmul(2.0, aa));
        const double u1 = dmul(k, dadd(-yN, root));
        const double u2 = dmul(k, dsub(-yN, root));
        const double third = 1.0 / 3.0;
        const double n1 = u1 < 0 ? -pow(-u1, third) : pow(u1, third);
        const double n2 = u2 < 0 ? -pow(-u2, third) : pow(u2, third);
        const double alpha = dadd(dadd(xN, n1), n2);
        if (!hap_in_range(alpha, minhap, maxhap)) return false;
        return choose_f11(Td, alpha, P, Q, flags, prm, lg, a, b, s);
    } else {
        const double delta = pow(dmul(ddiv(yN, 2.0), aa), 1.0 / 3.0);
        const double alpha = dadd(xN, delta);
        const double gamma = dsub(xN, dmul(2.0, delta));
        if (isnan(alpha) || isnan(gamma)) return false;
        int possible = 0;
        double best = 1.7976931348623157e308, chosen = alpha;
        if (hap_in_range(alpha, minhap, maxhap)) { ++possible; best = chisq_unphased(t, Td, alpha, P, Q); }
        if (hap_in_range(gamma, minhap, maxhap)) {
            ++possible;
            const double x = chisq_unphased(t, Td, gamma, P, Q);
            if (x < best) { chosen = gamma; best = x; }
        }
        if (possible == 0) return false;
        return choose_f11(Td, chosen, P, Q, flags, prm, lg, a, b, s);
    }
}

// ---------------------------------------------------------------- stats kernel
// One thread per candidate, but NOT in buffer order. Fisher's walk is 10^1..10^3 recurrence steps
// (two fp64 divisions each) and its length follows the spread of the hypergeometric distribution of
// the pair, sigma^2 = n1. n.1 (n - n1.)(n - n.1) / (n^2 (n - 1)): in buffer order the 32 lanes of a
// warp hold pairs of one row variant with 32 unrelated column variants, so a warp runs for the
// longest of 32 walks while most lanes idle (C1, R2 >= 0, every pair reaches Fisher: 94 ms for
// 4.9e7 candidates at ~25 % lane use). Here a block takes STATS_PER_BLOCK consecutive candidates,
// buckets them by estimated walk length (64 log-spaced bins, shared-memory counting sort) and hands
// each warp 32 neighbours of that order, so the lanes of a warp finish together. Every candidate
// is still evaluated by ONE thread with the reference's operation order: results are unchanged.
constexpr int STATS_THREADS = 256;
constexpr int STATS_PER_BLOCK = 1024;
// Record staging area of a block, 32 x 106 B per warp.
constexpr uint32_t STATS_STAGE_BYTES = (STATS_THREADS / 32) * 32 * 106;

// Estimated walk length of kt_fisher_exact for a candidate, as a bin 0..63 (4 bins per octave).
__device__ __forceinline__ uint32_t fisher_cost_bin(const Candidate& cd) {
    float n1, m1, n;
    if (cd.mode == 0) {  // fisher_two_sided(c0, c4, c1, c5): n1. = c0 + c4, n.1 = c0 + c1
        n1 = (float)cd.c[0] + (float)cd.c[2];
        m1 = (float)cd.c[0] + (float)cd.c[1];
        n = n1 + (float)cd.c[1] + (float)cd.c[3];
    } else {             // 3x3 genotype table: the allele marginals of the estimated 2x2 table
        const float r0 = (float)(cd.c[0] + cd.c[1] + cd.c[2]), r1 = (float)(cd.c[3] + cd.c[4] + cd.c[5]);
        const float k0 = (float)(cd.c[0] + cd.c[3] + cd.c[6]), k1 = (float)(cd.c[1] + cd.c[4] + cd.c[7]);
        const float T = r0 + r1 + (float)(cd.c[6] + cd.c[7] + cd.c[8]);
        n = 2.0f * T;
        n1 = 2.0f * r0 + r1;
        m1 = 2.0f * k0 + k1;
    }
    if (!(n > 1.0f)) return 0u;
    const float range = fminf(fminf(n1, m1), fminf(n - n1, n - m1));  // support size - 1
    const float var = (n1 / n) * (m1 / n) * (n - n1) * ((n - m1) / (n - 1.0f));
    const float est = fminf(range, 22.0f * sqrtf(fmaxf(var, 0.0f)) + 24.0f);
    const int bin = (int)(4.0f * __log2f(1.0f + fmaxf(est, 0.0f)));
    return (uint32_t)min(max(bin, 0), 63);
}

// UNPHASED = false: every candidate carries a 2x2 table (mode 0) -- the instantiation of the phased passes,
// without the cubic solver's registers.
//
// Records leave through shared memory: a warp builds the records of its passing lanes in a staging area in output
// order and copies the contiguous run out in aligned 8-byte words (the run starts at slot0 * 106, which is only 2-byte
// aligned: up to 3 leading and 3 trailing half-words go out singly). Written straight from the lanes a record is 53
// two-byte stores at a 106-byte lane stride -- 32 sectors per store instruction; C1 (every pair a record): 29.5 -> 24.0 ms.
// (Tried and dropped: a per-block copy of the log-factorial table in shared memory, 40 KB at 2,504 samples. With three
// blocks per SM it leaves the unified L1 ~20 KB for the candidate / metadata gathers and the kernel slows to 26-28 ms;
// the table's ~225 divergent lookups per candidate already hit L1. profiles/round2_stats_c1_staging.log)
template <bool UNPHASED, int MIN_BLOCKS>
__global__ void __launch_bounds__(STATS_THREADS, MIN_BLOCKS)
stats_kernel(const Candidate* __restrict__ cands, uint32_t n_cands, const DevVariant* __restrict__ meta,
             DevParams prm, const double* __restrict__ lgamma_tab, uint8_t* __restrict__ records,
             unsigned long long rec_capacity, unsigned long long* __restrict__ rec_count) {
    __shared__ __align__(16) uint8_t s_stage_all[STATS_STAGE_BYTES];
    __shared__ uint16_t s_perm[STATS_PER_BLOCK];
    __shared__ uint8_t s_bin[STATS_PER_BLOCK];
    __shared__ uint32_t s_off[64];
    uint8_t* s_stage = s_stage_all + (threadIdx.x >> 5) * (32 * 106);  // this warp's 32 records
    const uint32_t base = blockIdx.x * (uint32_t)STATS_PER_BLOCK;
    const uint32_t cnt = min((uint32_t)STATS_PER_BLOCK, n_cands - base);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s_off[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t idx = threadIdx.x; idx < cnt; idx += STATS_THREADS) {
        const uint32_t b = fisher_cost_bin(cands[base + idx]);
        s_bin[idx] = (uint8_t)b;
        atomicAdd(&s_off[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive prefix sum of the 64 bin counts
        const uint32_t a = s_off[2 * lane], b = s_off[2 * lane + 1];
        uint32_t incl = a + b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t excl = incl - (a + b);
        s_off[2 * lane] = excl;
        s_off[2 * lane + 1] = excl + a;
    }
    __syncthreads();
    for (uint32_t idx = threadIdx.x; idx < cnt; idx += STATS_THREADS) s_perm[atomicAdd(&s_off[s_bin[idx]], 1u)] = (uint16_t)idx;
    __syncthreads();
    const LgTable lg{lgamma_tab, prm.lgamma_len};
    for (uint32_t p0 = 0; p0 < cnt; p0 += STATS_THREADS) {  // uniform trip count: every lane takes part in the ballots
        const uint32_t pos = p0 + threadIdx.x;
        bool pass = false;
        PairStats s;
        DevVariant a, b;
        if (pos < cnt) {
            const Candidate cd = cands[base + s_perm[pos]];
            a = meta[cd.i];
            b = meta[cd.j];
            if (!UNPHASED || cd.mode == 0) {
                unsigned long long c0 = cd.c[0], c1 = cd.c[1], c4 = cd.c[2], c5 = cd.c[3];
                // Q3: the reference's run-length comparator (low allele counts, missing data)
                // stores the two mixed cells in swapped slots (ld_engine.cpp:1023,1055 vs :683-684);
                // CalculatePhasedBitmap* (-p -m -M) sends every masked pair there (:2393-2397, :2477-2481).
                if (prm.emulate_quirks && ((a.flags | b.flags) & VF_GT_MISSING) && (prm.bitmap_mode || a.ac + b.ac < prm.thresh_miss_phased)) {
                    unsigned long long tmp = c1; c1 = c4; c4 = tmp;
                }
                pass = phased_stats(c0, c1, c4, c5, prm, lg, a, b, s);
            } else {
                pass = unphased_stats(cd.c, prm, lg, a, b, s);
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, pass);
        if (ballot == 0) continue;
        const int leader = __ffs(ballot) - 1;
        const uint32_t n_pass = (uint32_t)__popc(ballot);
        unsigned long long slot0 = 0;
        if (lane == leader) slot0 = atomicAdd(rec_count, (unsigned long long)n_pass);
        slot0 = __shfl_sync(0xffffffffu, slot0, leader);
        if (pass) write_record(s_stage + (uint32_t)__popc(ballot & ((1u << lane) - 1)) * 106u, s, a, b);
        __syncwarp();
        // the records that fit the buffer (the host re-runs the batch after an overflow), as one contiguous run
        const unsigned long long room = slot0 < rec_capacity ? rec_capacity - slot0 : 0ull;
        const uint32_t total_h = (uint32_t)(room < n_pass ? room : n_pass) * 53u;  // half-words to copy
        uint8_t* g = records + slot0 * 106ull;
        const uint16_t* st16 = reinterpret_cast<const uint16_t*>(s_stage);
        uint32_t head_h = (uint32_t)((8u - (uint32_t)(reinterpret_cast<uintptr_t>(g) & 7u)) & 7u) >> 1;  // to the first 8-byte boundary
        head_h = min(head_h, total_h);
        if ((uint32_t)lane < head_h) reinterpret_cast<uint16_t*>(g)[lane] = st16[lane];
        const uint32_t n_words = (total_h - head_h) >> 2;
        unsigned long long* g8 = reinterpret_cast<unsigned long long*>(g + 2u * head_h);
        if ((head_h & 1u) == 0u) {  // staging offset 4-byte aligned
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(st16 + head_h);
            for (uint32_t w = lane; w < n_words; w += 32) g8[w] = (unsigned long long)s32[2 * w] | ((unsigned long long)s32[2 * w + 1] << 32);
        } else {
            const uint16_t* sh = st16 + head_h;
            for (uint32_t w = lane; w < n_words; w += 32) {
                const uint32_t lo = (uint32_t)sh[4 * w] | ((uint32_t)sh[4 * w + 1] << 16);
                const uint32_t hi = (uint32_t)sh[4 * w + 2] | ((uint32_t)sh[4 * w + 3] << 16);
                g8[w] = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
        }
        const uint32_t done_h = head_h + 4u * n_words;
        if ((uint32_t)lane < total_h - done_h) reinterpret_cast<uint16_t*>(g)[done_h + lane] = st16[done_h + lane];
        __syncwarp();  // the staging area is rewritten by the next round
    }
}

// ---------------------------------------------------------------- decay consumer
// LD decay over distance straight from the device-resident records (reference two_reader::Decay,
// lib/two_reader.cpp:424-475, which reads them back from the .two file): a record with ridA == ridB and posA < posB
// adds its R2 and 1 to bin min((posB - posA) / bin_width, n_bins - 1). Per-block partial sums in shared memory
// (DECAY_SMEM_BINS bins at most; larger tables go straight to global atomics), one global atomic per touched bin.
constexpr int DECAY_SMEM_BINS = 1024;
__global__ void __launch_bounds__(256)
decay_kernel(const uint8_t* __restrict__ records, unsigned long long n_records, uint32_t bin_width, uint32_t n_bins,
             double* __restrict__ sum_r2, unsigned long long* __restrict__ count) {
    __shared__ double s_sum[DECAY_SMEM_BINS];
    __shared__ unsigned int s_cnt[DECAY_SMEM_BINS];
    const bool use_smem = n_bins <= (uint32_t)DECAY_SMEM_BINS;
    if (use_smem)
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x) { s_sum[b] = 0.0; s_cnt[b] = 0u; }
    __syncthreads();
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_records;
         r += (unsigned long long)gridDim.x * blockDim.x) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(records + r * 106ull);  // records are 2-byte aligned
        const uint32_t ridA = h[1] | ((uint32_t)h[2] << 16), ridB = h[3] | ((uint32_t)h[4] << 16);
        const uint32_t posA = (h[5] | ((uint32_t)h[6] << 16)) >> 2, posB = (h[7] | ((uint32_t)h[8] << 16)) >> 2;
        if (ridA != ridB || !(posA < posB)) continue;
        const unsigned long long bits = (unsigned long long)h[37] | ((unsigned long long)h[38] << 16) | ((unsigned long long)h[39] << 32) |
                                        ((unsigned long long)h[40] << 48);  // R2 at byte 74
        const double r2 = __longlong_as_double((long long)bits);
        const uint32_t bin = min((posB - posA) / bin_width, n_bins - 1u);
        if (use_smem) { atomicAdd(&s_sum[bin], r2); atomicAdd(&s_cnt[bin], 1u); }
        else { atomicAdd(&sum_r2[bin], r2); atomicAdd(&count[bin], 1ull); }
    }
    __syncthreads();
    if (use_smem)
        for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x)
            if (s_cnt[b]) { atomicAdd(&sum_r2[b], s_sum[b]); atomicAdd(&count[b], (unsigned long long)s_cnt[b]); }
}

// ---------------------------------------------------------------- aggregate consumer
// `tomahawk aggregate` straight from the device-resident records (reference two_reader::Aggregate, lib/two_reader.cpp:543-853,
// twk_agg_slave::FindRangesUnsorted / BuildMatrix, lib/aggregation.h:127-175). Pass 1: position range per contig; pass 2: raster
// of twk_sstats (include/core.h:929-990) over forward + reverse orientation of every record.
struct AggBin { unsigned long long n; double total, total_squared, min, max; };
struct AggLayout {
    const unsigned long long* base;  // [n_contigs] coordinate of position cmin[rid]: rid_offsets.range - (max - min)
    const uint32_t* cmin;            // [n_contigs]
    uint32_t n_contigs, xrange, yrange, xbins, ybins;
    int field;
};

__device__ __forceinline__ uint32_t rec_u32(const uint16_t* h, int byte) { return h[byte >> 1] | ((uint32_t)h[(byte >> 1) + 1] << 16); }
__device__ __forceinline__ double rec_f64(const uint16_t* h, int byte) {
    const int k = byte >> 1;
    const unsigned long long bits = (unsigned long long)h[k] | ((unsigned long long)h[k + 1] << 16) | ((unsigned long long)h[k + 2] << 32) |
                                    ((unsigned long long)h[k + 3] << 48);
    return __longlong_as_double((long long)bits);
}

__global__ void __launch_bounds__(256)
agg_range_kernel(const uint8_t* __restrict__ records, unsigned long long n_records, uint32_t n_contigs, uint32_t* __restrict__ cmin,
                 uint32_t* __restrict__ cmax, unsigned long long* __restrict__ bad) {
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_records;
         r += (unsigned long long)gridDim.x * blockDim.x) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(records + r * 106ull);
        const uint32_t rid[2] = {rec_u32(h, 2), rec_u32(h, 6)};
        const uint32_t pos[2] = {rec_u32(h, 10) >> 2, rec_u32(h, 14) >> 2};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (rid[k] >= n_contigs) { atomicAdd(bad, 1ull); continue; }
            // most records improve nothing: look before paying for the atomic
            if (pos[k] < cmin[rid[k]]) atomicMin(&cmin[rid[k]], pos[k]);
            if (pos[k] > cmax[rid[k]]) atomicMax(&cmax[rid[k]], pos[k]);
        }
    }
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long old = *reinterpret_cast<unsigned long long*>(addr);
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long seen = atomicCAS(reinterpret_cast<unsigned long long*>(addr), old, (unsigned long long)__double_as_longlong(v));
        if (seen == old) break;
        old = seen;
    }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
    unsigned long long old = *reinterpret_cast<unsigned long long*>(addr);
    while (v > __longlong_as_double((long long)old)) {
        const unsigned long long seen = atomicCAS(reinterpret_cast<unsigned long long*>(addr), old, (unsigned long long)__double_as_longlong(v));
        if (seen == old) break;
        old = seen;
    }
}

__global__ void __launch_bounds__(256)
agg_bin_kernel(const uint8_t* __restrict__ records, unsigned long long n_records, AggLayout L, AggBin* __restrict__ bins) {
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_records;
         r += (unsigned long long)gridDim.x * blockDim.x) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(records + r * 106ull);
        const uint32_t ridA = rec_u32(h, 2), ridB = rec_u32(h, 6);
        if (ridA >= L.n_contigs || ridB >= L.n_contigs) continue;
        const uint32_t posA = rec_u32(h, 10) >> 2, posB = rec_u32(h, 14) >> 2;
        double v;
        switch (L.field) {
            case 1: v = rec_f64(h, 66); break;   // R
            case 2: v = rec_f64(h, 50); break;   // D
            case 3: v = rec_f64(h, 58); break;   // Dprime
            case 4: v = rec_f64(h, 82); break;   // P
            case 5: case 6: {
                const double c0 = rec_f64(h, 18), c1 = rec_f64(h, 26), c2 = rec_f64(h, 34), c3 = rec_f64(h, 42);
                const double tot = __dadd_rn(__dadd_rn(__dadd_rn(c0, c1), c2), c3);
                v = L.field == 5 ? __ddiv_rn(__dadd_rn(c1, c2), tot) : __ddiv_rn(c3, tot);
                break;
            }
            default: v = rec_f64(h, 74); break;  // R2
        }
        // aggregation.h:157: (range - (max - min)) + (pos - min)
        const unsigned long long ca = L.base[ridA] + (unsigned long long)(uint32_t)(posA - L.cmin[ridA]);
        const unsigned long long cb = L.base[ridB] + (unsigned long long)(uint32_t)(posB - L.cmin[ridB]);
        const uint32_t x[2] = {(uint32_t)min(ca / L.xrange, (unsigned long long)(L.xbins - 1u)), (uint32_t)min(cb / L.xrange, (unsigned long long)(L.xbins - 1u))};
        const uint32_t y[2] = {(uint32_t)min(cb / L.yrange, (unsigned long long)(L.ybins - 1u)), (uint32_t)min(ca / L.yrange, (unsigned long long)(L.ybins - 1u))};
        const double vv = __dmul_rn(v, v);
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // forward, reverse
            AggBin* b = bins + (size_t)x[k] * L.ybins + y[k];
            atomicAdd(&b->n, 1ull);
            atomicAdd(&b->total, v);
            atomicAdd(&b->total_squared, vv);
            atomic_min_f64(&b->min, v);
            atomic_max_f64(&b->max, v);
        }
    }
}

}  // namespace twkb
