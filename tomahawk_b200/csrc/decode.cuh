// decode.cuh -- device-side decode of the `.twk` run-length genotype records into the
// reference-layout bit rows (+ mask rows) every other kernel starts from.
//
// B200 counterpart of twk_igt_vec::Build (reference lib/core.cpp:349-383): for every run
// (include/core.h:188-256; len samples, allele codes a, b) the reference sets bit 2s of `data`
// when a == 1 and bit 2s+1 when b == 1 for each sample s of the run, and both bits of `mask`
// when either allele is missing (code 2, only in the 2-bit encoding; :379-380).
//
// One warp per variant. Per step the 32 lanes take 32 consecutive run words, a warp prefix sum
// of the run lengths gives every lane its sample range, and the lane ORs the 2-bit-periodic
// pattern of its run into the row: boundary words with atomicOr (neighbouring runs share them),
// fully covered words with plain stores (no other run can touch them). ref/ref runs -- almost
// all of the samples of a rare variant -- write nothing: the rows are zero-filled beforehand by
// a memset. Runs whose interior is longer than 64 words are written by the whole warp.
// HBM-bound on the row writes (M * stride bytes); the run bytes are read once.
#pragma once
#include "../../include/twkb.h"
#include "common.cuh"

namespace twkb {

constexpr int DEC_WARPS = 8;
constexpr uint32_t DEC_LONG_WORDS = 64;

// OR `pat` into bits [s, e) of a row of 32-bit words; s < e. Interior words are stored.
__device__ __forceinline__ void dec_or_range(uint32_t* row, uint32_t s, uint32_t e, uint32_t pat, bool interior) {
    const uint32_t w0 = s >> 5, w1 = (e - 1) >> 5;
    const uint32_t m0 = ~0u << (s & 31);
    const uint32_t m1 = (e & 31) ? ((1u << (e & 31)) - 1u) : ~0u;
    if (w0 == w1) {
        atomicOr(row + w0, pat & m0 & m1);
        return;
    }
    atomicOr(row + w0, pat & m0);
    if (interior)
        for (uint32_t w = w0 + 1; w < w1; ++w) row[w] = pat;
    atomicOr(row + w1, pat & m1);
}

// err[0]: 0 ok, 1 = runs exceed / do not cover the samples; err[1]: lowest offending variant.
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_runs_kernel(const uint8_t* __restrict__ bytes, const twkb_run_desc* __restrict__ desc, uint32_t n_variants, uint32_t H,
                   uint32_t* __restrict__ rows, uint32_t* __restrict__ masks /* nullable */, size_t stride32, uint32_t* err) {
    const uint32_t v = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (v >= n_variants) return;
    const uint32_t lane = threadIdx.x & 31;
    const twkb_run_desc d = desc[v];
    const uint8_t* p = bytes + d.offset;
    uint32_t* row = rows + (size_t)v * stride32;
    uint32_t* mrow = masks ? masks + (size_t)v * stride32 : nullptr;
    const uint32_t lshift = d.miss ? 4u : 2u, ashift = d.miss ? 2u : 1u, amask = d.miss ? 3u : 1u;
    uint64_t base = 0;  // samples covered so far (64-bit: a corrupt file must not wrap)
    bool bad = false;
    for (uint32_t k0 = 0; k0 < d.n_runs; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint32_t word = 0;
        if (k < d.n_runs) {
            const uint8_t* q = p + (size_t)k * d.width;
            word = q[0];
            if (d.width >= 2) word |= (uint32_t)q[1] << 8;
            if (d.width == 4) word |= (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24;
        }
        const uint32_t len = word >> lshift;
        const uint32_t a = (word >> ashift) & amask, b = word & amask;
        unsigned long long incl = len;  // warp inclusive scan, 64-bit: 32 corrupt run lengths of 2^30 must not wrap
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        const uint64_t s_samp = base + (incl - len), e_samp = base + incl;
        base += __shfl_sync(0xffffffffu, incl, 31);
        uint32_t pat = (a == 1 ? 0x55555555u : 0u) | (b == 1 ? 0xAAAAAAAAu : 0u);
        uint32_t mpat = (mrow && (a == 2 || b == 2)) ? 0xFFFFFFFFu : 0u;
        if (2 * e_samp > H) { bad = true; pat = 0; mpat = 0; }
        if (len == 0) { pat = 0; mpat = 0; }
        const uint32_t s = (uint32_t)(2 * s_samp), e = (uint32_t)(2 * e_samp);
        const bool is_long = (pat | mpat) != 0 && ((e - 1) >> 5) - (s >> 5) > DEC_LONG_WORDS;
        if (pat) dec_or_range(row, s, e, pat, !is_long);
        if (mpat) dec_or_range(mrow, s, e, mpat, !is_long);
        uint32_t long_lanes = __ballot_sync(0xffffffffu, is_long);
        while (long_lanes) {  // interiors of long non-reference runs: all 32 lanes write
            const int src = __ffs(long_lanes) - 1;
            long_lanes &= long_lanes - 1;
            const uint32_t ls = __shfl_sync(0xffffffffu, s, src), le = __shfl_sync(0xffffffffu, e, src);
            const uint32_t lp = __shfl_sync(0xffffffffu, pat, src), lm = __shfl_sync(0xffffffffu, mpat, src);
            const uint32_t w0 = ls >> 5, w1 = (le - 1) >> 5;
            for (uint32_t w = w0 + 1 + lane; w < w1; w += 32) {
                if (lp) row[w] = lp;
                if (lm) mrow[w] = lm;
            }
        }
    }
    if (2 * base != H) bad = true;
    if (__any_sync(0xffffffffu, bad) && lane == 0) {
        atomicExch(err, 1u);
        atomicMin(err + 1, v);
    }
}

}  // namespace twkb
