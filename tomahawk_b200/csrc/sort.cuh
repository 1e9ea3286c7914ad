// sort.cuh -- device-side sort of the resident records ("a GPU/CPU sorter equivalent to two_reader::Sort so
// output is queryable"). Reference: `tomahawk sort` orders a .two by twk1_two_t::operator< (lib/core.cpp:458-468: ridA, ridB,
// Apos, Bpos) after `calc` wrote forward and reverse copies in arrival order (lib/ld/ld_engine.cpp:1290-1298,
// lib/two_reader.cpp:162-420). Here the forward records are still in HBM when the computation ends: both orientations get a
// 128-bit key + a 32-bit reference, the references are ordered by an LSD radix sort (8-bit digits; digits that are equal in
// all keys -- contig ids on one chromosome, the high position bits -- are skipped), and one gather writes the records in
// file order, so the host only cuts blocks and compresses (SortedTwoWriter).
//
// The sort is stable; the key is the one hostio.cpp's sorter uses (rid biased to unsigned, the raw position words), so both
// produce the same file.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace twkb {

constexpr int SORT_THREADS = 256;              // 8 warps
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS_PER_LANE = 16;
constexpr int SORT_TILE = 32 * SORT_ITEMS_PER_LANE;  // items one warp ranks, in order

__device__ __forceinline__ uint32_t srec_u32(const uint16_t* h, int byte) { return h[byte >> 1] | ((uint32_t)h[(byte >> 1) + 1] << 16); }

// item t = 2 * record + orientation (0 forward, 1 reverse). or_and[0..1] = OR of (hi, lo), [2..3] = AND.
__global__ void __launch_bounds__(SORT_THREADS)
sort_keys_kernel(const uint8_t* __restrict__ records, unsigned long long n_items, unsigned long long* __restrict__ key_hi,
                 unsigned long long* __restrict__ key_lo, uint32_t* __restrict__ ref, unsigned long long* __restrict__ or_and) {
    unsigned long long o_hi = 0, o_lo = 0, a_hi = ~0ull, a_lo = ~0ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_items;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint16_t* h = reinterpret_cast<const uint16_t*>(records + (t >> 1) * 106ull);
        uint32_t ridA = srec_u32(h, 2), ridB = srec_u32(h, 6), pA = srec_u32(h, 10), pB = srec_u32(h, 14);
        if (t & 1ull) { uint32_t x = ridA; ridA = ridB; ridB = x; x = pA; pA = pB; pB = x; }
        // rid is int32 in the reference's comparison: bias so that unsigned order == signed order
        const unsigned long long hi = ((unsigned long long)(ridA ^ 0x80000000u) << 32) | (ridB ^ 0x80000000u);
        const unsigned long long lo = ((unsigned long long)pA << 32) | pB;
        key_hi[t] = hi; key_lo[t] = lo; ref[t] = (uint32_t)t;
        o_hi |= hi; o_lo |= lo; a_hi &= hi; a_lo &= lo;
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        o_hi |= __shfl_xor_sync(0xffffffffu, o_hi, s); o_lo |= __shfl_xor_sync(0xffffffffu, o_lo, s);
        a_hi &= __shfl_xor_sync(0xffffffffu, a_hi, s); a_lo &= __shfl_xor_sync(0xffffffffu, a_lo, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], o_hi); atomicOr(&or_and[1], o_lo);
        atomicAnd(&or_and[2], a_hi); atomicAnd(&or_and[3], a_lo);
    }
}

__device__ __forceinline__ uint32_t sort_digit(const unsigned long long* __restrict__ key_hi, const unsigned long long* __restrict__ key_lo,
                                               unsigned long long t, int pass) {
    const unsigned long long k = pass < 8 ? key_lo[t] : key_hi[t];
    return (uint32_t)(k >> ((pass & 7) * 8)) & 0xffu;
}

// counts[digit * n_tiles + tile]: a warp owns tile = items [tile * SORT_TILE, ...)
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const unsigned long long* __restrict__ key_hi, const unsigned long long* __restrict__ key_lo, unsigned long long n_items,
                  int pass, uint32_t n_tiles, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_cnt[SORT_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * SORT_WARPS + warp;
    for (int d = lane; d < 256; d += 32) s_cnt[warp][d] = 0;
    __syncwarp();
    if (tile < n_tiles) {
        const unsigned long long base = (unsigned long long)tile * SORT_TILE;
#pragma unroll 4
        for (int s = 0; s < SORT_ITEMS_PER_LANE; ++s) {
            const unsigned long long t = base + (unsigned long long)s * 32 + lane;
            if (t < n_items) atomicAdd(&s_cnt[warp][sort_digit(key_hi, key_lo, t, pass)], 1u);
        }
        __syncwarp();
        for (int d = lane; d < 256; d += 32) counts[(size_t)d * n_tiles + tile] = s_cnt[warp][d];
    }
}

// Exclusive prefix sum over `n` counters, in place, one block (n = 256 * n_tiles: a few million at most).
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(uint32_t* __restrict__ data, unsigned long long n) {
    __shared__ unsigned long long s_part[1024];
    const unsigned long long per = (n + 1023) / 1024;
    const unsigned long long lo = (unsigned long long)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0;
    for (unsigned long long i = lo; i < hi; ++i) sum += data[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele over the 1,024 partial sums
        unsigned long long v = threadIdx.x >= (unsigned)off ? s_part[threadIdx.x - off] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = threadIdx.x ? s_part[threadIdx.x - 1] : 0ull;
    for (unsigned long long i = lo; i < hi; ++i) {
        const uint32_t c = data[i];
        data[i] = (uint32_t)run;
        run += c;
    }
}

// Stable scatter: a warp walks its tile in order, 32 items per step; lanes with the same digit rank themselves by lane id.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const unsigned long long* __restrict__ key_hi, const unsigned long long* __restrict__ key_lo,
                     const uint32_t* __restrict__ ref, unsigned long long n_items, int pass, uint32_t n_tiles,
                     const uint32_t* __restrict__ offsets, unsigned long long* __restrict__ out_hi, unsigned long long* __restrict__ out_lo,
                     uint32_t* __restrict__ out_ref) {
    __shared__ uint32_t s_off[SORT_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * SORT_WARPS + warp;
    if (tile >= n_tiles) return;
    for (int d = lane; d < 256; d += 32) s_off[warp][d] = offsets[(size_t)d * n_tiles + tile];
    __syncwarp();
    const unsigned long long base = (unsigned long long)tile * SORT_TILE;
    for (int s = 0; s < SORT_ITEMS_PER_LANE; ++s) {
        const unsigned long long t = base + (unsigned long long)s * 32 + lane;
        const bool valid = t < n_items;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        if (!active) break;
        if (valid) {
            const unsigned long long hi = key_hi[t], lo = key_lo[t];
            const uint32_t d = (uint32_t)((pass < 8 ? lo : hi) >> ((pass & 7) * 8)) & 0xffu;
            const unsigned peers = __match_any_sync(active, d);
            const uint32_t pos = s_off[warp][d] + __popc(peers & ((1u << lane) - 1u));
            __syncwarp(active);
            if (lane == __ffs(peers) - 1) s_off[warp][d] += __popc(peers);
            __syncwarp(active);
            out_hi[pos] = hi; out_lo[pos] = lo; out_ref[pos] = ref[t];
        }
    }
}

// out = the records in sorted order, reverse copies with (rid, position) of A and B swapped -- everything else stays
// A-major like the reference's reverse record. One thread per 16-bit word of the output (records are 2-byte aligned).
__global__ void __launch_bounds__(256)
sort_gather_kernel(const uint8_t* __restrict__ records, const uint32_t* __restrict__ ref, unsigned long long first_item,
                   unsigned long long n_items, uint8_t* __restrict__ out) {
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_items * 53ull) return;
    const unsigned long long k = e / 53ull;
    uint32_t w = (uint32_t)(e - k * 53ull);
    const uint32_t r = ref[first_item + k];
    if ((r & 1u) && w >= 1u && w <= 8u) w = w <= 4u ? (w <= 2u ? w + 2u : w - 2u) : (w <= 6u ? w + 2u : w - 2u);  // words 1,2 <-> 3,4 and 5,6 <-> 7,8
    reinterpret_cast<uint16_t*>(out)[e] = reinterpret_cast<const uint16_t*>(records + (unsigned long long)(r >> 1) * 106ull)[w];
}

}  // namespace twkb
