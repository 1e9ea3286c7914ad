// pack.cuh -- builds the device-resident, transposed bit planes from the
// reference's row-major twk_igt_vec rows (lib/core.cpp:349-383: haplotype p of a
// variant is bit p%64 of word p/64; mask has both bits of a sample set when
// either allele is missing).
//
// Output planes are word-major ("transposed"): plane[p][k][v], v contiguous,
// Mpad (multiple of 128) variants per row, K32 (multiple of 16) words per
// variant, zero padded. Phased planes hold one bit per haplotype; unphased
// planes one bit per sample (het = a0^a1, hom = a0&a1, valid), which halves the
// word count of the 3x3 path.
#pragma once
#include "common.cuh"

namespace twkb {

// keep bits 0,2,4,... of x and squeeze them into the low 32 bits
__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

// mode: CountMode. One thread block transposes a 32(variant) x 32(word) patch.
__global__ void pack_planes_kernel(const uint64_t* __restrict__ data, const uint64_t* __restrict__ mask,
                                   size_t stride64, uint32_t n_variants, uint32_t n_samples, int mode,
                                   uint32_t* __restrict__ planes, uint32_t K32, uint32_t Mpad) {
    __shared__ uint32_t tile[3][32][33];
    const uint32_t v0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    const bool unphased = mode >= 2;
    const uint32_t n_bits = 2 * n_samples;
    const int np = (mode == 0) ? 1 : (mode == 3 ? 3 : 2);
    for (int r = ty; r < 32; r += 8) {
        const uint32_t v = v0 + r, k = k0 + tx;
        uint32_t p0 = 0, p1 = 0, p2 = 0;
        if (v < n_variants) {
            if (!unphased) {
                // 32-bit word k of the haplotype bit vector
                const uint32_t w64 = k >> 1;
                if ((size_t)w64 < stride64 && (uint64_t)k * 32 < n_bits) {
                    const uint64_t d = data[(size_t)v * stride64 + w64];
                    const uint64_t m = mask ? mask[(size_t)v * stride64 + w64] : 0ull;
                    uint32_t d32 = (uint32_t)(d >> ((k & 1) * 32)), m32 = (uint32_t)(m >> ((k & 1) * 32));
                    uint32_t live = 0xffffffffu;
                    if ((uint64_t)k * 32 + 32 > n_bits) live = (1u << (n_bits - k * 32)) - 1u;
                    const uint32_t valid = ~m32 & live;
                    p0 = d32 & valid;
                    p1 = valid;
                }
            } else {
                // 32 samples = 64-bit word k of the haplotype bit vector
                if ((size_t)k < stride64 && (uint64_t)k * 32 < n_samples) {
                    const uint64_t d = data[(size_t)v * stride64 + k];
                    const uint64_t m = mask ? mask[(size_t)v * stride64 + k] : 0ull;
                    uint32_t live = 0xffffffffu;
                    if ((uint64_t)k * 32 + 32 > n_samples) live = (1u << (n_samples - k * 32)) - 1u;
                    const uint32_t valid = ~compress_even(m | (m >> 1)) & live;
                    p0 = compress_even(d ^ (d >> 1)) & valid;
                    p1 = compress_even(d & (d >> 1)) & valid;
                    p2 = valid;
                }
            }
        }
        tile[0][r][tx] = p0;
        tile[1][r][tx] = p1;
        tile[2][r][tx] = p2;
    }
    __syncthreads();
    const size_t plane_stride = (size_t)K32 * Mpad;
    for (int r = ty; r < 32; r += 8) {
        const uint32_t k = k0 + r, v = v0 + tx;
        if (k < K32 && v < Mpad) {
            for (int p = 0; p < np; ++p) planes[p * plane_stride + (size_t)k * Mpad + v] = tile[p][tx][r];
        }
    }
}

// per-variant popcount of every plane: pp[p][v]
__global__ void plane_popc_kernel(const uint32_t* __restrict__ planes, int np, uint32_t K32, uint32_t Mpad,
                                  uint32_t* __restrict__ pp) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (v >= Mpad || p >= np) return;
    const uint32_t* src = planes + (size_t)p * K32 * Mpad + v;
    uint32_t s = 0;
    for (uint32_t k = 0; k < K32; ++k) s += __popc(src[(size_t)k * Mpad]);
    pp[(size_t)p * Mpad + v] = s;
}

}  // namespace twkb
