// count_sparse.cuh -- rare-variant (list-encoded) count kernel.
//
// B200 counterpart of the reference's list path
//   twk_igt_list::Build      include/core.h:517-672   (r_pos: offsets of the 128-bit registers
//                                                      that hold at least one alt allele)
//   PhasedListVector         lib/ld/ld_engine.cpp:185-267 (n11 = sum popc(A[r] & B[r]) over the
//                                                      registers of the sparser variant; the
//                                                      other three cells from ac, :244-246)
// The reference walks the non-empty registers of the sparser variant of every pair. Here the
// variants whose haplotype row has at most `T` non-zero 32-bit words ("sparse" variants) are
// kept as a CSR list of (word index k, word value w) entries; all other variants stay dense.
// The engine orders the resident matrix [dense variants | sparse variants] (both in file
// order), runs the tensor-core kernel on the dense x dense triangle only, and this kernel on
// every pair that has a sparse member:
//
//   n11(r, j) = sum over the entries (k, w) of r of popc(w & plane[k][j])
//
// plane[k][j] is the word-major bit plane every other kernel reads (pack.cuh), so for a fixed
// entry the 256 threads of a CTA read 1 KB of consecutive columns: coalesced, and at most
// nnz(r) <= T word operations per pair instead of ceil(2N/32) (POPC kernel) or 2N MACs
// (tensor kernel). A CTA owns SP_ROWS sparse rows x SP_TJ columns; a thread owns SP_CPT
// columns 256 apart and walks the rows' entries (staged through shared memory, 256 at a
// time); after the last entry of a row the per-pair epilogue runs: pair rules, fp32
// conservative screen, exact fp64 decision (pair_decide), warp-aggregated compaction.
//
// Pairs are oriented by the ORIGINAL variant order (A = lower file index), as everywhere else.
#pragma once
#include "common.cuh"
#include "count_popc.cuh"

namespace twkb {

constexpr int SP_THREADS = 256;
constexpr int SP_CPT = 4;                       // columns per thread
constexpr int SP_TJ = SP_THREADS * SP_CPT;      // 1024 columns per tile
constexpr int SP_ROWS = 32;                     // sparse rows per tile
constexpr int SP_CHUNK = 256;                   // entries staged per step

struct SparseArgs {
    const uint32_t* sp_off;   // [nS + 1] entry offsets; sparse row s is resident variant nD + s
    const uint2* sp_ent;      // (k, w): word k of the variant's haplotype row equals w != 0
    const uint32_t* orig;     // [Mpad] original (file-order) index of every resident variant
    uint32_t nD;              // number of dense variants = resident index of the first sparse one
    uint32_t nS;
};

// non-zero 32-bit words per variant row (reference rows: u64 words, LSB-first)
__global__ void row_nnz32_kernel(const uint64_t* __restrict__ rows, size_t stride64, uint32_t n_variants, uint32_t n_bits,
                                 uint32_t* __restrict__ nnz) {
    const uint32_t v = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= n_variants) return;
    const uint32_t n64 = (n_bits + 63) / 64;
    uint32_t c = 0;
    for (uint32_t w = lane; w < n64; w += 32) {
        uint64_t x = rows[(size_t)v * stride64 + w];
        if ((uint64_t)w * 64 + 64 > n_bits) x &= (1ull << (n_bits - w * 64)) - 1ull;
        c += ((uint32_t)x != 0u) + ((uint32_t)(x >> 32) != 0u);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) nnz[v] = c;
}

// dst row x = src row perm[x] (the [dense | sparse] ordering of the resident matrix)
__global__ void gather_rows_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, const uint32_t* __restrict__ perm,
                                   size_t stride64, uint32_t n_variants) {
    const uint32_t x = blockIdx.x;
    if (x >= n_variants) return;
    const uint64_t* s = src + (size_t)perm[x] * stride64;
    uint64_t* d = dst + (size_t)x * stride64;
    for (size_t w = threadIdx.x; w < stride64; w += blockDim.x) d[w] = s[w];
}

// CSR entries of the sparse rows (one warp per row, entries in increasing k)
__global__ void build_sparse_entries_kernel(const uint64_t* __restrict__ rows, size_t stride64, uint32_t nD, uint32_t nS,
                                            uint32_t n_bits, const uint32_t* __restrict__ sp_off, uint2* __restrict__ ent) {
    const uint32_t s = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= nS) return;
    const uint64_t* row = rows + (size_t)(nD + s) * stride64;
    const uint32_t n32 = (n_bits + 31) / 32;
    uint32_t out = sp_off[s];
    for (uint32_t k0 = 0; k0 < n32; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint32_t w = 0;
        if (k < n32) {
            const uint64_t x = row[k >> 1];
            w = (uint32_t)(x >> ((k & 1) * 32));
            if ((uint64_t)k * 32 + 32 > n_bits) w &= (1u << (n_bits - k * 32)) - 1u;
        }
        const unsigned b = __ballot_sync(0xffffffffu, w != 0u);
        if (w != 0u) ent[out + __popc(b & ((1u << lane) - 1u))] = make_uint2(k, w);
        out += __popc(b);
    }
}

template <bool SCREEN>
__global__ void __launch_bounds__(SP_THREADS, 4) count_sparse_kernel(CountArgs args, SparseArgs sp, DevParams prm) {
    __shared__ uint2 s_ent[SP_CHUNK];
    __shared__ uint32_t s_off[SP_ROWS + 1];
    const uint2 tile = args.tiles[blockIdx.x];
    const uint32_t r0 = tile.x, j0 = tile.y;  // resident indices: first sparse row, first column
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t M = prm.n_variants;
    const uint32_t n_rows = min((uint32_t)SP_ROWS, M - r0);
    if (tid <= (int)n_rows) s_off[tid] = sp.sp_off[r0 - sp.nD + tid];

    // per-column constants (SP_CPT columns, 256 apart)
    uint32_t jj[SP_CPT], oj[SP_CPT], acj[SP_CPT];
    int coff[SP_CPT];  // column offset from col_base (columns past Mpad read column j0: never used)
    float dB[SP_CPT];
    const float Tf = (float)(2u * prm.n_samples);
    const float thr = (float)prm.screenR2 * (1.0f - 1.0e-5f);
#pragma unroll
    for (int c = 0; c < SP_CPT; ++c) {
        jj[c] = j0 + (uint32_t)(c * SP_THREADS + tid);
        const bool in = jj[c] < M;
        acj[c] = in ? args.meta[jj[c]].ac : 0u;
        oj[c] = in ? sp.orig[jj[c]] : 0xffffffffu;
        coff[c] = jj[c] < args.Mpad ? c * SP_THREADS : -tid;
        const float acB = (float)acj[c];
        dB[c] = acB * (Tf - acB);
    }
    const uint32_t* col_base = args.planes + j0 + tid;  // plane 0 (haplotype bits), word-major
    __syncthreads();

    const uint32_t e_begin = s_off[0], e_end = s_off[n_rows];
    uint32_t row = 0, row_end = s_off[1];
    uint32_t acc[SP_CPT];
#pragma unroll
    for (int c = 0; c < SP_CPT; ++c) acc[c] = 0;

    // Per-row epilogue; called convergently by the whole CTA.
    auto finish_row = [&](uint32_t rr) {
        const uint32_t r = r0 + rr;
        const DevVariant vi = args.meta[r];
        const uint32_t oi = sp.orig[r];
        const float acA = (float)vi.ac, dA = acA * (Tf - acA);
#pragma unroll
        for (int c = 0; c < SP_CPT; ++c) {
            const uint32_t j = jj[c];
            // pairs of this kernel: sparse row r with every dense column and every LATER sparse column
            bool pre = j < M && (j < sp.nD || j > r) && (vi.ac + acj[c] > 2);
            if (SCREEN && pre) {
                const float n11 = (float)acc[c];
                const float pab = acA * (float)acj[c];
                const float x = fabsf(fmaf(n11, Tf, -pab));
                const float slack = 4.0f + 4.0e-7f * fmaxf(n11 * Tf, pab);
                pre = (x + slack) * (x + slack) >= thr * (dA * dB[c]);
            }
            if (__any_sync(0xffffffffu, pre)) {
                PairAcc<1> pa;
                pa.v[0][0] = acc[c];
                const bool swap = oj[c] < oi;  // A is the variant that comes first in the file
                DevVariant vjc{0, 0, 0, 0};
                if (pre) vjc = args.meta[j];
                emit_pair_with<MODE_PHASED_NOMISS>(args, prm, swap ? j : r, swap ? r : j, swap ? vjc : vi, swap ? vi : vjc, pa, lane, pre);
            }
            acc[c] = 0;
        }
    };

    for (uint32_t e0 = e_begin; e0 < e_end; e0 += SP_CHUNK) {
        const uint32_t n = min((uint32_t)SP_CHUNK, e_end - e0);
        __syncthreads();
        if ((uint32_t)tid < n) s_ent[tid] = sp.sp_ent[e0 + tid];
        __syncthreads();
        uint32_t e = 0;
        while (e < n) {
            while (e0 + e >= row_end) {  // rows that ended at or before this entry (empty rows included)
                finish_row(row);
                ++row;
                row_end = s_off[row + 1];
            }
            const uint32_t seg = min(n, row_end - e0);
#pragma unroll 4
            for (; e < seg; ++e) {
                const uint2 kw = s_ent[e];
                const uint32_t* p = col_base + (size_t)kw.x * args.Mpad;
#pragma unroll
                for (int c = 0; c < SP_CPT; ++c) acc[c] += __popc(kw.y & __ldg(p + coff[c]));
            }
        }
    }
    for (; row < n_rows; ++row) finish_row(row);  // the last row, and trailing empty rows
}

}  // namespace twkb
