// count_popc.cuh -- tiled LOP3+POPC contingency-count kernel family with the
// fused R2 pre-screen + warp-aggregated stream compaction.
//
// One kernel replaces the reference's per-pair comparators
//   PhasedListVector / PhasedVectorizedNoMissing   lib/ld/ld_engine.cpp:185-267, 636-707
//   PhasedVectorized (masked)                      :513-634
//   UnphasedVectorized / ...NoMissing              :709-1009
//   Phased/UnphasedRunlength                       :1011-1160 (same counts)
// and the pair loops of twk_ld_slave::{Phased,Unphased,Calculate*} (:1898-2838).
//
// Data layout: every bit plane of the genotype matrix is stored
// transposed, word-major: plane[k][v] is 32-bit word k of variant v, v contiguous.
// A CTA computes a TI x TJ tile of variant pairs. It streams
// [TK words] x [TI | TJ variants] slabs of every plane into a 4-stage shared
// memory ring with 1-D TMA bulk copies (cp.async.bulk, one 256/512-byte row per
// lane) signalled through mbarriers (warp 0 doubles as the producer, S-1 chunks
// ahead); all 8 warps each own a register tile
// of TM x TN pairs x NP^2 plane products and do acc += popc(a & b).
// Shared-memory reads are LDS.128: A rows broadcast inside a half-warp, B
// columns are 16 consecutive 16-byte chunks (conflict free).
//
// Epilogue (per pair, in registers): build the exact 2x2 / 3x3 table, apply
// the pair rules of the reference (i<j on diagonal tiles, ac_i+ac_j<=2 skip,
// window rule Q7), a conservative R2 screen in fp64, and append survivors to
// the candidate buffer with one atomic per warp-ballot.
#pragma once
#include "common.cuh"

namespace twkb {

enum CountMode : int {
    MODE_PHASED_NOMISS = 0,   // planes: data
    MODE_PHASED_MISS = 1,     // planes: data&valid, valid          (haplotype bits)
    MODE_UNPHASED_NOMISS = 2, // planes: het, hom                   (sample bits)
    MODE_UNPHASED_MISS = 3,   // planes: het&valid, hom&valid, valid (sample bits)
};

template <int MODE> struct PopcCfg;
template <> struct PopcCfg<MODE_PHASED_NOMISS>   { static constexpr int NP = 1, TM = 8, TN = 8; };
template <> struct PopcCfg<MODE_PHASED_MISS>     { static constexpr int NP = 2, TM = 4, TN = 8; };
template <> struct PopcCfg<MODE_UNPHASED_NOMISS> { static constexpr int NP = 2, TM = 4, TN = 8; };
template <> struct PopcCfg<MODE_UNPHASED_MISS>   { static constexpr int NP = 3, TM = 4, TN = 4; };

constexpr int POPC_TK = 16;       // 32-bit words per pipeline stage
constexpr int POPC_STAGES = 4;
constexpr int POPC_CONSUMERS = 256;
constexpr int POPC_THREADS = POPC_CONSUMERS;

template <int MODE> __host__ __device__ constexpr int popc_tile_i() { return 16 * PopcCfg<MODE>::TM; }
template <int MODE> __host__ __device__ constexpr int popc_tile_j() { return 16 * PopcCfg<MODE>::TN; }
template <int MODE> __host__ __device__ constexpr size_t popc_smem_bytes() {
    return (size_t)POPC_STAGES * PopcCfg<MODE>::NP * POPC_TK * (popc_tile_i<MODE>() + popc_tile_j<MODE>()) * 4 + 128;
}

struct CountArgs {
    const uint32_t* planes;      // [NP][K32][Mpad]
    uint32_t K32;                // words per variant per plane (multiple of POPC_TK)
    uint32_t Mpad;               // padded variant count (multiple of 128)
    const uint2* tiles;          // (i0, j0) of every tile of this launch
    const DevVariant* meta;      // [Mpad]
    const uint32_t* plane_popc;  // [NP][Mpad] per-variant popcount of each plane
    DevBlocks blocks;            // window mode only
    Candidate* cands;
    unsigned long long* cand_count;
    unsigned long long cand_capacity;
    uint32_t row_begin, row_end; // variant index limits of this problem (rows)
    uint32_t col_begin, col_end; // (cols)
    uint32_t screen_off;         // 1: every enumerated pair becomes a candidate (debug/test)
    uint32_t debug_flags;        // profiling aids (env TWKB_DEBUG_FLAGS); 0 in production
};

// Window-mode pair rules of the reference. DevParams::window selects:
//   1  -p / -u (CalculatePhasedWindow ld_engine.cpp:2553-2560, CalculateUnphasedWindow :2658-2664):
//      ld_balancing.h:189-196 prunes the rest of a block row by positions only, and the slave
//      abandons a block pair at its first out-of-window pair;
//   2  auto mode (twk_ld_slave::Calculate :2737-2838 has no per-pair test): the row prune only;
//   3  -p -m -M (CalculatePhasedBitmapWindow :2466-2471, :2490-2495): the row prune, and a pair is
//      skipped only when the contigs DIFFER and the wrapping position difference exceeds the window.
__device__ __forceinline__ bool window_pair_allowed(uint32_t kind, uint32_t i, uint32_t j, const DevVariant& vi, const DevVariant& vj,
                                                    const DevVariant* meta, const DevBlocks& bl, uint32_t w) {
    const uint32_t bi = bl.blk_of[i], bj = bl.blk_of[j];
    if (bi != bj && bj >= bl.blk_prune[bi]) return false;
    if (kind == 2u) return true;
    if (kind == 3u) return !(vi.rid != vj.rid && (vj.pos - vi.pos) > w);
    const uint32_t fi = bl.blk_first[bi], lj = bl.blk_last[bj];
    const DevVariant vf = meta[fi], vl = meta[lj];
    if (vf.rid == vl.rid && (vl.pos - vf.pos) > w) return i == fi && !((vj.pos - vi.pos) > w);
    return true;
}

// fp64 screens. Integer-valued doubles below 2^53 make the numerator exact.
__device__ __forceinline__ bool screen_phased(uint32_t c0, uint32_t c1, uint32_t c4, uint32_t c5, const DevParams& prm) {
    if (!(prm.minR2 > 0.0)) return true;
    const double T = (double)c0 + (double)c1 + (double)c4 + (double)c5;
    const double rA = (double)c1 + (double)c5, rB = (double)c4 + (double)c5;
    const double num = (double)c5 * T - rA * rB;
    const double den = rA * (T - rA) * rB * (T - rB);
    return num * num >= prm.screenR2 * den;
}
// Upper bound on the unphased R2: D = f11 - P*Q with f11 confined to
// [minhap - 1e-5, maxhap + 1e-5] by the reference (ld_engine.cpp:1460-1485).
__device__ __forceinline__ bool screen_unphased(const uint32_t* t, const DevParams& prm) {
    if (!(prm.minR2 > 0.0)) return true;
    const double T = (double)t[0] + t[1] + t[2] + t[3] + t[4] + t[5] + t[6] + t[7] + t[8];
    if (T < 5.0) return false;
    const double inv2T = 1.0 / (2.0 * T);
    const double P = (2.0 * ((double)t[0] + t[1] + t[2]) + ((double)t[3] + t[4] + t[5])) * inv2T;
    const double Q = (2.0 * ((double)t[0] + t[3] + t[6]) + ((double)t[1] + t[4] + t[7])) * inv2T;
    const double n11 = 2.0 * t[0] + t[1] + t[3];
    const double lo = n11 * inv2T - 1.0e-5 - P * Q, hi = (n11 + t[4]) * inv2T + 1.0e-5 - P * Q;
    const double dmax = fmax(fabs(lo), fabs(hi));
    const double den = P * (1.0 - P) * Q * (1.0 - Q);
    if (!(den > 0.0)) return true;
    return dmax * dmax * (1.0 + 1.0e-9) >= prm.minR2 * den;
}

template <int NP> struct PairAcc { uint32_t v[NP][NP]; };

// Exact per-pair decision for a pair inside the problem's ranges: pair rules of the reference
// (ac_i+ac_j<=2 skip, auto-mode pass filter, window rule Q7), the exact 2x2 / 3x3 table from the
// plane products, and the fp64 R2 screen. c[] / mode are the candidate fields.
template <int MODE>
__device__ __forceinline__ bool pair_decide(const CountArgs& args, const DevParams& prm, uint32_t i, uint32_t j, const DevVariant& vi,
                                            const DevVariant& vj, const PairAcc<PopcCfg<MODE>::NP>& pa, uint32_t (&c)[9],
                                            uint32_t& mode) {
    constexpr int NP = PopcCfg<MODE>::NP;
    bool ok = prm.single || (vi.ac + vj.ac > 2);  // ld_engine.cpp:1918 (not in CalculateSingle)
    if (prm.pair_filter) {
        const bool miss = ((vi.flags | vj.flags) & VF_HAS_MISSING) != 0;
        ok = ok && (prm.pair_filter == 1u ? !miss : miss);
    }
    if (ok && prm.window) ok = window_pair_allowed(prm.window, i, j, vi, vj, args.meta, args.blocks, prm.l_window);
    // position shard: pairs among the halo blocks belong to the shard that owns their earlier member
    if (ok && prm.shard_blocks) ok = min(args.blocks.blk_of[i], args.blocks.blk_of[j]) < prm.shard_blocks;
    if (!ok) return false;
    if (MODE == MODE_PHASED_NOMISS) {
        // ld_engine.cpp:244-246 / :682-685
        const uint32_t n11 = pa.v[0][0];
        c[3] = n11;
        c[1] = vi.ac - n11;
        c[2] = vj.ac - n11;
        c[0] = 2u * prm.n_samples - ((vi.ac + vj.ac) - n11);
        ok = args.screen_off || screen_phased(c[0], c[1], c[2], c[3], prm);
    } else if (MODE == MODE_PHASED_MISS) {
        // planes (alt&valid, valid): the four masked counts of ld_engine.cpp:555-581
        const uint32_t n11 = pa.v[0][0], nA = pa.v[0][NP - 1], nB = pa.v[NP - 1][0], nV = pa.v[NP - 1][NP - 1];
        c[3] = n11;
        c[1] = nA - n11;
        c[2] = nB - n11;
        c[0] = nV - nA - nB + n11;
        ok = args.screen_off || screen_phased(c[0], c[1], c[2], c[3], prm);
    } else {
        // unphased: plane 0 = het, 1 = hom, (2 = valid); 3x3 table of ld_engine.cpp:835-844
        uint32_t hetA_v, homA_v, hetB_v, homB_v, vv;
        if (MODE == MODE_UNPHASED_NOMISS) {
            const uint32_t* pp = args.plane_popc;
            hetA_v = pp[i]; homA_v = pp[args.Mpad + i];
            hetB_v = pp[j]; homB_v = pp[args.Mpad + j];
            vv = prm.n_samples;
        } else {
            hetA_v = pa.v[0][NP - 1]; homA_v = pa.v[1 % NP][NP - 1];
            hetB_v = pa.v[NP - 1][0]; homB_v = pa.v[NP - 1][1 % NP];
            vv = pa.v[NP - 1][NP - 1];
        }
        const uint32_t c11 = pa.v[0][0], c12 = pa.v[0][1 % NP], c21 = pa.v[1 % NP][0], c22 = pa.v[1 % NP][1 % NP];
        mode = 1;
        c[4] = c11; c[5] = c12; c[7] = c21; c[8] = c22;
        c[3] = hetA_v - c11 - c12;  // A het, B 0/0
        c[6] = homA_v - c21 - c22;  // A 1/1, B 0/0
        c[1] = hetB_v - c11 - c21;  // A 0/0, B het
        c[2] = homB_v - c12 - c22;  // A 0/0, B 1/1
        c[0] = vv - (c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7] + c[8]);
        ok = args.screen_off || screen_unphased(c, prm);
    }
    return ok;
}

// Per-pair epilogue: pair rules of the reference, exact table, screen, compaction.
// Called convergently by all 32 lanes of a warp; vj is only read by lanes with pre_ok.
template <int MODE>
__device__ __forceinline__ void emit_pair_with(const CountArgs& args, const DevParams& prm, uint32_t i, uint32_t j, const DevVariant& vi,
                                               const DevVariant& vj, const PairAcc<PopcCfg<MODE>::NP>& pa, int lane, bool pre_ok) {
    const uint32_t M = prm.n_variants;
    bool ok = pre_ok && i >= args.row_begin && i < args.row_end && j >= args.col_begin && j < args.col_end && i < M && j < M;
    if (prm.diag) ok = ok && (i < j);
    uint32_t c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = 0;
    uint32_t mode = 0;
    if (ok) ok = pair_decide<MODE>(args, prm, i, j, vi, vj, pa, c, mode);
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    if (ballot == 0) return;
    const int leader = __ffs(ballot) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(args.cand_count, (unsigned long long)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) {
        const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1));
        if (slot < args.cand_capacity) {
            uint4* dst = reinterpret_cast<uint4*>(args.cands + slot);
            dst[0] = make_uint4(i, j, c[0], c[1]);
            dst[1] = make_uint4(c[2], c[3], c[4], c[5]);
            dst[2] = make_uint4(c[6], c[7], c[8], mode);
        }
    }
}

template <int MODE>
__device__ __noinline__ void emit_pair(const CountArgs& args, const DevParams& prm, uint32_t i, uint32_t j,
                                       DevVariant vi, PairAcc<PopcCfg<MODE>::NP> pa, int lane, bool pre_ok = true) {
    const uint32_t M = prm.n_variants;
    const bool in = pre_ok && i < M && j < M;
    DevVariant vj{0, 0, 0, 0};
    if (in) vj = args.meta[j];
    emit_pair_with<MODE>(args, prm, i, j, vi, vj, pa, lane, in);
}

template <int MODE>
__global__ void __launch_bounds__(POPC_THREADS, 1) count_popc_kernel(CountArgs args, DevParams prm) {
    using Cfg = PopcCfg<MODE>;
    constexpr int NP = Cfg::NP, TM = Cfg::TM, TN = Cfg::TN;
    constexpr int TI = 16 * TM, TJ = 16 * TN, TK = POPC_TK, S = POPC_STAGES;
    constexpr int STAGE_WORDS = NP * TK * (TI + TJ);

    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * STAGE_WORDS * 4);
    uint64_t* empty_bar = full_bar + S;

    const uint2 tile = args.tiles[blockIdx.x];
    const uint32_t i0 = tile.x, j0 = tile.y;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t nchunks = args.K32 / TK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], POPC_CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Producer duty (warp 0, every lane issues one 256/512-byte row per step):
    // 1-D TMA bulk copies of chunk c into ring slot c % S.
    const size_t plane_stride = (size_t)args.K32 * args.Mpad;
    auto issue_chunk = [&](uint32_t c) {
        const int s = c % S;
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], STAGE_WORDS * 4);
        __syncwarp();
        uint32_t* sA = ring + (size_t)s * STAGE_WORDS;
        uint32_t* sB = sA + NP * TK * TI;
        for (int r = lane; r < NP * TK; r += 32) {
            const int p = r / TK, k = r % TK;
            const uint32_t* src = args.planes + p * plane_stride + (size_t)(c * TK + k) * args.Mpad;
            tma_bulk_g2s(sA + r * TI, src + i0, TI * 4, &full_bar[s]);
            tma_bulk_g2s(sB + r * TJ, src + j0, TJ * 4, &full_bar[s]);
        }
    };
    if (warp == 0) {
        for (uint32_t c = 0; c < (uint32_t)(S - 1) && c < nchunks; ++c) issue_chunk(c);
    }

    // ========================= consumer warps: LOP3 + POPC =========================
    const int tx = tid & 15, ty = tid >> 4;
    uint32_t acc[NP][NP][TM][TN];
#pragma unroll
    for (int a = 0; a < NP; ++a)
#pragma unroll
        for (int b = 0; b < NP; ++b)
#pragma unroll
            for (int ii = 0; ii < TM; ++ii)
#pragma unroll
                for (int jj = 0; jj < TN; ++jj) acc[a][b][ii][jj] = 0;

    for (uint32_t c = 0; c < nchunks; ++c) {
        const int s = c % S;
        if (warp == 0 && c + S - 1 < nchunks) {
            // slot (c-1)%S is refilled once every warp has released chunk c-1
            if (c >= 1) mbar_wait(&empty_bar[(c - 1) % S], ((c - 1) / S) & 1);
            issue_chunk(c + S - 1);
        }
        mbar_wait(&full_bar[s], (c / S) & 1);
        const uint32_t* sA = ring + (size_t)s * STAGE_WORDS;
        const uint32_t* sB = sA + NP * TK * TI;
#pragma unroll 4
        for (int k = 0; k < TK; ++k) {
            uint32_t a[NP][TM], b[NP][TN];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t* ra = sA + (p * TK + k) * TI + ty * TM;
#pragma unroll
                for (int q = 0; q < TM / 4; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(ra + 4 * q);
                    a[p][4 * q + 0] = v.x; a[p][4 * q + 1] = v.y; a[p][4 * q + 2] = v.z; a[p][4 * q + 3] = v.w;
                }
                const uint32_t* rb = sB + (p * TK + k) * TJ + 4 * tx;
#pragma unroll
                for (int q = 0; q < TN / 4; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rb + 64 * q);
                    b[p][4 * q + 0] = v.x; b[p][4 * q + 1] = v.y; b[p][4 * q + 2] = v.z; b[p][4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int pa = 0; pa < NP; ++pa)
#pragma unroll
                for (int pb = 0; pb < NP; ++pb)
#pragma unroll
                    for (int ii = 0; ii < TM; ++ii)
#pragma unroll
                        for (int jj = 0; jj < TN; ++jj) acc[pa][pb][ii][jj] += __popc(a[pa][ii] & b[pb][jj]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ================================ epilogue ================================
    // Fully unrolled so every accumulator index is static; the per-pair work
    // lives in one non-inlined function to keep the unrolled body small.
#pragma unroll
    for (int ii = 0; ii < TM; ++ii) {
        const uint32_t i = i0 + ty * TM + ii;
        const DevVariant vi = args.meta[i];
#pragma unroll
        for (int jj = 0; jj < TN; ++jj) {
            const uint32_t j = j0 + 4 * tx + (jj & 3) + 64 * (jj >> 2);
            PairAcc<NP> v;
#pragma unroll
            for (int a = 0; a < NP; ++a)
#pragma unroll
                for (int b = 0; b < NP; ++b) v.v[a][b] = acc[a][b][ii][jj];
            emit_pair<MODE>(args, prm, i, j, vi, v, lane);
        }
    }
}

}  // namespace twkb
