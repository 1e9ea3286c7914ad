// count_umma.cuh -- tensor-core variant of the phased (no missing data) count
// kernel: the all-pairs haplotype co-occurrence n11[i][j] = sum_h a[i][h]*a[j][h]
// is a dense 0/1 contraction, computed as an int8 x int8 -> int32 GEMM on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, accumulator in TMEM).
//
// Replaces (with bit-identical counts) the same reference comparators as
// count_popc.cuh: PhasedListVector / PhasedVectorizedNoMissing,
// lib/ld/ld_engine.cpp:185-267, 636-707; the three other cells follow from the
// allele counts exactly as :244-246.
//
// Operand: the haplotype matrix expanded to one byte per haplotype (0/1),
// row-major [Mpad][Kbytes], Kbytes = 2N rounded up to 128. int8 products
// accumulate exactly in int32 for any N < 2^31.
//
// One kernel: count_umma3_kernel, a persistent CTA pair (cta_group::2) per TPC -- see the
// comment above Umma3Cfg. (Two earlier variants, a single-CTA 128 x 128 kernel and a
// non-persistent CTA-pair kernel, measured 97.6 / 78.7 ms on C2 against 28 ms and were removed.)
#pragma once
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <string>

#include "common.cuh"
#include "count_popc.cuh"

namespace twkb {

// Profiling switches (CountArgs::debug_flags, env TWKB_DEBUG_FLAGS) switch parts of the kernel off and
// make its results INVALID. They exist only in the separate profiling build (libtwkb_prof.so,
// -DTWKB_PROFILING: bench.py's MMA-only ceiling, scripts/ceiling.py); the product library compiles
// them out.
#ifdef TWKB_PROFILING
#define TWKB_DBG(args, bit) (((args).debug_flags & (bit)) != 0u)
#else
#define TWKB_DBG(args, bit) false
#endif

constexpr uint32_t UMMA_TILE_M = 128;
constexpr uint32_t UMMA_TILE_N = 128;
constexpr uint32_t UMMA_BLOCK_K = 128;  // bytes of K per pipeline stage (one 128B swizzle atom)
constexpr uint32_t UMMA_K = 32;         // K of one kind::i8 instruction
constexpr int UMMA_STAGES = 3;
constexpr int UMMA_THREADS = 256;
constexpr uint32_t UMMA_TMEM_COLS = 128;
constexpr uint32_t UMMA_STAGE_BYTES = (UMMA_TILE_M + UMMA_TILE_N) * UMMA_BLOCK_K;
constexpr size_t UMMA_SMEM_BYTES = 1024 /*align*/ + (size_t)UMMA_STAGES * UMMA_STAGE_BYTES + 128 * sizeof(DevVariant) + 256;

struct UmmaOperand {
    bool valid = false;
    uint8_t* d_bytes = nullptr;  // [Mpad][Kbytes]
    size_t capacity = 0;
    uint32_t Kbytes = 0;         // bytes per operand row
    uint32_t Kelems = 0;         // padded haplotypes per row (= Kbytes for int8, 2*Kbytes for e2m1)
    uint32_t Mpad = 0;
    bool fp4 = false;            // e2m1 nibbles instead of int8 bytes
    CUtensorMap tmap;            // box 128 B x 128 rows (A, and B of the int8 kernels)
    CUtensorMap tmap_b;          // box 128 B x B rows of the persistent kernel
    int mode = 0;                // CountMode the operand was built for (planes kernels: 1..3)
    uint8_t* d_bytes_b = nullptr;  // planes kernels: the plane-major B copy [RB][Kbytes]
    size_t capacity_b = 0;
    uint32_t* d_pace = nullptr;    // K-sweep pacing counters of the persistent kernel (long rows only), [waves][chunks per tile]
    size_t pace_capacity = 0;      // in counters
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 inputs, int32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, 128B swizzle (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) =
// 1024 B between 8-row groups | version 1 [46,48) | layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 @ bit 4), A/B unsigned
// 8-bit (0 @ bits 7, 10), both K-major, N>>3 @ bit 17, M>>4 @ bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_i8(uint32_t M, uint32_t N) { return (2u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

// One byte per haplotype from the reference-layout rows (row-major u64 words).
__global__ void expand_bits_to_bytes_kernel(const uint64_t* __restrict__ rows, size_t stride64, uint32_t n_variants,
                                            uint32_t n_bits, uint8_t* __restrict__ out, uint32_t Kbytes, uint32_t Mpad) {
    const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;  // 64-haplotype group
    const uint32_t v = blockIdx.x;
    if (w * 64 >= Kbytes || v >= Mpad) return;
    uint64_t x = 0;
    if (v < n_variants && (size_t)w < stride64 && (uint64_t)w * 64 < n_bits) {
        x = rows[(size_t)v * stride64 + w];
        if ((uint64_t)w * 64 + 64 > n_bits) x &= (1ull << (n_bits - w * 64)) - 1ull;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)v * Kbytes + (size_t)w * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t nib = (uint32_t)(x >> (16 * q + 4 * r)) & 0xFu;
            b[r] = (nib * 0x00204081u) & 0x01010101u;  // bit t of the nibble -> byte t
        }
        dst[q] = make_uint4(b[0], b[1], b[2], b[3]);
    }
}

// =====================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a
// 256 x 256 tile. Each CTA stages only ITS 128 rows of A and ITS 128 rows of B per K
// block (32 KB instead of 64 KB for the same 256x256x128 MACs), the leader CTA issues
// M=256 x N=256 MMAs that read both CTAs' shared memory, and each CTA's TMEM receives
// its 128 accumulator rows x 256 columns. This halves the L2->SM operand traffic per
// MAC, which is what bounds the 1-CTA kernel (ncu: xbar2l1tex at 45 % with the tensor
// pipe at 43 %; profiles/round1_umma_1cta.md).
constexpr uint32_t UMMA2_TILE = 256;
constexpr int UMMA2_STAGES = 3;
constexpr uint32_t UMMA2_TMEM_COLS = 256;
constexpr uint32_t UMMA2_STAGE_BYTES = 2 * 128 * UMMA_BLOCK_K;  // per CTA: 128 A rows + 128 B rows
constexpr size_t UMMA2_SMEM_BYTES = 1024 + (size_t)UMMA2_STAGES * UMMA2_STAGE_BYTES + 256 * sizeof(DevVariant) + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load executed by both CTAs of the pair; the transaction bytes are credited to the
// LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared).
// `l2_policy`: an L2 eviction-priority descriptor (createpolicy encoding; L2_EVICT_* below).
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                uint64_t l2_policy) {
    const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_addr), "r"(c0), "r"(c1), "l"(l2_policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the same-offset mbarrier of both CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// =====================================================================================
// Persistent 2-CTA variant: one CTA pair per TPC loops over 256 x TILE_N tiles. The operand
// ring is 6 stages deep (~190 KB of the SM's shared memory in flight hides the ~1-2 us
// L2 latency that starves the 3-stage kernels), TMEM holds two accumulators, and the
// epilogue warps drain tile n while the tensor pipe already works on tile n+1.
//
// Two operand encodings share the kernel (template parameter FP4):
//   FP4 = false  int8 0/1 operands, tcgen05.mma kind::i8, int32 accumulators, 256x256 tiles;
//   FP4 = true   e2m1 0/1 operands (nibble 0x2 = 1.0), tcgen05.mma kind::mxf4 block-scaled
//                with every UE8M0 scale factor = 2^0, fp32 accumulators. The products are
//                0 or 1 and every partial sum is an integer < 2^24, so the fp32 accumulation
//                is exact (enforced by the host: 2N < 2^24, and proven bit for bit against
//                the POPC kernel by the tests). Twice the MACs per instruction and half the
//                operand bytes of int8. The scale factors occupy TMEM columns [480,512), so
//                the two accumulators are 240 columns wide: 256 x 240 tiles.
template <bool FP4>
struct Umma3Cfg {
    static constexpr uint32_t TILE_N = FP4 ? 240u : 256u;
    static constexpr uint32_t B_ROWS = TILE_N / 2;                          // B rows staged per CTA
    static constexpr uint32_t STAGE_BYTES = (128u + B_ROWS) * UMMA_BLOCK_K; // per CTA
    static constexpr int STAGES = 6;
    static constexpr uint32_t SF_COL = 480u;                                // FP4 only
    static constexpr size_t SMEM_BYTES =
        1024 + (size_t)STAGES * STAGE_BYTES + 2 * 256 * sizeof(DevVariant) + 2 * 256 * sizeof(float2) + 1024 * 16 /*survivor ring*/ + 256;
};
constexpr uint32_t UMMA3_TMEM_COLS = 512;

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t target_cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(target_cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void epilogue_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// One lane of a converged warp (cute::elect_one_sync).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0u;
}

// Block-scaled e2m1 MMA (K = 64 per instruction), scale factors read from TMEM.
__device__ __forceinline__ void umma_mxf4_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                              uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// Instruction descriptor of the block-scaled kinds (cute::UMMA::InstrDescriptorBlockScaled):
// A/B format E2M1 (1 @ bits 7, 10), K-major, N>>3 @ 17, scale format UE8M0 (1 @ 23), M>>4 @ 24,
// scale-factor ids 0, K = 64.
__host__ __device__ constexpr uint32_t umma_idesc_mxf4(uint32_t M, uint32_t N) {
    return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}
// 32 lanes x 32 columns of one constant.
__device__ __forceinline__ void tmem_fill_32x32(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
        "r"(v)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Epilogue of the persistent kernel: 8 warps, two per TMEM lane quarter (a warp may only read
// the quarter warp_id % 4), each draining half of the columns of accumulator n & 1 of tile n
// while the tensor pipe works on tile n + 1.
//
// Fast path, 5 fp32 instructions per pair: with X = n11*T - acA*acB, DA = acA(T-acA),
// DB = acB(T-acB) the pair can only reach R2 >= minR2 if |X| >= sqrt(minR2*DA*DB). Per row
// s_i = sqrt(DA) and per column s_j = sqrt(thr*DB), thr = minR2*(1-1e-12)*(1-2e-5), are
// precomputed; m = max_j(|X| + 1e-6*acA*acB - s_i*s_j) over the 32 columns of a chunk (the 1e-6
// term and the 2e-5 relative slack dominate every fp32 rounding error of the expression, so
// the test is conservative). Only chunks in which some lane has m >= -1 (a few % of the chunks
// at R2 >= 0.1) take the exact per-column path. Invalid rows / columns carry s = +inf
// (-> -inf or NaN, both ignored by fmaxf).
//
// The epilogue is a chain of latencies (metadata loads, TMEM loads, votes), not of issue
// slots, so everything that can be taken off the per-tile critical path is: the metadata of
// tile n + 1 is fetched while tile n is drained, the TMEM load of chunk c + 1 is in flight
// while chunk c is screened, and the release of the accumulator is a CTA-scope arrive.
constexpr int UMMA3_EPI_WARPS = 8;
constexpr int UMMA3_THREADS = 128 + 32 * UMMA3_EPI_WARPS;
constexpr uint32_t UMMA3_PACE_MAX_SPINS = 20000;  // polls of ~100 ns before a CTA pair stops waiting for the others (see the producer warp)
constexpr uint32_t UMMA3_PACE_MIN_KBLOCKS = 128;  // operand rows of >= 16 KB sweep K in paced chunks (count_umma3_kernel)

__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// The "+r" operands tie the wait to the registers of the load it completes.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ DevVariant lds_variant(uint32_t saddr) {
    DevVariant v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.pos), "=r"(v.ac), "=r"(v.rid), "=r"(v.flags) : "r"(saddr));
    return v;
}
// arrive on an mbarrier of another CTA of the cluster (CTA-scope release: no gpu-wide fence)
__device__ __forceinline__ void mbar_arrive_remote_cta(uint64_t* bar, uint32_t target_cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(target_cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void epilogue_bar_sync8() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct EpiRow {  // per-thread row constants of one tile
    uint32_t i;
    DevVariant vi;
    bool i_ok;
    float acA, dA, sA;
};

// Direct path of a 32-column chunk (the instantiation without a screen, minR2 = 0, where every
// pair is a candidate): per-column pair rules, exact decision, then ONE atomic per chunk (warp
// scan) and the candidate stores.
template <bool FP4>
__device__ __forceinline__ void umma_direct_chunk(const CountArgs& args, const DevParams& prm, const uint32_t (&r)[32], const EpiRow& row,
                                                  uint32_t meta_saddr, uint32_t j0, int chunk, float Tf, float thr, bool no_screen,
                                                  int lane) {
    constexpr uint32_t TILE_N = Umma3Cfg<FP4>::TILE_N;
    const uint32_t M = prm.n_variants;
    uint32_t passmask = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const int jl = chunk * 32 + c;
        const uint32_t j = j0 + jl;
        const DevVariant vj = lds_variant(meta_saddr + (uint32_t)jl * 16u);
        bool pass = row.i_ok && j >= args.col_begin && j < args.col_end && j < M && (!prm.diag || row.i < j) && (row.vi.ac + vj.ac > 2);
        if (TILE_N % 32 != 0) pass = pass && (jl < (int)TILE_N);
        if (!no_screen) {
            const float n11 = FP4 ? __uint_as_float(r[c]) : (float)r[c];
            const float acB = (float)vj.ac;
            const float pab = row.acA * acB;
            const float x = fabsf(fmaf(n11, Tf, -pab));
            const float slack = 4.0f + 4.0e-7f * fmaxf(n11 * Tf, pab);
            const float lhs = (x + slack) * (x + slack);
            const float rhs = thr * (row.dA * (acB * (Tf - acB)));
            pass = pass && (lhs >= rhs);
        }
        passmask |= (pass ? 1u : 0u) << c;
    }
    if (!__any_sync(0xffffffffu, passmask != 0u)) return;
    // exact decision for the flagged pairs of this lane (a handful at most)
    uint32_t keep = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if ((passmask >> c) & 1u) {
            const int jl = chunk * 32 + c;
            const DevVariant vj = lds_variant(meta_saddr + (uint32_t)jl * 16u);
            PairAcc<1> pa;
            pa.v[0][0] = FP4 ? (uint32_t)__uint_as_float(r[c]) : r[c];
            uint32_t cc[9], mode = 0;
            if (pair_decide<MODE_PHASED_NOMISS>(args, prm, row.i, j0 + jl, row.vi, vj, pa, cc, mode)) keep |= 1u << c;
        }
    }
    // warp-aggregated compaction: one atomic per chunk
    const uint32_t cnt = __popc(keep);
    uint32_t scan = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, scan, d);
        if (lane >= d) scan += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, scan, 31);
    if (total == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(args.cand_count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0) + (scan - cnt);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if ((keep >> c) & 1u) {
            const unsigned long long slot = base + __popc(keep & ((1u << c) - 1u));
            if (slot < args.cand_capacity) {
                const int jl = chunk * 32 + c;
                const DevVariant vj = lds_variant(meta_saddr + (uint32_t)jl * 16u);
                const uint32_t n11 = FP4 ? (uint32_t)__uint_as_float(r[c]) : r[c];
                uint4* dst = reinterpret_cast<uint4*>(args.cands + slot);
                dst[0] = make_uint4(row.i, j0 + jl, 2u * prm.n_samples - ((row.vi.ac + vj.ac) - n11), row.vi.ac - n11);
                dst[1] = make_uint4(vj.ac - n11, n11, 0u, 0u);
                dst[2] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

// ---- survivor queue of the persistent kernel -------------------------------------------------
// The pairs the fast screen flags are rare (1e-5 of the pairs at C2) but spread over every
// tile, and their exact decision needs global round trips (metadata, the candidate-slot
// atomic): ~2-5 us, a whole tile time, if an epilogue warp does it inline. Instead the epilogue
// warps push (i, j, n11) into a shared-memory ring (~0.3 us) and the otherwise idle warp 3
// drains it: exact pair rules + fp64 screen, warp-aggregated global atomic, candidate stores.
constexpr uint32_t UMMA3_QCAP = 1024;  // entries (16 KB)
struct __align__(16) QEntry { uint32_t i, j, n11, seq; };
struct SurvivorQueue {
    QEntry* ring;        // [UMMA3_QCAP]
    uint32_t* ctrl;      // [0] tail (reservations), [1] head (consumed), [2] finished producer warps
};
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// Reserves one slot per flagged lane (one atomic per warp) and waits, with a condition that is the
// SAME in every lane, until the whole reservation fits in the ring. A per-lane wait would deadlock
// when the ring fills: the lanes whose slots fit would be held at the loop's reconvergence point
// behind a spinning lane, never publish, and the drain warp (which consumes in slot order) would
// never free the slot the spinning lane waits for. Called convergently by all 32 lanes.
__device__ __forceinline__ uint32_t queue_reserve_warp(uint32_t* ctrl, unsigned flagged, uint32_t capacity, int lane) {
    const int leader = __ffs(flagged) - 1;
    const uint32_t cnt = (uint32_t)__popc(flagged);
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&ctrl[0], cnt);
    base = __shfl_sync(0xffffffffu, base, leader);
    for (;;) {  // ring full: wait for the drain warp; lane 0 reads, so the exit is uniform by construction
        uint32_t head = 0;
        if (lane == 0) head = ld_volatile_shared(&ctrl[1]);
        head = __shfl_sync(0xffffffffu, head, 0);
        if ((base + cnt - 1u) - head < capacity) break;
        __nanosleep(64);
    }
    return base + (uint32_t)__popc(flagged & ((1u << lane) - 1u));
}
// Warp-wide push of the flagged lanes' (i, j, n11); every lane of the warp must call it.
__device__ __forceinline__ void queue_push_warp(const SurvivorQueue& q, bool flag, uint32_t i, uint32_t j, uint32_t n11, int lane) {
    const unsigned fl = __ballot_sync(0xffffffffu, flag);
    if (fl == 0) return;
    const uint32_t slot = queue_reserve_warp(q.ctrl, fl, UMMA3_QCAP, lane);
    if (flag) {
        QEntry* e = q.ring + (slot % UMMA3_QCAP);
        e->i = i; e->j = j; e->n11 = n11;
        __threadfence_block();
        st_volatile_shared(&e->seq, slot / UMMA3_QCAP + 1u);  // publishes the entry
    }
    __syncwarp();
}

// One poll of the drain warp. The control words are read by lane 0 and broadcast, and the warp is
// re-converged first: `head`, `tail` and every decision derived from them must be identical in all
// 32 lanes (a lane that read `tail` a few cycles later than its neighbours would consume a
// different number of entries and wait for slots nobody will ever publish). Returns false when
// every producer has finished and the ring is empty; sleeps when there is nothing to do yet.
__device__ __forceinline__ bool drain_poll(uint32_t* ctrl, uint32_t head, uint32_t& tail, int lane) {
    __syncwarp();
    uint32_t t = 0, fin = 0;
    if (lane == 0) {
        t = ld_volatile_shared(&ctrl[0]);
        if (t == head) {
            fin = ld_volatile_shared(&ctrl[2]);
            if (fin == (uint32_t)UMMA3_EPI_WARPS) t = ld_volatile_shared(&ctrl[0]);  // pushes precede the "finished" mark
        }
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    fin = __shfl_sync(0xffffffffu, fin, 0);
    tail = t;
    if (t == head) {
        if (fin == (uint32_t)UMMA3_EPI_WARPS) return false;
        __nanosleep(200);
    }
    return true;
}

// Warp 3: drains the ring until all epilogue warps have finished and the ring is empty.
__device__ __noinline__ void umma_drain_loop(const CountArgs& args, const DevParams& prm, SurvivorQueue q, int lane) {
    const uint32_t M = prm.n_variants;
    uint32_t head = 0;
    for (;;) {
        uint32_t tail;
        if (!drain_poll(q.ctrl, head, tail, lane)) break;
        if (tail == head) continue;
        const uint32_t n = min(tail - head, 32u);
        const bool have = (uint32_t)lane < n;
        uint32_t i = 0, j = 0, n11 = 0;
        if (have) {
            const uint32_t slot = head + (uint32_t)lane;
            QEntry* e = q.ring + (slot % UMMA3_QCAP);
            while (ld_volatile_shared(&e->seq) != slot / UMMA3_QCAP + 1u) __nanosleep(32);  // reserved, not yet written
            __threadfence_block();
            i = e->i; j = e->j; n11 = e->n11;
        }
        __syncwarp();
        head += n;
        __threadfence_block();
        if (lane == 0) st_volatile_shared(&q.ctrl[1], head);  // the slots may be reused
        bool ok = have && i >= args.row_begin && i < args.row_end && j >= args.col_begin && j < args.col_end && i < M && j < M;
        if (prm.diag) ok = ok && (i < j);
        uint32_t c[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) c[k] = 0;
        uint32_t mode = 0;
        if (ok) {
            const DevVariant vi = args.meta[i], vj = args.meta[j];
            PairAcc<1> pa;
            pa.v[0][0] = n11;
            ok = pair_decide<MODE_PHASED_NOMISS>(args, prm, i, j, vi, vj, pa, c, mode);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, ok);
        if (ballot == 0) continue;
        const int leader = __ffs(ballot) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(args.cand_count, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok) {
            const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1));
            if (slot < args.cand_capacity) {
                uint4* dst = reinterpret_cast<uint4*>(args.cands + slot);
                dst[0] = make_uint4(i, j, c[0], c[1]);
                dst[1] = make_uint4(c[2], c[3], 0u, 0u);
                dst[2] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

// One 32-column chunk held in r[]. SCREEN: fast screen, flagged pairs go to the survivor queue.
// !SCREEN (minR2 = 0, nothing can be screened out): the direct path for every chunk.
template <bool FP4, bool SCREEN>
__device__ __forceinline__ void umma_epilogue_chunk(const CountArgs& args, const DevParams& prm, uint32_t (&r)[32], const EpiRow& row,
                                                    uint32_t meta_saddr, uint32_t colf_saddr, uint32_t j0, int chunk, float Tf,
                                                    float thr, int lane, const SurvivorQueue& q) {
    if (SCREEN) {
        float m = __int_as_float(0xff800000);
#ifdef TWKB_PROFILING
        if (TWKB_DBG(args, 32u)) {  // ablation: the screen arithmetic without its shared-memory loads (results invalid)
            const float4 cb = make_float4(row.acA, row.sA, row.dA, row.sA);
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) {
                {
                    const float n11 = FP4 ? __uint_as_float(r[2 * c2]) : (float)r[2 * c2];
                    const float pab = row.acA * (cb.x + (float)c2);
                    const float x = fmaf(n11, Tf, -pab);
                    m = fmaxf(m, fmaf(-row.sA, cb.y, fmaf(pab, 1.0e-6f, fabsf(x))));
                }
                {
                    const float n11 = FP4 ? __uint_as_float(r[2 * c2 + 1]) : (float)r[2 * c2 + 1];
                    const float pab = row.acA * (cb.z + (float)c2);
                    const float x = fmaf(n11, Tf, -pab);
                    m = fmaxf(m, fmaf(-row.sA, cb.w, fmaf(pab, 1.0e-6f, fabsf(x))));
                }
            }
            if (m == 12345.678f) q.ctrl[3] = 1u;  // keeps the arithmetic alive
            return;
        }
        if (TWKB_DBG(args, 64u)) {  // ablation: the shared-memory loads without the arithmetic
            float acc = 0.0f;
#pragma unroll
            for (int c2 = 0; c2 < 16; ++c2) {
                const float4 cb = lds_f4(colf_saddr + (uint32_t)(chunk * 32 + 2 * c2) * 8u);
                acc += cb.x;
            }
            if (acc == 12345.678f) q.ctrl[3] = 1u;
            return;
        }
#endif
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) {
            const float4 cb = lds_f4(colf_saddr + (uint32_t)(chunk * 32 + 2 * c2) * 8u);  // {ac_j, s_j} of two columns
            {
                const float n11 = FP4 ? __uint_as_float(r[2 * c2]) : (float)r[2 * c2];
                const float pab = row.acA * cb.x;
                const float x = fmaf(n11, Tf, -pab);
                m = fmaxf(m, fmaf(-row.sA, cb.y, fmaf(pab, 1.0e-6f, fabsf(x))));
            }
            {
                const float n11 = FP4 ? __uint_as_float(r[2 * c2 + 1]) : (float)r[2 * c2 + 1];
                const float pab = row.acA * cb.z;
                const float x = fmaf(n11, Tf, -pab);
                m = fmaxf(m, fmaf(-row.sA, cb.w, fmaf(pab, 1.0e-6f, fabsf(x))));
            }
        }
        if (!__any_sync(0xffffffffu, m >= -1.0f)) return;
        if (TWKB_DBG(args, 16u)) return;  // screen-only ablation: zero registers would flag everything
        // which pairs: the same margins once more, kept as a bit mask (rare path). Pairs the reference skips for
        // ac_i + ac_j <= 2 (ld_engine.cpp:1918) are dropped here: two singletons on one haplotype have R2 = 1, and real
        // cohorts carry many of them -- they would all travel through the ring only to be rejected by the drain warp.
        uint32_t mask = 0;
#pragma unroll
        for (int c2 = 0; c2 < 16; ++c2) {
            const float4 cb = lds_f4(colf_saddr + (uint32_t)(chunk * 32 + 2 * c2) * 8u);
            {
                const float n11 = FP4 ? __uint_as_float(r[2 * c2]) : (float)r[2 * c2];
                const float pab = row.acA * cb.x;
                const float x = fmaf(n11, Tf, -pab);
                mask |= (fmaf(-row.sA, cb.y, fmaf(pab, 1.0e-6f, fabsf(x))) >= -1.0f && row.acA + cb.x > 2.5f ? 1u : 0u) << (2 * c2);
            }
            {
                const float n11 = FP4 ? __uint_as_float(r[2 * c2 + 1]) : (float)r[2 * c2 + 1];
                const float pab = row.acA * cb.z;
                const float x = fmaf(n11, Tf, -pab);
                mask |= (fmaf(-row.sA, cb.w, fmaf(pab, 1.0e-6f, fabsf(x))) >= -1.0f && row.acA + cb.z > 2.5f ? 1u : 0u) << (2 * c2 + 1);
            }
        }
        if (__any_sync(0xffffffffu, mask != 0u)) {  // rare: some lane flagged a pair of this chunk
#pragma unroll
            for (int c = 0; c < 32; ++c)
                queue_push_warp(q, (mask >> c) & 1u, row.i, j0 + (uint32_t)(chunk * 32 + c),
                                FP4 ? (uint32_t)__uint_as_float(r[c]) : r[c], lane);
        }
    } else {
        umma_direct_chunk<FP4>(args, prm, r, row, meta_saddr, j0, chunk, Tf, thr, true, lane);
    }
}

template <bool FP4, bool SCREEN>
__device__ __forceinline__ void umma_epilogue_loop(const CountArgs& args, const DevParams& prm, DevVariant* s_meta, float2* s_colf,
                                                   uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint32_t tmem_base,
                                                   uint32_t cluster_id, uint32_t n_clusters, uint32_t n_tiles, uint32_t row_off,
                                                   uint32_t leader_cta, int warp, int lane, const SurvivorQueue& queue) {
    constexpr uint32_t TILE_N = Umma3Cfg<FP4>::TILE_N;
    constexpr int N_CHUNKS = (int)((TILE_N + 31) / 32);       // 8
    constexpr int CHUNKS_PER_WARP = N_CHUNKS / (UMMA3_EPI_WARPS / 4);  // 4
    static_assert(CHUNKS_PER_WARP % 2 == 0, "the chunk loop is unrolled by two");
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 4) >> 2;       // which half of the columns
    const int te = threadIdx.x - 128;       // 0..255: column whose metadata this thread prepares
    const uint32_t M = prm.n_variants;
    const float Tf = (float)(2u * prm.n_samples);
    const float thr = (float)prm.screenR2 * (1.0f - 1.0e-5f);
    const float thr_fast = (float)prm.screenR2 * (1.0f - 2.0e-5f);
    const float f_inf = __int_as_float(0x7f800000);
    const uint32_t meta_s0 = smem_u32(s_meta), colf_s0 = smem_u32(s_colf);

    auto store_column = [&](uint32_t buf, uint32_t j0, DevVariant vj) {
        const uint32_t j = j0 + (uint32_t)te;
        s_meta[buf * 256 + te] = vj;
        const bool j_ok = te < (int)TILE_N && j >= args.col_begin && j < args.col_end && j < M;
        const float acB = (float)vj.ac;
        s_colf[buf * 256 + te] = make_float2(acB, j_ok ? sqrtf(thr_fast * (acB * (Tf - acB))) : f_inf);
    };
    auto load_column = [&](uint32_t j0) {
        const uint32_t j = j0 + (uint32_t)te;
        return j < args.Mpad ? args.meta[j] : DevVariant{0, 0, 0, 0};
    };

    uint32_t t = cluster_id;
    uint2 tile = make_uint2(0, 0);
    DevVariant vi_cur{0, 0, 0, 0};
    if (t < n_tiles) {
        tile = args.tiles[t];
        store_column(0, tile.y, load_column(tile.y));
        vi_cur = args.meta[tile.x + row_off + 32 * q + lane];
    }
    for (uint32_t n = 0; t < n_tiles; t += n_clusters, ++n) {
        const uint32_t acc = n & 1;
        const uint32_t j0 = tile.y;
        EpiRow row;
        row.i = tile.x + row_off + 32 * q + lane;
        row.vi = vi_cur;
        row.i_ok = row.i >= args.row_begin && row.i < args.row_end && row.i < M;
        row.acA = (float)row.vi.ac;
        row.dA = row.acA * (Tf - row.acA);
        row.sA = row.i_ok ? sqrtf(row.dA) : f_inf;
        // buffer `acc` (written during the previous iteration) becomes visible; every warp is
        // done with tile n - 1, so buffer acc ^ 1 may be rewritten below
        epilogue_bar_sync8();
        // metadata of the next tile: the loads stay in flight across the drain of this one
        const uint32_t t_next = t + n_clusters;
        const bool has_next = t_next < n_tiles;
        uint2 tile_next = make_uint2(0, 0);
        DevVariant vj_next{0, 0, 0, 0}, vi_next{0, 0, 0, 0};
        if (has_next) {
            tile_next = args.tiles[t_next];
            vj_next = load_column(tile_next.y);
            vi_next = args.meta[tile_next.x + row_off + 32 * q + lane];
        }
        mbar_wait(&tmem_full_bar[acc], (n >> 1) & 1);
        tcgen05_fence_after();
        if (!TWKB_DBG(args, 2u)) {
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * TILE_N;
            const uint32_t meta_sa = meta_s0 + acc * 256u * 16u, colf_sa = colf_s0 + acc * 256u * 8u;
            const int c_begin = half * CHUNKS_PER_WARP;
            uint32_t r0[32], r1[32];
#ifdef TWKB_PROFILING
            // ablations of the epilogue (results invalid): 8 = TMEM loads only (no screen), 16 = screen only
            // (no TMEM loads: the registers stay zero, the rare flagged path is cut short)
            const bool do_ld = !TWKB_DBG(args, 16u), do_math = !TWKB_DBG(args, 8u);
#pragma unroll
            for (int z = 0; z < 32; ++z) r0[z] = r1[z] = 0u;
#else
            constexpr bool do_ld = true, do_math = true;
#endif
            // (FP4: the last chunk reads 16 columns past the accumulator; they carry s_j = +inf)
            if (do_ld) {
                tmem_ld_32x32_nowait(taddr + (uint32_t)(c_begin * 32), r0);
                tmem_ld_wait(r0);
            }
#pragma unroll 1
            for (int c = c_begin; c < c_begin + CHUNKS_PER_WARP; c += 2) {
                if (do_ld) tmem_ld_32x32_nowait(taddr + (uint32_t)((c + 1) * 32), r1);
                if (do_math) umma_epilogue_chunk<FP4, SCREEN>(args, prm, r0, row, meta_sa, colf_sa, j0, c, Tf, thr, lane, queue);
                if (do_ld) tmem_ld_wait(r1);
                if (do_ld && c + 2 < c_begin + CHUNKS_PER_WARP) tmem_ld_32x32_nowait(taddr + (uint32_t)((c + 2) * 32), r0);
                if (do_math) umma_epilogue_chunk<FP4, SCREEN>(args, prm, r1, row, meta_sa, colf_sa, j0, c + 1, Tf, thr, lane, queue);
                if (do_ld) tmem_ld_wait(r0);
            }
        }
        // this warp has read everything it needs from accumulator `acc`
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_cta(&tmem_empty_bar[acc], leader_cta);
        if (has_next) store_column(acc ^ 1u, tile_next.y, vj_next);
        tile = tile_next;
        vi_cur = vi_next;
    }
    // every push of this warp is written; tell the drain warp
    __syncwarp();
    __threadfence_block();
    if (lane == 0) atomicAdd(&queue.ctrl[2], 1u);
}

// ---- multi-plane tables on the tensor pipe (masked phased 2x2, unphased 3x3) ----------------
// The NP^2 joint plane counts of a pair (count_popc.cuh: PopcCfg<MODE>::NP planes per variant) are
// NP^2 entries of the SAME product A.B^T when every variant contributes NP operand rows, one per
// plane ([alt&valid; valid], [het; hom], [het&valid; hom&valid; valid] -- SURVEY.md 8d "GEMM view").
// The MMA pipeline is the e2m1 persistent kernel unchanged (256 x 240 tiles); only the row order of
// the two operand copies and the epilogue differ:
//   A copy: 32-row groups of G = 32/NP variants, row = 32*(v/G) + NP*(v%G) + plane (NP = 3 leaves two
//           zero rows per group), so the NP accumulator rows of a variant are NP adjacent TMEM
//           lanes of one warp: a tile holds TI = 8*G variants (128 / 80);
//   B copy: plane-major per tile of NV = 240/NP variants, row = 240*(v/NV) + NV*plane + v%NV, so a
//           thread reads the NP column planes of 8 consecutive variants with NP tcgen05.ld.x8.
// Epilogue: lane (variant vl, plane pa) gathers the NP x NP table of (vl, column c) with NP^2 warp
// shuffles per column; the lane with pa == c % NP runs the exact per-pair rules + fp64 screen of
// the POPC kernel (pair_decide) and survivors are appended with one atomic per warp ballot. The
// tensor work per pair is NP^2 times that of the 1-plane kernel, so this epilogue is hidden.
template <int MODE>
struct PlanesCfg {
    static constexpr int NP = PopcCfg<MODE>::NP;
    static constexpr uint32_t G = 32u / (NP > 0 ? NP : 1);
    static constexpr uint32_t TI = 8u * G;
    static constexpr uint32_t NV = 240u / (NP > 0 ? NP : 1);
    static constexpr uint32_t TJ = NV;
    static constexpr int CW = 8;
    static constexpr int N_CHUNKS = (int)(NV / CW);
    static_assert(NV % CW == 0, "column chunks");
};

__device__ __forceinline__ void tmem_ld_32x32_x8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}

// x[pl] for a lane-constant pl: NP - 1 SEL instructions on predicates that are loop invariant (is0 =
// pl == 0, is1 = pl == 1). Written as selp so that the compiler cannot turn the lane-varying choice
// into divergent branches (ncu of the first version: 60 BRA + 42 BSSY/BSYNC pairs per 8-column
// chunk, a third of the epilogue's issue slots).
__device__ __forceinline__ uint32_t selp_u32(uint32_t a, uint32_t b, uint32_t pred) {
    uint32_t d;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(pred));
    return d;
}
template <int NP>
__device__ __forceinline__ uint32_t sel_by_lane(uint32_t is0, uint32_t is1, const uint32_t (&x)[NP]) {
    if (NP == 2) return selp_u32(x[0], x[1], is0);
    return selp_u32(x[0], selp_u32(x[1], x[NP - 1], is1), is0);
}

// Conservative fp32 form of the exact fp64 screens of count_popc.cuh (screen_phased /
// screen_unphased), division free: every quantity is scaled by T^2 (phased) or (2T)^2 (unphased),
// the operands are integers < 2^25 and every rounding error (<= 2^-24 relative per operation) is
// covered by the explicit slack terms, so a pair the exact screen keeps is always flagged. Flagged
// pairs (~0.5 % at R2 >= 0.1) then take the exact path.
// The table entries arrive as the fp32 accumulator values (exact integers < 2^24) and every sum
// below stays an exact integer in fp32, so nothing is converted unless the pair is flagged.
template <int MODE>
__device__ __forceinline__ bool planes_fast_screen(const float (&v)[PopcCfg<MODE>::NP][PopcCfg<MODE>::NP], float n_samples, float hetA,
                                                   float homA, float hetB, float homB, float thr) {
    constexpr int NP = PopcCfg<MODE>::NP;
    if (MODE == MODE_PHASED_MISS) {
        const float n11 = v[0][0], nA = v[0][1], nB = v[1][0], nV = v[1][1];
        const float p1 = n11 * nV, p2 = nA * nB;
        const float x = fabsf(p1 - p2) + (2.5e-7f * (p1 + p2) + 1.0f);
        const float den = (nA * (nV - nA)) * (nB * (nV - nB));  // nV - nA: exact integers
        return (x * x) * (1.0f + 1.0e-5f) >= thr * den;
    } else {
        float hA, oA, hB, oB, vv;
        if (MODE == MODE_UNPHASED_NOMISS) { hA = hetA; oA = homA; hB = hetB; oB = homB; vv = n_samples; }
        else { hA = v[0][NP - 1]; oA = v[1 % NP][NP - 1]; hB = v[NP - 1][0]; oB = v[NP - 1][1 % NP]; vv = v[NP - 1][NP - 1]; }
        const float c11 = v[0][0], c12 = v[0][1 % NP], c21 = v[1 % NP][0], c22 = v[1 % NP][1 % NP];
        // t0 = both 0/0, t1 = (0/0, het), t3 = (het, 0/0), t4 = (het, het); n11 = 2 t0 + t1 + t3
        // (partial sums are integers in [-2 vv, 2 vv]: exact)
        const float t0 = ((vv - hA - oA) - (hB + oB)) + ((c11 + c12) + (c21 + c22));
        const float t1 = hB - c11 - c21, t3 = hA - c11 - c12;
        const float n11 = 2.0f * t0 + t1 + t3;
        const float S = 2.0f * vv;
        const float a = S - hA - 2.0f * oA, c = S - hB - 2.0f * oB;  // 2T * P, 2T * Q
        const float pq = a * c, eps = 1.0e-5f * (S * S);
        const float l1 = n11 * S, h1 = (n11 + c11) * S;
        const float lo = (l1 - pq) - eps, hi = (h1 - pq) + eps;
        const float dmax = fmaxf(fabsf(lo), fabsf(hi)) + (2.5e-7f * (h1 + pq) + 3.0e-12f * (S * S) + 1.0f);
        const float den = (a * (S - a)) * (c * (S - c));
        return (dmax * dmax) * (1.0f + 1.0e-5f) >= thr * den;
    }
}

// Exact path of one round: kept out of line (long fp64 code, rarely taken).
template <int MODE>
__device__ __noinline__ void umma_planes_exact(const CountArgs& args, const DevParams& prm, uint32_t i, uint32_t j, DevVariant vi,
                                               DevVariant vj, PairAcc<PopcCfg<MODE>::NP> pa, int lane, bool active) {
    emit_pair_with<MODE>(args, prm, i, j, vi, vj, pa, lane, active);
}

// ---- survivor queue of the planes kernels --------------------------------------------------------
// Same idea as the 1-plane kernel's queue, with the pair's NP x NP plane counts as payload (48-byte
// entries in the same 16 KB ring). With the exact decision inline, an epilogue warp ran the long
// fp64 path in ~1/4 of its rounds whenever ANY of its 30 pairs was flagged (C3: 0.5 % of the
// pairs), the 8 warps of a tile finished far apart and met at the per-tile barrier (ncu: 41 % of
// the stall cycles at that barrier, 17 % of the samples inside the fp64 code). Now the epilogue
// warps only push; warp 3 applies pair rules + exact fp64 screen and appends the candidates.
struct __align__(16) PQEntry { uint32_t i, j, v[9], seq; };
static_assert(sizeof(PQEntry) == 48, "planes queue entry");
constexpr uint32_t UMMA3_PQCAP = (UMMA3_QCAP * (uint32_t)sizeof(QEntry)) / (uint32_t)sizeof(PQEntry);  // 341

// Warp-aggregated push of the flagged lanes' pairs.
template <int NP>
__device__ __forceinline__ void planes_queue_push(PQEntry* ring, uint32_t* ctrl, bool flag, uint32_t i, uint32_t j,
                                                  const float (&tv)[NP][NP], int lane) {
    const unsigned fl = __ballot_sync(0xffffffffu, flag);
    if (fl == 0) return;
    const uint32_t slot = queue_reserve_warp(ctrl, fl, UMMA3_PQCAP, lane);
    if (flag) {
        PQEntry* e = ring + (slot % UMMA3_PQCAP);
        e->i = i; e->j = j;
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int b = 0; b < NP; ++b) e->v[a * NP + b] = (uint32_t)tv[a][b];
        __threadfence_block();
        st_volatile_shared(&e->seq, slot / UMMA3_PQCAP + 1u);  // publishes the entry
    }
    __syncwarp();
}

// Warp 3 of the planes kernels: drains the ring until all epilogue warps have finished.
template <int MODE>
__device__ __noinline__ void umma_planes_drain_loop(const CountArgs& args, const DevParams& prm, PQEntry* ring, uint32_t* ctrl, int lane) {
    constexpr int NP = PopcCfg<MODE>::NP;
    uint32_t head = 0;
    for (;;) {
        uint32_t tail;
        if (!drain_poll(ctrl, head, tail, lane)) break;
        if (tail == head) continue;
        const uint32_t n = min(tail - head, 32u);
        const bool have = (uint32_t)lane < n;
        uint32_t i = 0, j = 0;
        PairAcc<NP> pa;
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int b = 0; b < NP; ++b) pa.v[a][b] = 0;
        if (have) {
            const uint32_t slot = head + (uint32_t)lane;
            PQEntry* e = ring + (slot % UMMA3_PQCAP);
            while (ld_volatile_shared(&e->seq) != slot / UMMA3_PQCAP + 1u) __nanosleep(32);  // reserved, not yet written
            __threadfence_block();
            i = e->i; j = e->j;
#pragma unroll
            for (int a = 0; a < NP; ++a)
#pragma unroll
                for (int b = 0; b < NP; ++b) pa.v[a][b] = e->v[a * NP + b];
        }
        __syncwarp();
        head += n;
        __threadfence_block();
        if (lane == 0) st_volatile_shared(&ctrl[1], head);  // the slots may be reused
        const bool in = have && i < prm.n_variants && j < prm.n_variants;
        DevVariant vi{0, 0, 0, 0}, vj{0, 0, 0, 0};
        if (in) { vi = args.meta[i]; vj = args.meta[j]; }
        emit_pair_with<MODE>(args, prm, i, j, vi, vj, pa, lane, in);
    }
}

template <int MODE>
__device__ __forceinline__ void umma_planes_epilogue_loop(const CountArgs& args, const DevParams& prm, DevVariant* s_meta, uint2* s_colx,
                                                          uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint32_t tmem_base,
                                                          uint32_t cluster_id, uint32_t n_clusters, uint32_t n_tiles, uint32_t rank,
                                                          uint32_t leader_cta, int warp, int lane, PQEntry* q_ring, uint32_t* q_ctrl) {
    using PC = PlanesCfg<MODE>;
    constexpr int NP = PC::NP;
    constexpr int CW = PC::CW;
    constexpr int ROUNDS = (CW + NP - 1) / NP;
    constexpr uint32_t TILE_N = 240u;
    const int q = warp & 3;            // TMEM lane quarter
    const int half = (warp - 4) >> 2;  // which half of the column chunks
    const int te = threadIdx.x - 128;  // 0..255: column whose metadata this thread stages
    const uint32_t vl = (uint32_t)lane / NP, pl = (uint32_t)lane % NP;
    const bool lane_ok = vl < PC::G;
    const int base_lane = lane - (int)pl;
    const uint32_t is0 = pl == 0 ? 1u : 0u, is1 = pl == 1 ? 1u : 0u;
    int src_lane[NP];  // shuffle step s reads from the lane holding plane row (pl + s) % NP of this variant
#pragma unroll
    for (int sft = 0; sft < NP; ++sft) src_lane[sft] = base_lane + (int)((pl + (uint32_t)sft) % NP);
    const float ns_f = (float)prm.n_samples;
    const uint32_t meta_s0 = smem_u32(s_meta);
    const int c_begin = half ? (PC::N_CHUNKS + 1) / 2 : 0;
    const int c_end = half ? PC::N_CHUNKS : (PC::N_CHUNKS + 1) / 2;
    const bool no_screen = args.screen_off || !(prm.minR2 > 0.0);
    const float thr = (float)prm.minR2 * (1.0f - 1.0e-5f);
    const uint32_t* pp = args.plane_popc;

    struct Col { DevVariant v; uint2 x; };
    auto load_column = [&](uint32_t j0) {
        const uint32_t j = j0 + (uint32_t)te;
        Col c{DevVariant{0, 0, 0, 0}, make_uint2(0, 0)};
        if (te < (int)PC::NV && j < args.Mpad) {
            c.v = args.meta[j];
            if (MODE == MODE_UNPHASED_NOMISS) c.x = make_uint2(pp[j], pp[args.Mpad + j]);
        }
        return c;
    };
    auto store_column = [&](uint32_t buf, const Col& c) {
        s_meta[buf * 256u + te] = c.v;
        if (MODE == MODE_UNPHASED_NOMISS) s_colx[buf * 256u + te] = c.x;
    };
    struct Row { DevVariant v; uint32_t het, hom; };
    auto load_row = [&](uint32_t i0) {
        const uint32_t i = i0 + (4u * rank + (uint32_t)q) * PC::G + vl;
        Row r{DevVariant{0, 0, 0, 0}, 0, 0};
        if (lane_ok && i < args.Mpad) {
            r.v = args.meta[i];
            if (MODE == MODE_UNPHASED_NOMISS) { r.het = pp[i]; r.hom = pp[args.Mpad + i]; }
        }
        return r;
    };

    uint32_t t = cluster_id;
    uint2 tile = make_uint2(0, 0);
    Row row_cur{DevVariant{0, 0, 0, 0}, 0, 0};
    if (t < n_tiles) {
        tile = args.tiles[t];
        store_column(0, load_column(tile.y));
        row_cur = load_row(tile.x);
    }
    for (uint32_t n = 0; t < n_tiles; t += n_clusters, ++n) {
        const uint32_t acc = n & 1;
        const uint32_t j0 = tile.y;
        const uint32_t i = tile.x + (4u * rank + (uint32_t)q) * PC::G + vl;
        const Row row = row_cur;
        const float row_het_f = (float)row.het, row_hom_f = (float)row.hom;
        epilogue_bar_sync8();  // metadata buffer `acc` visible; buffer acc ^ 1 free
        const uint32_t t_next = t + n_clusters;
        const bool has_next = t_next < n_tiles;
        uint2 tile_next = make_uint2(0, 0);
        Col col_next{DevVariant{0, 0, 0, 0}, make_uint2(0, 0)};
        Row row_next{DevVariant{0, 0, 0, 0}, 0, 0};
        if (has_next) {
            tile_next = args.tiles[t_next];
            col_next = load_column(tile_next.y);
            row_next = load_row(tile_next.x);
        }
        mbar_wait(&tmem_full_bar[acc], (n >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * TILE_N;
        const uint32_t meta_sa = meta_s0 + acc * 256u * 16u;
        const uint2* colx = s_colx + acc * 256u;
        uint32_t r[NP][CW], rn[NP][CW];
#pragma unroll
        for (int pb = 0; pb < NP; ++pb) tmem_ld_32x32_x8_nowait(taddr + (uint32_t)pb * PC::NV + (uint32_t)(c_begin * CW), r[pb]);
#pragma unroll 1
        for (int ch = c_begin; ch < c_end; ++ch) {
#pragma unroll
            for (int pb = 0; pb < NP; ++pb) tmem_ld_wait8(r[pb]);
            if (ch + 1 < c_end) {  // next chunk in flight while this one is screened
#pragma unroll
                for (int pb = 0; pb < NP; ++pb) tmem_ld_32x32_x8_nowait(taddr + (uint32_t)pb * PC::NV + (uint32_t)((ch + 1) * CW), rn[pb]);
            }
#pragma unroll
            for (int g = 0; g < ROUNDS; ++g) {
                // lane (vl, pl) takes column NP*g + pl: step s moves, from the lane holding plane row
                // a = (pl + s) % NP of this variant, its NP column-plane values of that column
                uint32_t recv[NP][NP];
#pragma unroll
                for (int sft = 0; sft < NP; ++sft) {
#pragma unroll
                    for (int b = 0; b < NP; ++b) {
                        uint32_t cand[NP];  // what this lane sends to destination plane-lane d = (pl - sft) mod NP
#pragma unroll
                        for (int d = 0; d < NP; ++d) {
                            const int col = NP * g + d;
                            cand[(d + sft) % NP] = r[b][col < CW ? col : CW - 1];
                        }
                        recv[sft][b] = __shfl_sync(0xffffffffu, sel_by_lane<NP>(is0, is1, cand), src_lane[sft]);
                    }
                }
                float tv[NP][NP];
#pragma unroll
                for (int a = 0; a < NP; ++a) {
#pragma unroll
                    for (int b = 0; b < NP; ++b) {
                        uint32_t cand[NP];  // recv[(a - pl) mod NP][b]
#pragma unroll
                        for (int d = 0; d < NP; ++d) cand[d] = recv[(a - d + NP) % NP][b];
                        tv[a][b] = __uint_as_float(sel_by_lane<NP>(is0, is1, cand));
                    }
                }
                const uint32_t cl = (uint32_t)(NP * g) + pl;  // column of this lane inside the chunk
                const bool active = lane_ok && cl < (uint32_t)CW;
                const uint32_t jl = (uint32_t)(ch * CW) + (active ? cl : 0u);
                bool flag = active;
                if (!no_screen) {
                    float cxh = 0.0f, cxo = 0.0f;
                    if (MODE == MODE_UNPHASED_NOMISS) { const uint2 cx = colx[jl]; cxh = (float)cx.x; cxo = (float)cx.y; }
                    flag = active && planes_fast_screen<MODE>(tv, ns_f, row_het_f, row_hom_f, cxh, cxo, thr);
                }
                if (!no_screen && !TWKB_DBG(args, 4u)) {  // (flag 4: A/B aid, exact decision inline)
                    // flagged pairs (a fraction of a percent) go to the drain warp
                    planes_queue_push<NP>(q_ring, q_ctrl, flag, i, j0 + jl, tv, lane);
                } else if (__any_sync(0xffffffffu, flag)) {
                    // nothing can be screened out: every pair is a candidate, decided inline by all 8 warps
                    PairAcc<NP> pa;
#pragma unroll
                    for (int a = 0; a < NP; ++a)
#pragma unroll
                        for (int b = 0; b < NP; ++b) pa.v[a][b] = (uint32_t)tv[a][b];
                    const DevVariant vj = lds_variant(meta_sa + jl * 16u);
                    umma_planes_exact<MODE>(args, prm, i, j0 + jl, row.v, vj, pa, lane, flag);
                }
            }
            if (ch + 1 < c_end) {
#pragma unroll
                for (int pb = 0; pb < NP; ++pb)
#pragma unroll
                    for (int c = 0; c < CW; ++c) r[pb][c] = rn[pb][c];
            }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_cta(&tmem_empty_bar[acc], leader_cta);
        if (has_next) store_column(acc ^ 1u, col_next);
        tile = tile_next;
        row_cur = row_next;
    }
    // every push of this warp is written; tell the drain warp
    __syncwarp();
    __threadfence_block();
    if (lane == 0) atomicAdd(&q_ctrl[2], 1u);
}

template <bool FP4, bool SCREEN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UMMA3_THREADS, 1)
count_umma3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, CountArgs args,
                   DevParams prm, uint32_t num_kblocks, uint32_t n_tiles, uint32_t* pace, uint32_t pace_kb, uint32_t pace_depth,
                   uint64_t l2_policy_a, uint64_t l2_policy_b) {
    using Cfg = Umma3Cfg<FP4>;
    constexpr uint32_t TILE_N = Cfg::TILE_N;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    DevVariant* s_meta = reinterpret_cast<DevVariant*>(smem + (size_t)STAGES * Cfg::STAGE_BYTES);  // [2][256]
    float2* s_colf = reinterpret_cast<float2*>(s_meta + 2 * 256);                                  // [2][256] {ac_j, s_j}
    QEntry* s_ring = reinterpret_cast<QEntry*>(s_colf + 2 * 256);                                  // [UMMA3_QCAP]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ring + UMMA3_QCAP);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2], the leader's copy is the one used
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint32_t* q_ctrl = tmem_slot + 4;              // [4] tail, head, finished producers (own 16-byte line: tcgen05.alloc writes tmem_slot)
    const SurvivorQueue queue{s_ring, q_ctrl};

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 2 * UMMA3_EPI_WARPS);  // epilogue warps of both CTAs
        }
        mbar_fence_init();
        q_ctrl[0] = 0; q_ctrl[1] = 0; q_ctrl[2] = 0;
    }
    for (uint32_t e = threadIdx.x; e < UMMA3_QCAP * 4u; e += blockDim.x) reinterpret_cast<uint32_t*>(s_ring)[e] = 0;  // seq = 0 in either entry layout
    if (warp == 2) tmem_alloc_2sm(tmem_slot, UMMA3_TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (FP4) {
        // every scale factor = UE8M0 127 = 2^0, for all 128 lanes of both CTAs
        if (warp >= 4 && warp < 8) tmem_fill_32x32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + Cfg::SF_COL, 0x7F7F7F7Fu);
        tcgen05_fence_before();
        cluster_sync_all();
        tcgen05_fence_after();
    }

    // The two mainloop warps run CONVERGED (all 32 lanes execute the loops, one elected lane issues the
    // asynchronous instructions): with a single active lane (`if (lane == 0)`) ptxas cannot keep the
    // descriptors, barrier addresses and coordinates in uniform registers and wraps every UTCOMMA / UTMALDG /
    // UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop -- ~110 issue slots per K block on the scheduler the
    // issuer shares with two epilogue warps, which is what made the kernel 14 % slower with the (cheap) screen
    // arithmetic running than without it (profiles/round2_epilogue_ablation.log). Stage index and barrier
    // phase are carried as counters (no division by the ring depth).
    if (warp == 0) {
        // ============================== TMA producer ==============================
        uint32_t s = 0, wrap = 0;  // ring slot, number of completed passes over the ring
        // K-sweep pacing (long operand rows, pace != nullptr). The ~74 tiles in flight together (one per CTA pair, a
        // compact block of the super-tile order) read the same operand rows; whether a row's K block is fetched from
        // DRAM once or once per tile depends on the CTA pairs sweeping K in near lock-step, and at 1 M haplotypes
        // (500 KB per row, ~1 ms per tile) they drift apart by far more than L2 holds -- ncu: L2 hit 50 %, DRAM
        // 5.9 TB/s, tensor pipe half idle. So the leader's producer announces every chunk of pace_kb K blocks it
        // starts in a global counter (one per wave and chunk) and does not start chunk c before ALL pairs of the
        // wave have started chunk c - pace_depth: the spread of the wave's K positions -- and with it the L2
        // working set, rows of the wave x (pace_depth + 1) chunks -- is bounded. Pure pacing: no data depends on
        // it and the slowest pair never waits. The wait is BOUNDED: the grid is sized to be co-resident (<= one CTA per
        // SM), but when other work holds SMs -- a second context's persistent kernel on the same device, MPS -- part of
        // the grid may not be resident, and pairs spinning for it would never free the SMs it needs. A pair whose wait
        // exceeds UMMA3_PACE_MAX_SPINS polls (~2 ms) stops waiting for the rest of the launch (it keeps announcing its
        // own progress), so the worst case is an un-paced kernel, never a hang.
        const uint32_t pace_cpt = pace ? (num_kblocks + pace_kb - 1u) / pace_kb : 0u;  // chunks per tile
        bool pace_wait = true;
        uint32_t wave = 0;
        for (uint32_t t = cluster_id; t < n_tiles; t += n_clusters, ++wave) {
            const uint2 tile = args.tiles[t];
            // operand rows of this CTA's halves of the tile (planes: the re-ordered copies)
            const uint32_t a_row = (MODE == MODE_PHASED_NOMISS ? tile.x : (tile.x / PlanesCfg<MODE>::TI) * 256u) + 128u * rank;
            const uint32_t b_row = (MODE == MODE_PHASED_NOMISS ? tile.y : (tile.y / PlanesCfg<MODE>::TJ) * 240u) + Cfg::B_ROWS * rank;
            uint32_t pace_next = 0, pace_chunk = 0;
            for (uint32_t kb = 0; kb < num_kblocks; ++kb) {
                if (pace != nullptr && leader && kb == pace_next) {
                    bool gave_up = false;
                    if (elect_one_sync()) {
                        const uint32_t seq = wave * pace_cpt + pace_chunk;
                        atomicAdd(&pace[seq], 1u);
                        if (pace_wait && seq >= pace_depth) {
                            // wave of chunk seq - pace_depth (one chunk per tile: any depth; else pace_depth <= pace_cpt)
                            const uint32_t w_back = pace_cpt == 1u ? seq - pace_depth : (pace_chunk >= pace_depth ? wave : wave - 1u);
                            const uint32_t expect = min(n_clusters, n_tiles - w_back * n_clusters);
                            const volatile uint32_t* c = pace + (seq - pace_depth);
                            uint32_t spins = 0;
                            while (*c < expect) {
                                __nanosleep(64);
                                if (++spins > UMMA3_PACE_MAX_SPINS) { gave_up = true; break; }
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, gave_up)) pace_wait = false;
                    pace_next += pace_kb;
                    ++pace_chunk;
                }
                if (wrap) mbar_wait(&empty_bar[s], (wrap - 1u) & 1u);
                uint8_t* sA = stage_base + (size_t)s * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + 128 * UMMA_BLOCK_K;
                const bool skip_loads = TWKB_DBG(args, 1u) && wrap;  // profiling aid: no operand traffic after the first ring fill
                if (elect_one_sync()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], skip_loads ? 0u : 2u * Cfg::STAGE_BYTES);
                    if (!skip_loads) {
                        tma_load_2d_2sm(sA, &tmap_a, &full_bar[s], (int32_t)(kb * UMMA_BLOCK_K), (int32_t)a_row, l2_policy_a);
                        tma_load_2d_2sm(sB, &tmap_b, &full_bar[s], (int32_t)(kb * UMMA_BLOCK_K), (int32_t)b_row, l2_policy_b);
                    }
                }
                __syncwarp();
                if (++s == (uint32_t)STAGES) { s = 0; ++wrap; }
            }
        }
    } else if (warp == 1) {
        // ========================= MMA issuer (leader only) =========================
        if (leader) {
            constexpr uint32_t idesc = FP4 ? umma_idesc_mxf4(256, TILE_N) : umma_idesc_i8(256, TILE_N);
            uint32_t s = 0, phase = 0, n = 0;
            for (uint32_t t = cluster_id; t < n_tiles; t += n_clusters, ++n) {
                const uint32_t acc = n & 1;
                if (n >= 2) mbar_wait(&tmem_empty_bar[acc], ((n >> 1) - 1) & 1);  // epilogue drained tile n-2
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TILE_N;
                for (uint32_t kb = 0; kb < num_kblocks; ++kb) {
                    mbar_wait(&full_bar[s], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base) + s * Cfg::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + 128 * UMMA_BLOCK_K;
                    const uint64_t adesc = umma_smem_desc(a_addr), bdesc = umma_smem_desc(b_addr);
                    if (elect_one_sync()) {
                        // 32 bytes of K per instruction in both encodings (32 int8 / 64 e2m1)
#pragma unroll
                        for (uint32_t k = 0; k < UMMA_BLOCK_K / UMMA_K; ++k) {
                            if (FP4)
                                umma_mxf4_2sm(d_tmem, adesc + (uint64_t)(k * UMMA_K >> 4), bdesc + (uint64_t)(k * UMMA_K >> 4), idesc,
                                              (kb | k) != 0 ? 1u : 0u, tmem_base + Cfg::SF_COL, tmem_base + Cfg::SF_COL + 8);
                            else
                                umma_i8_2sm(d_tmem, adesc + (uint64_t)(k * UMMA_K >> 4), bdesc + (uint64_t)(k * UMMA_K >> 4), idesc,
                                            (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_2sm(&empty_bar[s]);
                        if (kb + 1 == num_kblocks) umma_commit_2sm(&tmem_full_bar[acc]);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)STAGES) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        if constexpr (MODE == MODE_PHASED_NOMISS)
            umma_epilogue_loop<FP4, SCREEN>(args, prm, s_meta, s_colf, tmem_full_bar, tmem_empty_bar, tmem_base, cluster_id, n_clusters,
                                            n_tiles, 128u * rank, 0u, warp, lane, queue);
        else
            umma_planes_epilogue_loop<MODE>(args, prm, s_meta, reinterpret_cast<uint2*>(s_colf), tmem_full_bar, tmem_empty_bar, tmem_base,
                                            cluster_id, n_clusters, n_tiles, rank, 0u, warp, lane, reinterpret_cast<PQEntry*>(s_ring), q_ctrl);
    } else if (warp == 3) {
        if constexpr (MODE == MODE_PHASED_NOMISS) umma_drain_loop(args, prm, queue, lane);
        else umma_planes_drain_loop<MODE>(args, prm, reinterpret_cast<PQEntry*>(s_ring), q_ctrl, lane);
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, UMMA3_TMEM_COLS);
    }
}

// One e2m1 nibble (0x2 = 1.0, 0x0 = 0.0) per haplotype from the reference-layout rows.
__global__ void expand_bits_to_e2m1_kernel(const uint64_t* __restrict__ rows, size_t stride64, uint32_t n_variants,
                                           uint32_t n_bits, uint8_t* __restrict__ out, uint32_t Kbytes, uint32_t Mpad) {
    const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;  // 64-haplotype group = 32 output bytes
    const uint32_t v = blockIdx.x;
    if (w * 32 >= Kbytes || v >= Mpad) return;
    uint64_t x = 0;
    if (v < n_variants && (size_t)w < stride64 && (uint64_t)w * 64 < n_bits) {
        x = rows[(size_t)v * stride64 + w];
        if ((uint64_t)w * 64 + 64 > n_bits) x &= (1ull << (n_bits - w * 64)) - 1ull;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)v * Kbytes + (size_t)w * 32);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t y = (uint32_t)(x >> (32 * q + 8 * r)) & 0xFFu;  // bit t -> nibble t
            y = (y | (y << 12)) & 0x000F000Fu;
            y = (y | (y << 6)) & 0x03030303u;
            y = (y | (y << 3)) & 0x11111111u;
            b[r] = y << 1;
        }
        dst[q] = make_uint4(b[0], b[1], b[2], b[3]);
    }
}

// ------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

inline bool umma_supported() {
    if (const char* e = getenv("TWKB_DISABLE_UMMA")) {
        if (e[0] == '1') return false;
    }
    return get_tmap_encoder() != nullptr;
}

// e2m1 operands: fp32 accumulation of 0/1 products is exact while every count stays below 2^24.
inline bool umma_fp4_possible(uint32_t n_samples) { return 2ull * n_samples < (1ull << 24); }
inline void umma_tile(bool fp4, uint32_t& TI, uint32_t& TJ) {
    TI = UMMA2_TILE;
    TJ = fp4 ? Umma3Cfg<true>::TILE_N : UMMA2_TILE;
}

static inline int umma_encode_map(CUtensorMap* map, uint8_t* base, uint32_t Kbytes, uint32_t Mpad, uint32_t box_rows, std::string& err) {
    PFN_tmapEncodeTiled enc = get_tmap_encoder();
    if (!enc) { err = "cuTensorMapEncodeTiled unavailable"; return -3; }
    cuuint64_t gdim[2] = {Kbytes, Mpad};
    cuuint64_t gstride[1] = {Kbytes};
    cuuint32_t box[2] = {UMMA_BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return -3; }
    return 0;
}

// Builds (once per matrix and encoding) the expanded operand and its tensor maps.
inline int umma_prepare(UmmaOperand& op, bool fp4, const uint64_t* d_rows, size_t stride64, uint32_t n_variants, uint32_t Mpad,
                        uint32_t n_samples, cudaStream_t stream, std::string& err, uint64_t* launches) {
    if (op.valid && op.fp4 == fp4 && op.mode == 0) return 0;
    op.valid = false;
    op.mode = 0;
    const uint32_t n_bits = 2 * n_samples;
    // one pipeline stage = 128 bytes of K = 128 int8 or 256 e2m1 elements
    const uint32_t k_per_stage = fp4 ? 2 * UMMA_BLOCK_K : UMMA_BLOCK_K;
    const uint32_t Kelems = (n_bits + k_per_stage - 1) / k_per_stage * k_per_stage;
    const uint32_t Kbytes = fp4 ? Kelems / 2 : Kelems;
    const size_t need = (size_t)Mpad * Kbytes;
    if (op.capacity < need) {
        if (op.d_bytes) cudaFree(op.d_bytes);
        op.d_bytes = nullptr;
        op.capacity = 0;
        cudaError_t e = cudaMalloc((void**)&op.d_bytes, need);
        if (e != cudaSuccess) { err = std::string("cudaMalloc(tensor-core operand): ") + cudaGetErrorString(e); return -3; }
        op.capacity = need;
    }
    op.Kbytes = Kbytes;
    op.Kelems = Kelems;
    op.Mpad = Mpad;
    op.fp4 = fp4;
    if (fp4) {
        dim3 grid(Mpad, (Kbytes / 32 + 127) / 128), block(128);
        expand_bits_to_e2m1_kernel<<<grid, block, 0, stream>>>(d_rows, stride64, n_variants, n_bits, op.d_bytes, Kbytes, Mpad);
    } else {
        dim3 grid(Mpad, (Kbytes / 64 + 127) / 128), block(128);
        expand_bits_to_bytes_kernel<<<grid, block, 0, stream>>>(d_rows, stride64, n_variants, n_bits, op.d_bytes, Kbytes, Mpad);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("expand kernel: ") + cudaGetErrorString(e); return -3; }
    if (launches) *launches += 1;
    int rc = umma_encode_map(&op.tmap, op.d_bytes, Kbytes, Mpad, UMMA_TILE_M, err);
    if (rc) return rc;
    rc = umma_encode_map(&op.tmap_b, op.d_bytes, Kbytes, Mpad, fp4 ? Umma3Cfg<true>::B_ROWS : UMMA_TILE_M, err);
    if (rc) return rc;
    op.valid = true;
    return 0;
}

// e2m1 operand copies of the planes kernels from the word-major bit planes plane[p][k][v]
// (pack.cuh): word k of plane p of variant v -> 32 nibbles at byte 16k of the variant's rows
// in the A order (32-row groups) and in the B order (plane-major tiles); see PlanesCfg.
template <int NP>
__global__ void expand_planes_to_e2m1_kernel(const uint32_t* __restrict__ planes, uint32_t K32, uint32_t Mpad, uint32_t n_variants,
                                             uint8_t* __restrict__ outA, uint8_t* __restrict__ outB, uint32_t Kbytes) {
    constexpr uint32_t G = 32u / NP, NV = 240u / NP;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = blockIdx.y;
    if (v >= n_variants || k >= K32) return;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t x = planes[((size_t)p * K32 + k) * Mpad + v];
        uint32_t b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t y = (x >> (8 * r)) & 0xFFu;  // bit t -> nibble t, 1 -> 0x2 (e2m1 1.0)
            y = (y | (y << 12)) & 0x000F000Fu;
            y = (y | (y << 6)) & 0x03030303u;
            y = (y | (y << 3)) & 0x11111111u;
            b[r] = y << 1;
        }
        const uint4 val = make_uint4(b[0], b[1], b[2], b[3]);
        const size_t rowA = (size_t)(v / G) * 32u + (v % G) * NP + p;
        const size_t rowB = (size_t)(v / NV) * 240u + (size_t)p * NV + (v % NV);
        *reinterpret_cast<uint4*>(outA + rowA * Kbytes + (size_t)k * 16u) = val;
        *reinterpret_cast<uint4*>(outB + rowB * Kbytes + (size_t)k * 16u) = val;
    }
}

inline void planes_tile(int mode, uint32_t& TI, uint32_t& TJ) {
    const uint32_t np = mode == MODE_UNPHASED_MISS ? 3u : 2u;
    TI = 8u * (32u / np);
    TJ = 240u / np;
}

// Builds (once per matrix and mode) the two e2m1 operand copies of a planes kernel.
inline int umma_prepare_planes(UmmaOperand& op, int mode, const uint32_t* d_planes, uint32_t K32, uint32_t Mpad, uint32_t n_variants,
                               cudaStream_t stream, std::string& err, uint64_t* launches) {
    if (op.valid && op.mode == mode) return 0;
    op.valid = false;
    uint32_t TI, TJ;
    planes_tile(mode, TI, TJ);
    const uint32_t Kbytes = K32 * 16u;  // K32 is a multiple of 16 words -> whole 128-byte K blocks
    const size_t rowsA = (size_t)((n_variants + TI - 1) / TI) * 256u, rowsB = (size_t)((n_variants + TJ - 1) / TJ) * 240u;
    const size_t needA = rowsA * Kbytes, needB = rowsB * Kbytes;
    auto grow = [&](uint8_t*& ptr, size_t& cap, size_t need) -> bool {
        if (cap >= need) return true;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        if (cudaMalloc((void**)&ptr, need) != cudaSuccess) return false;
        cap = need;
        return true;
    };
    if (!grow(op.d_bytes, op.capacity, needA) || !grow(op.d_bytes_b, op.capacity_b, needB)) {
        err = std::string("cudaMalloc(tensor-core plane operands): ") + cudaGetErrorString(cudaGetLastError());
        return -3;
    }
    cudaMemsetAsync(op.d_bytes, 0, needA, stream);
    cudaMemsetAsync(op.d_bytes_b, 0, needB, stream);
    dim3 grid((n_variants + 127) / 128, K32), block(128);
    if (mode == MODE_UNPHASED_MISS)
        expand_planes_to_e2m1_kernel<3><<<grid, block, 0, stream>>>(d_planes, K32, Mpad, n_variants, op.d_bytes, op.d_bytes_b, Kbytes);
    else
        expand_planes_to_e2m1_kernel<2><<<grid, block, 0, stream>>>(d_planes, K32, Mpad, n_variants, op.d_bytes, op.d_bytes_b, Kbytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("expand planes kernel: ") + cudaGetErrorString(e); return -3; }
    if (launches) *launches += 3;
    op.Kbytes = Kbytes;
    op.Kelems = 2 * Kbytes;
    op.Mpad = Mpad;
    op.fp4 = true;
    op.mode = mode;
    PFN_tmapEncodeTiled enc = get_tmap_encoder();
    if (!enc) { err = "cuTensorMapEncodeTiled unavailable"; return -3; }
    auto encode = [&](CUtensorMap* map, uint8_t* base, size_t rows, uint32_t box_rows) -> bool {
        cuuint64_t gdim[2] = {Kbytes, rows};
        cuuint64_t gstride[1] = {Kbytes};
        cuuint32_t box[2] = {UMMA_BLOCK_K, box_rows};
        cuuint32_t estr[2] = {1, 1};
        return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!encode(&op.tmap, op.d_bytes, rowsA, 128) || !encode(&op.tmap_b, op.d_bytes_b, rowsB, Umma3Cfg<true>::B_ROWS)) {
        err = "cuTensorMapEncodeTiled failed (plane operands)";
        return -3;
    }
    op.valid = true;
    return 0;
}

constexpr uint64_t UMMA3_L2_POLICY_A = L2_EVICT_NORMAL, UMMA3_L2_POLICY_B = L2_EVICT_NORMAL;
template <bool FP4, bool SCREEN, int MODE>
inline cudaError_t umma3_launch(UmmaOperand& op, const CountArgs& args, const DevParams& prm, uint32_t n_tiles, cudaStream_t stream) {
    // function attributes and the SM count belong to a device: one process may drive several
    // (twkb_calc -g 0,1,..., one host thread per device)
    static std::atomic<uint64_t> configured{0};
    static std::atomic<int> n_sm_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load() & bit)) {
        cudaError_t e = cudaFuncSetAttribute(count_umma3_kernel<FP4, SCREEN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Umma3Cfg<FP4>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        n_sm_of[dev & 63].store(n);
        configured.fetch_or(bit);
    }
    const int n_sm = n_sm_of[dev & 63].load();
    const uint32_t n_clusters = std::min<uint32_t>(n_tiles, (uint32_t)std::max(1, n_sm / 2));
    const uint32_t num_kblocks = op.Kbytes / UMMA_BLOCK_K;
    // Pacing (see the producer warp), whenever there is more than one CTA pair.
    //  * rows of >= 16 KB: chunks of 16 K blocks (2 KB per row), look-ahead 2 chunks: wave rows (<= ~8,000) x ~3 chunks in
    //    flight = <= ~50 MB of L2;
    //  * shorter rows: one chunk per tile, look-ahead 8 tiles. The tiles are dealt round-robin, so the tiles in flight are a
    //    compact block of the super-tile order only while the pairs advance at the same rate; free running they drift apart
    //    over the ~4,400 tiles each runs at C2 (a pair that meets more survivors, or sits farther from the L2 slices it
    //    reads, never catches up) until the tiles in flight span several super-tiles and stop sharing operand rows in L2.
    //    ncu, C2, per launch (profiles/round2_drift_probe_c2.log): DRAM 52-75 GB free running, 7.2 GB paced (compulsory
    //    ~6.5 GB), L2 hit rate 76-80 % -> 96 %; back to back 25.2 -> 26.9 ms free running (the board reaches its power
    //    cap), 24.47 ms paced. Look-ahead 1 / 2 / 4 / 8 / 16 / 32 tiles: 25.24 / 24.85 / 24.59 / 24.47 / 24.5-24.9 / 24.9-25.4 ms.
    uint32_t pace_kb = 16, pace_depth = 2, pace_min_kblocks = 1;
    if (num_kblocks < UMMA3_PACE_MIN_KBLOCKS) { pace_kb = num_kblocks; pace_depth = 8; }
    uint32_t* pace = nullptr;
#ifdef TWKB_PROFILING
    if (const char* e = getenv("TWKB_PACE_KB")) pace_kb = (uint32_t)std::max(1, atoi(e));
    if (const char* e = getenv("TWKB_PACE_DEPTH")) pace_depth = (uint32_t)std::max(0, atoi(e));
    if (const char* e = getenv("TWKB_PACE_MIN_KB")) pace_min_kblocks = (uint32_t)std::max(1, atoi(e));
#endif
    if (num_kblocks >= pace_min_kblocks && n_clusters > 1 && pace_depth > 0) {
        const uint32_t cpt = (num_kblocks + pace_kb - 1) / pace_kb;
        if (cpt > 1) pace_depth = std::min(pace_depth, cpt);
        const size_t need = (size_t)((n_tiles + n_clusters - 1) / n_clusters) * cpt;
        if (op.pace_capacity < need) {
            if (op.d_pace) cudaFree(op.d_pace);
            op.d_pace = nullptr;
            op.pace_capacity = 0;
            cudaError_t e = cudaMalloc((void**)&op.d_pace, need * sizeof(uint32_t));
            if (e != cudaSuccess) return e;
            op.pace_capacity = need;
        }
        cudaError_t e = cudaMemsetAsync(op.d_pace, 0, need * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        pace = op.d_pace;
    }
    // L2 eviction priority of the operand loads (UMMA3_L2_POLICY_*: see the C2 sweep in profiles/round2_l2_sweep_c2.log)
    uint64_t pol_a = UMMA3_L2_POLICY_A, pol_b = UMMA3_L2_POLICY_B;
#ifdef TWKB_PROFILING
    auto pol_of = [](const char* name, uint64_t dflt) {
        const char* e = getenv(name);
        if (!e) return dflt;
        return e[0] == '1' ? L2_EVICT_FIRST : e[0] == '2' ? L2_EVICT_LAST : L2_EVICT_NORMAL;
    };
    pol_a = pol_of("TWKB_L2_HINT_A", pol_a);
    pol_b = pol_of("TWKB_L2_HINT_B", pol_b);
#endif
    count_umma3_kernel<FP4, SCREEN, MODE><<<2 * n_clusters, UMMA3_THREADS, Umma3Cfg<FP4>::SMEM_BYTES, stream>>>(
        op.tmap, op.tmap_b, args, prm, num_kblocks, n_tiles, pace, pace_kb, pace_depth, pol_a, pol_b);
    return cudaGetLastError();
}

inline cudaError_t umma_launch(UmmaOperand& op, const CountArgs& args, const DevParams& prm, uint32_t n_tiles, cudaStream_t stream) {
    // the instantiation with the fast screen + survivor queue, or (minR2 = 0 / screen off:
    // every pair is a candidate) the one that compacts every chunk directly
    const bool screen = !args.screen_off && prm.minR2 > 0.0;
    if (op.mode == MODE_PHASED_MISS) return umma3_launch<true, false, MODE_PHASED_MISS>(op, args, prm, n_tiles, stream);
    if (op.mode == MODE_UNPHASED_NOMISS) return umma3_launch<true, false, MODE_UNPHASED_NOMISS>(op, args, prm, n_tiles, stream);
    if (op.mode == MODE_UNPHASED_MISS) return umma3_launch<true, false, MODE_UNPHASED_MISS>(op, args, prm, n_tiles, stream);
    if (op.fp4)
        return screen ? umma3_launch<true, true, 0>(op, args, prm, n_tiles, stream) : umma3_launch<true, false, 0>(op, args, prm, n_tiles, stream);
    return screen ? umma3_launch<false, true, 0>(op, args, prm, n_tiles, stream) : umma3_launch<false, false, 0>(op, args, prm, n_tiles, stream);
}

inline void umma_release(UmmaOperand& op) {
    if (op.d_bytes) cudaFree(op.d_bytes);
    if (op.d_bytes_b) cudaFree(op.d_bytes_b);
    if (op.d_pace) cudaFree(op.d_pace);
    op.d_bytes = op.d_bytes_b = nullptr;
    op.d_pace = nullptr;
    op.capacity = op.capacity_b = op.pace_capacity = 0;
    op.valid = false;
}

}  // namespace twkb
