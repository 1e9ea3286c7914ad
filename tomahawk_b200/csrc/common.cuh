// common.cuh -- shared device/host definitions of the B200 LD engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace twkb {

// ---------------------------------------------------------------------------
// Device-side per-variant metadata (16 B, one LDG.128): the fields of twk1_t
// (reference include/core.h:291-295) that the pair loop, the flags and the
// record need.
struct __align__(16) DevVariant {
    uint32_t pos;
    uint32_t ac;
    uint32_t rid;
    uint32_t flags;  // VF_*
};
enum : uint32_t {
    VF_HAS_MISSING = 1u << 0,  // an != 0
    VF_BAD_HWE = 1u << 1,      // hwe < 1e-4
    VF_GT_MISSING = 1u << 2,   // gt_missing flag of the record
};

// A pair that survived the in-kernel pre-screen, with its exact contingency
// counts. 48 B = 3 x STG.128.
//   mode 0 (phased math):   c[0..3] = REFREF, slot1 (A alt/B ref), slot4 (A ref/B alt), ALTALT
//   mode 1 (unphased math): c[0..8] = 3x3 genotype table t[gA][gB]
struct __align__(16) Candidate {
    uint32_t i, j;
    uint32_t c[9];
    uint32_t mode;
};
static_assert(sizeof(Candidate) == 48, "Candidate must be 48 bytes");

// Thresholds and mode flags, mirrored from twkb_settings for the kernels.
struct DevParams {
    double minP, minR2, maxR2, minDprime, maxDprime;
    double screenR2;       // minR2 * (1 - 1e-12): conservative in-kernel R2 pre-screen
    uint32_t n_samples;
    uint32_t n_variants;
    uint32_t window;       // window mode: 0 off, 1 -p/-u rule, 2 auto mode (row prune only), 3 -p -m -M rule
    uint32_t l_window;
    uint32_t emulate_quirks;
    uint32_t thresh_miss_phased;  // (uint32)(0.0047*n_s + 5.2913), ld_engine.cpp:1910
    uint32_t unphased;            // 1: unphased math for every pair
    uint32_t diag;                // 1: row range == col range, only i<j
    uint32_t lgamma_len;
    uint32_t bitmap_mode;         // -p -m -M with emulate_quirks: every masked pair takes the run-length slots (Q3)
    uint32_t single;              // scalc: no ac_i + ac_j <= 2 skip (ld_engine.cpp:2265, 2290: commented out there)
    uint32_t pair_filter;         // auto mode passes: 0 all pairs, 1 only pairs without a variant
                                  // with missing alleles, 2 only pairs with one (ld_engine.cpp:2775)
    uint32_t shard_blocks;        // position-sharded window runs: > 0 = only pairs whose earlier member lies in the first
                                  // shard_blocks .twk blocks of the loaded variants (the later blocks are the halo)
};

// Window-mode block structure (reference .twk blocks):
// per-variant block id, per-block first/last variant and row-prune limit.
struct DevBlocks {
    const uint32_t* blk_of;      // [n_variants]
    const uint32_t* blk_first;   // [n_blocks]
    const uint32_t* blk_last;    // [n_blocks] (inclusive)
    const uint32_t* blk_prune;   // [n_blocks] first block column that the row prune removes
};

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copies (sm_90+/sm_100a).
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP). dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace twkb
