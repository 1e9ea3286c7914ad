// engine.cu -- host side of libtwkb.so: context, device-resident matrix, tile
// scheduler (the B200 replacement of twk_ld_balancer / twk_ld_dynamic_balancer,
// reference lib/ld/ld_balancing.h:13-242), batch loop and the C-ABI of
// include/twkb.h. No CPU compute path exists in this file: every count and
// every statistic is produced by the CUDA kernels included below.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/twkb.h"
#include "comm.cuh"
#include "common.cuh"
#include "count_popc.cuh"
#include "count_sparse.cuh"
#include "count_umma.cuh"
#include "decode.cuh"
#include "hostio.h"
#include "pack.cuh"
#include "sort.cuh"
#include "stats.cuh"

namespace twkb {

static std::string g_create_error;
static std::mutex g_create_mutex;

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ctx->fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                        \
            return TWKB_ECUDA;                                                                \
        }                                                                                     \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// Tuning knobs of the development scripts (super-tile edge, batch size, kernel choice overrides ...) are read from the
// environment by the profiling build only; the product library has no switch that changes what runs.
#ifdef TWKB_PROFILING
static inline const char* tuning_env(const char* name) { return getenv(name); }
#else
static inline const char* tuning_env(const char*) { return nullptr; }
#endif

// The host keeps 8 of the 32 bytes of a caller's twkb_variant: contig and position are all the scheduler consults (window
// rules, .twk blocks, -c chunks). Every load rewrites this copy for every variant, on every rank: at 566,000 variants x 8
// ranks the full 32-byte copy was a third of the host-memory traffic of a load that is bound by exactly that.
struct HostVar { uint32_t rid, pos; };

struct Problem {  // one rectangular sub-problem of the pair grid
    uint32_t row_begin, row_end, col_begin, col_end;
    bool diag;
};

struct Context {
    twkb_settings st{};
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_begin = nullptr, ev_end = nullptr;

    // matrix
    bool loaded = false;
    uint32_t n_samples = 0, n_variants = 0, Mpad = 0;
    bool any_missing = false;
    int mode = -1;   // CountMode of the resident planes
    int np = 0;
    uint32_t K32 = 0;
    DevBuf<uint64_t> d_raw_data, d_raw_mask;  // row-major upload (kept: re-pack on mode change)
    size_t raw_stride = 0;
    DevBuf<uint32_t> d_planes, d_plane_popc;
    DevBuf<DevVariant> d_meta;
    DevBuf<double> d_lgamma;
    uint32_t lgamma_len = 0;
    std::vector<HostVar> h_meta;             // what the scheduler reads of a variant, resident order
    UmmaOperand umma;  // int8-expanded operand of the tensor-core kernel

    // Rare-variant (list) class. When active the resident matrix is ordered
    // [dense variants | sparse variants], both in file order; every index the kernels see is a
    // resident index, h_orig maps it back to the file order that orients a pair (A = lower).
    bool permuted = false;
    uint32_t nD = 0, nS = 0, sparse_T = 0;
    std::vector<uint32_t> h_orig;            // resident -> original index (identity when !permuted)
    bool h_orig_identity = false;            // h_orig currently holds 0, 1, 2, ...
    std::vector<HostVar> h_meta_orig;   // metadata in file order -- filled only when the rows were re-ordered (meta_orig())
    DevVariant* h_dm = nullptr;              // pinned staging of the device metadata
    size_t h_dm_cap = 0;
    uint32_t lgamma_ready = 0;               // length of the log-factorial table resident in d_lgamma
    DevBuf<uint32_t> d_orig, d_sp_off;
    DevBuf<uint2> d_sp_ent;
    uint64_t sp_entries = 0;
    std::vector<uint2> sp_plan_tiles;
    uint64_t sp_plan_pairs = 0;
    std::string sp_plan_key;
    DevBuf<uint2> d_sp_tiles;
    double est_cand_per_sp_tile = -1.0;
    std::vector<uint32_t> h_sp_off;               // [nS + 1] CSR offsets
    std::vector<uint64_t> sp_tile_words_prefix;   // prefix sum of word operations per sparse tile

    // window-mode block structure
    DevBuf<uint32_t> d_blk_of, d_blk_first, d_blk_last, d_blk_prune;
    std::vector<uint32_t> h_blk_first, h_blk_last, h_blk_prune, h_blk_of_orig;  // file order
    std::vector<uint32_t> file_blocks;  // twkb_set_blocks: first variant of every .twk block (empty: assume twk_block_size)

    // tile plan of the last run, reused while the sub-problem and the partition are unchanged
    std::vector<uint2> plan_tiles;
    uint64_t plan_pairs = 0;
    std::string plan_key;
    uint64_t matrix_epoch = 0;
    double est_cand_per_tile = -1.0;  // survivors per tile seen by the previous run (batch sizing)

    // work buffers
    DevBuf<uint2> d_tiles;
    DevBuf<Candidate> d_cands;
    DevBuf<unsigned long long> d_counters;  // [0] cand count, [1], [2] record counts of the two record buffers
    DevBuf<uint8_t> d_records[2];           // double-buffered: the statistics kernel fills one while the other drains
    int rec_cur = 0;                        // buffer the statistics kernel appends to
    bool stats_timing_pending = false;      // ev2/ev3 of the last statistics launch not read yet
    struct Flusher* flusher = nullptr;      // record drain thread (D2H + sink), created on first use
    size_t cand_cap = 0, rec_cap = 0;
    uint8_t* h_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    unsigned long long* h_counters = nullptr;  // pinned

    twkb_stats stats{};
    double ms_decode = 0.0;  // decode_runs_kernel time of the last twkb_load_runs

    // downstream consumer fed from the device-resident records (twkb_compute_decay): active while decay_bins != 0
    uint32_t decay_bins = 0, decay_width = 0;
    DevBuf<double> d_decay_sum;
    DevBuf<unsigned long long> d_decay_cnt;
    // twkb_compute_sorted: the forward records of the run are collected in d_collect instead of leaving the device
    bool collect = false;
    DevBuf<uint8_t> d_collect;
    uint64_t n_collected = 0;
    // twkb_compute_aggregate: 0 off, 1 = pass 1 (contig position ranges), 2 = pass 2 (raster)
    int agg_pass = 0;
    AggLayout agg_layout{};
    DevBuf<uint32_t> d_agg_min, d_agg_max;
    DevBuf<unsigned long long> d_agg_base;  // [n_contigs] coordinate bases + 1 counter of records with an unknown contig
    DevBuf<AggBin> d_agg_bins;

    // multi-GPU data plane (comm.cuh): set by twkb_comm_init, used by the sliced loads only
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    std::vector<cudaEvent_t> chunk_events;

    int fail(const std::string& m) {
        err = m;
        return TWKB_ECUDA;
    }
};

// Metadata in FILE order (block structure, visited pairs): h_meta itself unless the rare-variant class re-ordered the rows.
static const std::vector<HostVar>& meta_orig(const Context* ctx) { return ctx->permuted ? ctx->h_meta_orig : ctx->h_meta; }

static void settings_defaults(twkb_settings* s) {
    // reference lib/core.cpp:297-306
    std::memset(s, 0, sizeof(*s));
    s->square = 1;
    s->emulate_quirks = 1;
    s->c_level = 1;
    s->bl_size = 500;
    s->b_size = 10000;
    s->l_window = 1000000;
    s->n_threads = 1;
    s->l_surrounding = 500000;
    s->n_chunks = 1;
    s->c_chunk = 0;
    s->minP = 1;
    s->minR2 = 0.1;
    s->maxR2 = 100;
    s->minDprime = 0;
    s->maxDprime = 100;
    s->device = 0;
    s->part_index = 0;
    s->part_count = 1;
    s->kernel = TWKB_KERNEL_AUTO;
    s->twk_block_size = 500;
}

static int validate_settings(const twkb_settings* s, std::string& why) {
    if (s->single) {  // twk_ld::ComputeSingle, lib/ld/ld.cpp:679-687
        if (s->n_chunks != 1) { why = "Cannot use chunking in single mode!"; return TWKB_EINVAL; }
        if (s->window) { why = "Cannot use window in single mode!"; return TWKB_EINVAL; }
        if (s->single_targets < 0) { why = "illegal single_targets"; return TWKB_EINVAL; }
        if (s->part_count > 1) { why = "single mode runs on one device"; return TWKB_EINVAL; }
    }
    if (s->force_phased && s->forced_unphased) { why = "cannot force both phased and unphased"; return TWKB_EINVAL; }
    if (s->window && s->n_chunks != 1) { why = "Cannot use chunking in window mode!"; return TWKB_EINVAL; }  // ld.cpp:485
    if (s->n_chunks < 1 || s->c_chunk < 0 || s->c_chunk >= s->n_chunks) { why = "illegal chunk selection"; return TWKB_EINVAL; }
    if (s->part_count < 0 || (s->part_count > 0 && (s->part_index < 0 || s->part_index >= s->part_count))) {
        why = "illegal part_index/part_count";
        return TWKB_EINVAL;
    }
    if (s->minR2 < 0 || s->minR2 > 1) { why = "minR2 out of range"; return TWKB_EINVAL; }  // calc.h range checks
    if (s->shard_blocks < 0 || (s->shard_blocks > 0 && !s->window)) { why = "shard_blocks needs window mode"; return TWKB_EINVAL; }
    return TWKB_OK;
}

// Which per-pair window rule the reference applies for these flags (count_popc.cuh:
// window_pair_allowed). Without emulate_quirks every mode uses the -p / -u rule.
static uint32_t window_kind(const twkb_settings& s) {
    if (!s.window) return 0u;
    if (!s.emulate_quirks) return 1u;
    if (s.force_phased && s.low_memory && s.bitmaps) return 3u;  // ld_engine.cpp:1832-1834
    if (s.force_phased || s.forced_unphased) return 1u;
    return 2u;  // auto mode: twk_ld_slave::Calculate, ld_engine.cpp:1842-1843
}

static DevParams make_params(const Context* ctx, const Problem& pb) {
    DevParams p{};
    const twkb_settings& s = ctx->st;
    p.minP = s.minP; p.minR2 = s.minR2; p.maxR2 = s.maxR2; p.minDprime = s.minDprime; p.maxDprime = s.maxDprime;
    p.screenR2 = s.minR2 * (1.0 - 1e-12);
    p.n_samples = ctx->n_samples;
    p.n_variants = ctx->n_variants;
    p.window = window_kind(s);
    p.bitmap_mode = (s.emulate_quirks && s.force_phased && s.low_memory && s.bitmaps && !s.single) ? 1u : 0u;
    p.single = s.single ? 1u : 0u;
    p.l_window = (uint32_t)s.l_window;
    p.emulate_quirks = s.emulate_quirks ? 1u : 0u;
    p.thresh_miss_phased = (uint32_t)(0.0047 * ctx->n_samples + 5.2913);
    p.unphased = (ctx->mode >= 2) ? 1u : 0u;
    p.diag = pb.diag ? 1u : 0u;
    p.lgamma_len = ctx->lgamma_len;
    p.shard_blocks = s.window ? (uint32_t)s.shard_blocks : 0u;
    return p;
}

// Which planes a run needs: -u => unphased; -p => phased; neither ("auto",
// ld_engine.cpp:2737-2838) => unphased iff any variant has missing alleles is decided
// per pair by the reference; here auto mode is served as two passes by compute().
static int wanted_mode(const Context* ctx, bool unphased) {
    if (unphased) return ctx->any_missing ? MODE_UNPHASED_MISS : MODE_UNPHASED_NOMISS;
    return ctx->any_missing ? MODE_PHASED_MISS : MODE_PHASED_NOMISS;
}

static int ensure_planes(Context* ctx, int mode) {
    if (ctx->mode == mode) return TWKB_OK;
    const bool unphased = mode >= 2;
    const int np = (mode == 0) ? 1 : (mode == 3 ? 3 : 2);
    const uint32_t bits = unphased ? ctx->n_samples : 2 * ctx->n_samples;
    uint32_t K32 = (bits + 31) / 32;
    K32 = (K32 + POPC_TK - 1) / POPC_TK * POPC_TK;
    CUDA_TRY(ctx->d_planes.alloc((size_t)np * K32 * ctx->Mpad));
    CUDA_TRY(ctx->d_plane_popc.alloc((size_t)3 * ctx->Mpad));
    dim3 grid(ctx->Mpad / 32, (K32 + 31) / 32), block(32, 8);
    pack_planes_kernel<<<grid, block, 0, ctx->stream>>>(ctx->d_raw_data.p, ctx->any_missing ? ctx->d_raw_mask.p : nullptr,
                                                        ctx->raw_stride, ctx->n_variants, ctx->n_samples, mode,
                                                        ctx->d_planes.p, K32, ctx->Mpad);
    CUDA_TRY(cudaGetLastError());
    dim3 g2((ctx->Mpad + 127) / 128, np);
    plane_popc_kernel<<<g2, 128, 0, ctx->stream>>>(ctx->d_planes.p, np, K32, ctx->Mpad, ctx->d_plane_popc.p);
    CUDA_TRY(cudaGetLastError());
    ctx->stats.other_launches += 2;
    ctx->mode = mode;
    ctx->np = np;
    ctx->K32 = K32;
    ctx->umma.valid = false;
    return TWKB_OK;
}

// .twk block structure for the window rule (blocks of <= bs variants, one contig
// per block: lib/importer.cpp:196-236) and the row-prune limit of
// ld_balancing.h:189-196.
// First variant of every .twk block in file order (+ M at the end): from the file's index when the caller supplied it
// (twkb_set_blocks), else blocks of twk_block_size variants that never span two contigs (lib/importer.cpp:196-236).
static std::vector<uint32_t> block_starts(const Context* ctx, const std::vector<HostVar>& mo) {
    const uint32_t M = ctx->n_variants;
    std::vector<uint32_t> first;
    if (!ctx->file_blocks.empty() && ctx->file_blocks.back() < M) {
        first = ctx->file_blocks;
        first.push_back(M);
        return first;
    }
    const uint32_t bs = ctx->st.twk_block_size > 0 ? (uint32_t)ctx->st.twk_block_size : 500u;
    for (uint32_t v = 0; v < M;) {
        uint32_t e = v + 1;
        while (e < M && e - v < bs && mo[e].rid == mo[v].rid) ++e;
        first.push_back(v);
        v = e;
    }
    first.push_back(M);
    return first;
}

static void build_blocks_host(Context* ctx, std::vector<uint32_t>& blk_of_orig) {
    const uint32_t M = ctx->n_variants;
    const std::vector<HostVar>& mo = meta_orig(ctx);  // blocks are defined on the file order
    blk_of_orig.assign(M, 0);
    ctx->h_blk_first.clear();
    ctx->h_blk_last.clear();
    const std::vector<uint32_t> first = block_starts(ctx, mo);
    for (size_t b = 0; b + 1 < first.size(); ++b) {
        for (uint32_t x = first[b]; x < first[b + 1]; ++x) blk_of_orig[x] = (uint32_t)b;
        ctx->h_blk_first.push_back(first[b]);
        ctx->h_blk_last.push_back(first[b + 1] - 1);
    }
    const uint32_t nb = (uint32_t)ctx->h_blk_first.size();
    const uint32_t w = (uint32_t)ctx->st.l_window;
    ctx->h_blk_prune.assign(nb, nb);
    for (uint32_t bi = 0; bi < nb; ++bi) {
        const uint32_t last_pos = mo[ctx->h_blk_last[bi]].pos;
        for (uint32_t bj = bi + 1; bj < nb; ++bj) {
            if ((uint32_t)(mo[ctx->h_blk_first[bj]].pos - last_pos) > w) {
                ctx->h_blk_prune[bi] = bj;
                break;
            }
        }
    }
    ctx->h_blk_of_orig = blk_of_orig;
}

static int build_blocks(Context* ctx) {
    const uint32_t M = ctx->n_variants;
    std::vector<uint32_t> blk_of_orig;
    build_blocks_host(ctx, blk_of_orig);
    const uint32_t nb = (uint32_t)ctx->h_blk_first.size();
    // device copies in resident indices (identical to the file order unless the matrix is
    // ordered [dense | sparse])
    std::vector<uint32_t> inv(M), blk_of(ctx->Mpad, 0), d_first(nb), d_last(nb);
    for (uint32_t x = 0; x < M; ++x) inv[ctx->h_orig[x]] = x;
    for (uint32_t x = 0; x < M; ++x) blk_of[x] = blk_of_orig[ctx->h_orig[x]];
    for (uint32_t b = 0; b < nb; ++b) { d_first[b] = inv[ctx->h_blk_first[b]]; d_last[b] = inv[ctx->h_blk_last[b]]; }
    CUDA_TRY(ctx->d_blk_of.alloc(ctx->Mpad));
    CUDA_TRY(ctx->d_blk_first.alloc(nb));
    CUDA_TRY(ctx->d_blk_last.alloc(nb));
    CUDA_TRY(ctx->d_blk_prune.alloc(nb));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_blk_of.p, blk_of.data(), ctx->Mpad * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_blk_first.p, d_first.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_blk_last.p, d_last.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_blk_prune.p, ctx->h_blk_prune.data(), nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return TWKB_OK;
}

// Pairs the reference counts as visited (progress n_var; ld_engine.cpp:1933,2015,2607).
static uint64_t visited_pairs(const Context* ctx, const Problem& pb) {
    if (!ctx->st.window) {
        const uint64_t nr = pb.row_end - pb.row_begin, nc = pb.col_end - pb.col_begin;
        return pb.diag ? (nr * nr - nr) / 2 : nr * nc;
    }
    uint64_t total = 0;
    const uint32_t nb = (uint32_t)ctx->h_blk_first.size();
    const uint32_t w = (uint32_t)ctx->st.l_window;
    const uint32_t kind = window_kind(ctx->st);
    const uint32_t own = ctx->st.shard_blocks > 0 ? std::min(nb, (uint32_t)ctx->st.shard_blocks) : nb;  // a position shard counts its own block rows
    for (uint32_t bi = 0; bi < own; ++bi) {
        const uint64_t ni = ctx->h_blk_last[bi] - ctx->h_blk_first[bi] + 1;
        const HostVar& vf = meta_orig(ctx)[ctx->h_blk_first[bi]];
        for (uint32_t bj = bi; bj < ctx->h_blk_prune[bi] || bj == bi; ++bj) {
            if (bj >= nb) break;
            const HostVar& vl = meta_orig(ctx)[ctx->h_blk_last[bj]];
            const bool aborted = kind == 1u && vf.rid == vl.rid && (uint32_t)(vl.pos - vf.pos) > w;
            if (!aborted) {
                const uint64_t nj = ctx->h_blk_last[bj] - ctx->h_blk_first[bj] + 1;
                total += (bi == bj) ? (ni * ni - ni) / 2 : ni * nj;
            }
        }
    }
    return total;
}

// The sub-problem of this run: -c/-C chunk selection of twk_ld_balancer::Build
// (ld_balancing.h:23-80) over .twk blocks, expressed in variant ranges.
static int select_problem(Context* ctx, Problem& pb) {
    const uint32_t M = ctx->n_variants;
    pb = {0, M, 0, M, true};
    const int parts = ctx->st.n_chunks;
    if (parts <= 1) return TWKB_OK;
    const std::vector<uint32_t> first = block_starts(ctx, meta_orig(ctx));
    const uint32_t nb = (uint32_t)first.size() - 1;
    if ((uint32_t)parts > nb) { ctx->err = "more sub-problems than blocks"; return TWKB_EINVAL; }
    uint32_t factor = 0;
    for (uint32_t i = 1; i < (uint32_t)parts; ++i)
        if (((i * i - i) / 2) + i == (uint32_t)parts) { factor = i; break; }
    if (factor == 0) { ctx->err = "number of sub-problems is not k(k+1)/2"; return TWKB_EINVAL; }
    const uint32_t chunk = nb / factor;
    uint32_t k = 0;
    for (uint32_t i = 0; i < factor; ++i)
        for (uint32_t j = i; j < factor; ++j, ++k) {
            if (k != (uint32_t)ctx->st.c_chunk) continue;
            // The reference takes from = to - chunk_size (ld_balancing.h:64-67), which silently
            // drops blocks [chunk*(factor-1), nb - chunk) of the last chunk whenever nb is not a
            // multiple of factor; here the last chunk starts where the previous one ended, so
            // the chunks always tile the whole triangle (explained disagreements).
            const uint32_t tR = (j + 1 == factor ? nb : chunk * (j + 1)), fR = chunk * j;
            const uint32_t tL = (i + 1 == factor ? nb : chunk * (i + 1)), fL = chunk * i;
            pb.row_begin = first[fL]; pb.row_end = first[tL];
            pb.col_begin = first[fR]; pb.col_end = first[tR];
            pb.diag = (i == j);
            return TWKB_OK;
        }
    ctx->err = "chunk not found";
    return TWKB_EINVAL;
}

// Tile list of one sub-problem for a TI x TJ kernel, this context's share only.
// Tiles are emitted in super-tile order (SUPER x SUPER tiles) so that CTAs that are
// resident together share operand rows in L2; parts take interleaved tiles.
static void build_tiles(const Context* ctx, const Problem& pb, uint32_t TI, uint32_t TJ, uint32_t SUPER,
                        std::vector<uint2>& tiles, uint64_t* pairs_out = nullptr) {
    tiles.clear();
    uint64_t pairs = 0;
    // pairs (i in rows, j in cols, i<j when diag) that fall inside one tile
    auto tile_pairs = [&](uint32_t i0, uint32_t j0) -> uint64_t {
        const uint64_t ia = std::max(i0, pb.row_begin), ib = std::min<uint64_t>(i0 + TI, pb.row_end);
        const uint64_t ja = std::max(j0, pb.col_begin), jb = std::min<uint64_t>(j0 + TJ, pb.col_end);
        if (ia >= ib || ja >= jb) return 0;
        if (!pb.diag || ib <= ja) return (ib - ia) * (jb - ja);  // every i is below every j
        uint64_t n = 0;
        for (uint64_t i = ia; i < ib; ++i) {
            const uint64_t lo = std::max<uint64_t>(ja, i + 1);
            if (lo < jb) n += jb - lo;
        }
        return n;
    };
    const uint32_t ti0 = pb.row_begin / TI, ti1 = (pb.row_end + TI - 1) / TI;
    const uint32_t tj0 = pb.col_begin / TJ, tj1 = (pb.col_end + TJ - 1) / TJ;
    // tile-level window skip: only the -p / -u rule drops every pair farther apart than the window
    // (the auto-mode and -M rules keep them inside un-pruned block pairs; the kernels decide per pair)
    const bool window = window_kind(ctx->st) == 1u;
    // auto mode / -M: a tile is dropped when the balancer's row prune (ld_balancing.h:189-196) removes every
    // block pair it touches (file order == resident order required: not with the rare-variant class)
    const bool prune_only = window_kind(ctx->st) >= 2u && !ctx->permuted && ctx->h_blk_of_orig.size() == ctx->n_variants;
    const uint32_t w = (uint32_t)ctx->st.l_window;
    const uint32_t M = ctx->n_variants;
    const int parts = std::max(1, ctx->st.part_count), part = ctx->st.part_index;
    const uint32_t shard_blocks = (ctx->st.window && ctx->st.shard_blocks > 0 && ctx->h_blk_of_orig.size() == M && ctx->h_orig.size() == M)
                                      ? (uint32_t)ctx->st.shard_blocks : 0u;
    auto blk_res = [&](uint32_t x) { return ctx->h_blk_of_orig[ctx->h_orig[x]]; };  // .twk block of a resident index
    // A part's share. Many super-tiles (>= 16 per part): whole super-tiles are dealt, each to the part
    // with the fewest tiles so far (every part computes the same assignment), so that a part's
    // operand working set is the rows of ITS super-tiles -- dealing single tiles round-robin makes
    // every part stream every operand row of every super-tile for 1/parts of the tiles (DRAM traffic
    // per tile x4.6 at 8 parts). Few super-tiles: single tiles round-robin in emission order, which
    // balances any grid size.
    uint64_t group = 0;
    enum { COUNT_ONLY, EMIT_ALL, EMIT_ROUND_ROBIN };
    // Order inside a super-tile: row-major. (Profiling build: row-major over INNER_I x INNER_J sub-blocks of tiles, so that
    // the tiles in flight together are a compact rectangle -- measured at C2 with paced CTA pairs: DRAM 7.3 -> 6.8 GB per
    // launch, time +0.4 %, not adopted; scripts/l2_sweep.py, profiles/round2_l2_sweep_c2.log.)
    uint32_t INNER_I = SUPER, INNER_J = SUPER;
#ifdef TWKB_PROFILING
    if (const char* e = getenv("TWKB_INNER_I")) { if (atoi(e) > 0) INNER_I = std::min<uint32_t>((uint32_t)atoi(e), SUPER); }
    if (const char* e = getenv("TWKB_INNER_J")) { if (atoi(e) > 0) INNER_J = std::min<uint32_t>((uint32_t)atoi(e), SUPER); }
#endif
    auto walk_super = [&](uint32_t si, uint32_t sj, int what) -> uint64_t {
        uint64_t n = 0;
        const uint32_t se_i = std::min(si + SUPER, ti1), se_j = std::min(sj + SUPER, tj1);
        for (uint32_t bi = si; bi < se_i; bi += INNER_I)
        for (uint32_t bj = sj; bj < se_j; bj += INNER_J)
        for (uint32_t ti = bi; ti < std::min(bi + INNER_I, se_i); ++ti)
            for (uint32_t tj = bj; tj < std::min(bj + INNER_J, se_j); ++tj) {
                const uint32_t i0 = ti * TI, j0 = tj * TJ;
                if (pb.diag && j0 + TJ - 1 <= i0) continue;  // no i<j in this tile
                if (window) {
                    // every pair of the tile is farther apart than the window on one contig
                    const uint32_t il = std::min(i0 + TI, std::min(M, pb.row_end)) - 1;
                    const uint32_t jf = std::max(j0, pb.col_begin), jl = std::min(j0 + TJ, std::min(M, pb.col_end)) - 1;
                    if (jf > il) {
                        const HostVar &a0 = ctx->h_meta[std::max(i0, pb.row_begin)], &a1 = ctx->h_meta[il];
                        const HostVar &b0 = ctx->h_meta[jf], &b1 = ctx->h_meta[jl];
                        if (a0.rid == a1.rid && a1.rid == b0.rid && b0.rid == b1.rid && b0.pos >= a1.pos &&
                            (uint32_t)(b0.pos - a1.pos) > w)
                            continue;
                    }
                }
                if (shard_blocks) {
                    // position shard: both members of every pair of the tile lie in halo blocks (file order is monotone
                    // within the range a tile plan covers, so the first row / column has the lowest block)
                    const uint32_t ia = std::max(i0, pb.row_begin), ja = std::max(j0, pb.col_begin);
                    if (ia < M && ja < M && blk_res(ia) >= shard_blocks && blk_res(ja) >= shard_blocks) continue;
                }
                if (prune_only) {
                    const uint32_t ia = std::max(i0, pb.row_begin), il = std::min(i0 + TI, std::min(M, pb.row_end)) - 1;
                    const uint32_t jf = std::max(j0, pb.col_begin);
                    if (ia <= il && jf > il && jf < M) {
                        const uint32_t bjf = ctx->h_blk_of_orig[jf];
                        bool all_pruned = true;
                        for (uint32_t b = ctx->h_blk_of_orig[ia]; b <= ctx->h_blk_of_orig[il] && all_pruned; ++b)
                            all_pruned = bjf >= ctx->h_blk_prune[b];
                        if (all_pruned) continue;
                    }
                }
                ++n;
                if (what == COUNT_ONLY) continue;
                if (what == EMIT_ALL || (group % parts) == (uint64_t)part) {
                    tiles.push_back(make_uint2(i0, j0));
                    pairs += tile_pairs(i0, j0);
                }
                ++group;
            }
        return n;
    };
    auto each_super = [&](auto&& fn) {
        for (uint32_t si = ti0; si < ti1; si += SUPER)
            for (uint32_t sj = tj0; sj < tj1; sj += SUPER) {
                if (pb.diag && (uint64_t)(sj + SUPER) * TJ <= (uint64_t)si * TI) continue;
                fn(si, sj);
            }
    };
    bool by_super = false;
    std::vector<int> owner;
    if (parts > 1) {
        std::vector<uint64_t> n_in;
        each_super([&](uint32_t si, uint32_t sj) { n_in.push_back(walk_super(si, sj, COUNT_ONLY)); });
        size_t non_empty = 0;
        for (uint64_t n : n_in) non_empty += n ? 1 : 0;
        if (non_empty >= (size_t)16 * parts) {
            by_super = true;
            owner.assign(n_in.size(), 0);
            std::vector<uint64_t> load(parts, 0);
            for (size_t k = 0; k < n_in.size(); ++k) {
                int best = 0;
                for (int q = 1; q < parts; ++q)
                    if (load[q] < load[best]) best = q;
                owner[k] = best;
                load[best] += n_in[k];
            }
        }
    }
    size_t k_super = 0;
    each_super([&](uint32_t si, uint32_t sj) {
        if (!by_super) walk_super(si, sj, EMIT_ROUND_ROBIN);
        else if (owner[k_super] == part) walk_super(si, sj, EMIT_ALL);
        ++k_super;
    });
    if (pairs_out) *pairs_out = pairs;
}

// Tiles of the sparse kernel (count_sparse.cuh): SP_ROWS sparse rows x SP_TJ columns, covering every
// pair with a sparse member exactly once -- sparse row r with every dense column (resident index
// < nD) and every LATER sparse column. Window mode drops tiles whose column parts are all out of
// reach of all rows (both classes keep the file order, so first/last bound the positions).
static void build_sparse_tiles(const Context* ctx, std::vector<uint2>& tiles, uint64_t* pairs_out) {
    tiles.clear();
    const uint32_t M = ctx->n_variants, nD = ctx->nD;
    const bool window = window_kind(ctx->st) == 1u;
    const uint32_t w = (uint32_t)ctx->st.l_window;
    const int parts = std::max(1, ctx->st.part_count), part = ctx->st.part_index;
    const uint32_t shard_blocks = (ctx->st.window && ctx->st.shard_blocks > 0 && ctx->h_blk_of_orig.size() == M && ctx->h_orig.size() == M)
                                      ? (uint32_t)ctx->st.shard_blocks : 0u;
    auto blk_res = [&](uint32_t x) { return ctx->h_blk_of_orig[ctx->h_orig[x]]; };
    uint64_t group = 0, pairs = 0;
    for (uint32_t r0 = nD; r0 < M; r0 += SP_ROWS) {
        const uint32_t r1 = std::min<uint32_t>(r0 + SP_ROWS, M);
        const HostVar &a0 = ctx->h_meta[r0], &a1 = ctx->h_meta[r1 - 1];
        for (uint32_t j0 = 0; j0 < M; j0 += SP_TJ) {
            const uint32_t j1 = std::min<uint32_t>(j0 + SP_TJ, M);
            const uint32_t da = j0, db = std::min(j1, nD);        // dense columns of the tile
            const uint32_t sa = std::max(j0, r0 + 1), sb = j1;   // sparse columns later than the first row
            const bool has_d = da < db, has_s = sa < sb;
            if (!has_d && !has_s) continue;
            if (shard_blocks && blk_res(r0) >= shard_blocks && (!has_d || blk_res(da) >= shard_blocks) &&
                (!has_s || blk_res(sa) >= shard_blocks))
                continue;  // position shard: halo rows x halo columns
            if (window) {
                auto far = [&](uint32_t ca, uint32_t cb) {
                    const HostVar &b0 = ctx->h_meta[ca], &b1 = ctx->h_meta[cb - 1];
                    if (!(a0.rid == a1.rid && b0.rid == b1.rid && a0.rid == b0.rid)) return false;
                    if (b0.pos >= a1.pos && (uint32_t)(b0.pos - a1.pos) > w) return true;
                    if (a0.pos >= b1.pos && (uint32_t)(a0.pos - b1.pos) > w) return true;
                    return false;
                };
                if ((!has_d || far(da, db)) && (!has_s || far(sa, sb))) continue;
            }
            if ((group % parts) == (uint64_t)part) {
                tiles.push_back(make_uint2(r0, j0));
                uint64_t n = has_d ? (uint64_t)(r1 - r0) * (db - da) : 0;
                for (uint32_t r = r0; r < r1; ++r) {
                    const uint32_t lo = std::max(j0, r + 1);
                    if (lo < j1) n += j1 - lo;
                }
                pairs += n;
            }
            ++group;
        }
    }
    if (pairs_out) *pairs_out = pairs;
}

template <int MODE>
static cudaError_t launch_popc(Context* ctx, const CountArgs& args, const DevParams& prm, uint32_t n_tiles) {
    // function attributes belong to a device: one process may drive several (twkb_calc -g 0,1,...)
    static std::atomic<uint64_t> configured{0};
    const size_t smem = popc_smem_bytes<MODE>();
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(configured.load() & bit)) {
        cudaError_t e = cudaFuncSetAttribute(count_popc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit);
    }
    count_popc_kernel<MODE><<<n_tiles, POPC_THREADS, smem, ctx->stream>>>(args, prm);
    return cudaGetLastError();
}

static void tile_dims(int mode, bool umma, bool fp4, uint32_t& TI, uint32_t& TJ) {
    if (umma) { umma_tile(fp4, TI, TJ); return; }
    switch (mode) {
        case MODE_PHASED_NOMISS: TI = popc_tile_i<0>(); TJ = popc_tile_j<0>(); break;
        case MODE_PHASED_MISS: TI = popc_tile_i<1>(); TJ = popc_tile_j<1>(); break;
        case MODE_UNPHASED_NOMISS: TI = popc_tile_i<2>(); TJ = popc_tile_j<2>(); break;
        default: TI = popc_tile_i<3>(); TJ = popc_tile_j<3>(); break;
    }
}

static int ensure_work_buffers(Context* ctx, uint64_t max_tile_pairs) {
    size_t want_cand = 16u << 20, want_rec = 16u << 20;
    if (const char* e = getenv("TWKB_CAND_CAP")) want_cand = (size_t)atoll(e);
    if (const char* e = getenv("TWKB_REC_CAP")) want_rec = (size_t)atoll(e);
    want_cand = std::max<size_t>(want_cand, max_tile_pairs);
    want_rec = std::max<size_t>(want_rec, want_cand);
    if (ctx->cand_cap < want_cand) {
        CUDA_TRY(ctx->d_cands.alloc(want_cand));
        ctx->cand_cap = want_cand;
    }
    if (ctx->rec_cap < want_rec) {
        CUDA_TRY(ctx->d_records[0].alloc(want_rec * TWKB_RECORD_BYTES));
        CUDA_TRY(ctx->d_records[1].alloc(want_rec * TWKB_RECORD_BYTES));
        ctx->rec_cap = want_rec;
    }
    CUDA_TRY(ctx->d_counters.alloc(4));
    if (!ctx->h_counters) CUDA_TRY(cudaMallocHost((void**)&ctx->h_counters, 4 * sizeof(unsigned long long)));
    if (!ctx->h_stage[0]) {
        ctx->stage_bytes = (size_t)(32u << 20) / TWKB_RECORD_BYTES * TWKB_RECORD_BYTES;
        CUDA_TRY(cudaMallocHost((void**)&ctx->h_stage[0], ctx->stage_bytes));
        CUDA_TRY(cudaMallocHost((void**)&ctx->h_stage[1], ctx->stage_bytes));
    }
    return TWKB_OK;
}

// ---- record drain ------------------------------------------------------------------------------
// north_star (4): "compacted results stream to the host writer by async D2H on side streams while
// the next tiles compute". A context owns one drain thread. The batch loop hands it a FULL record
// buffer (all statistics kernels that wrote it have completed) and goes on launching count and
// statistics kernels into the other buffer; the thread copies the records to the host in pinned
// chunks on copy_stream (two staging buffers: chunk c + 1 is in flight while the sink consumes
// chunk c) and calls the sink. With R2 >= 0 (configs[0]: 5.2 GB of records) the run is then bound
// by the PCIe link, not by PCIe + compute.
struct Flusher {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    struct Job { int buf; uint64_t n; };
    std::deque<Job> jobs;
    bool busy[2] = {false, false};
    bool stop = false;
    int rc = TWKB_OK;
    std::string err;
    twkb_sink_fn sink = nullptr;
    void* user = nullptr;
    uint64_t bytes_d2h = 0;
};

// D2H of one record buffer in pinned chunks, overlapped with the sink (drain thread).
static int flush_records(Context* ctx, int buf, uint64_t n_records, twkb_sink_fn sink, void* user, std::string& err, uint64_t& bytes) {
    if (n_records == 0) return TWKB_OK;
    auto fail = [&](cudaError_t e, const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return TWKB_ECUDA; };
    const size_t total = (size_t)n_records * TWKB_RECORD_BYTES;
    size_t off = 0;
    int cur = 0;
    size_t pending_bytes = 0;
    int pending = -1;
    while (off < total || pending >= 0) {
        size_t nbytes = 0;
        if (off < total) {
            nbytes = std::min(ctx->stage_bytes, total - off);
            cudaError_t e = cudaMemcpyAsync(ctx->h_stage[cur], ctx->d_records[buf].p + off, nbytes, cudaMemcpyDeviceToHost, ctx->copy_stream);
            if (e != cudaSuccess) return fail(e, "record D2H");
        }
        if (pending >= 0) {
            if (sink && sink(user, ctx->h_stage[pending], pending_bytes / TWKB_RECORD_BYTES) != 0) {
                cudaStreamSynchronize(ctx->copy_stream);
                err = "record sink aborted the run";
                return TWKB_ESINK;
            }
            pending = -1;
        }
        if (nbytes) {
            cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
            if (e != cudaSuccess) return fail(e, "record D2H");
            pending = cur;
            pending_bytes = nbytes;
            off += nbytes;
            cur ^= 1;
            bytes += nbytes;
        }
    }
    return TWKB_OK;
}

static void flusher_loop(Context* ctx) {
    Flusher* f = ctx->flusher;
    cudaSetDevice(ctx->device);
    std::unique_lock<std::mutex> lk(f->mu);
    for (;;) {
        f->cv.wait(lk, [&] { return f->stop || !f->jobs.empty(); });
        if (f->jobs.empty()) return;  // stop requested and nothing left
        const Flusher::Job job = f->jobs.front();
        f->jobs.pop_front();
        const bool skip = f->rc != TWKB_OK;  // after a failure the remaining buffers are only released
        twkb_sink_fn sink = f->sink;
        void* user = f->user;
        lk.unlock();
        int rc = TWKB_OK;
        std::string err;
        uint64_t bytes = 0;
        if (!skip) {
            try {
                rc = flush_records(ctx, job.buf, job.n, sink, user, err, bytes);
            } catch (...) {  // a throwing sink must not take the process down from a foreign thread
                rc = TWKB_ESINK;
                err = "record sink threw an exception";
            }
        }
        lk.lock();
        f->bytes_d2h += bytes;
        if (rc != TWKB_OK && f->rc == TWKB_OK) { f->rc = rc; f->err = err; }
        f->busy[job.buf] = false;
        f->cv.notify_all();
    }
}

static void flusher_begin(Context* ctx, twkb_sink_fn sink, void* user) {
    if (!ctx->flusher) {
        ctx->flusher = new Flusher();
        ctx->flusher->th = std::thread(flusher_loop, ctx);
    }
    std::lock_guard<std::mutex> g(ctx->flusher->mu);
    ctx->flusher->sink = sink;
    ctx->flusher->user = user;
    ctx->flusher->rc = TWKB_OK;
    ctx->flusher->err.clear();
    ctx->flusher->bytes_d2h = 0;
}
static void flusher_submit(Context* ctx, int buf, uint64_t n) {
    Flusher* f = ctx->flusher;
    std::lock_guard<std::mutex> g(f->mu);
    f->busy[buf] = true;
    f->jobs.push_back({buf, n});
    f->cv.notify_all();
}
// Blocks until the drain thread has released `buf` (buf < 0: every buffer). Returns the drain's status.
static int flusher_wait(Context* ctx, int buf) {
    Flusher* f = ctx->flusher;
    std::unique_lock<std::mutex> lk(f->mu);
    f->cv.wait(lk, [&] { return buf >= 0 ? !f->busy[buf] : (!f->busy[0] && !f->busy[1]); });
    if (f->rc != TWKB_OK) ctx->err = f->err;
    return f->rc;
}
static void flusher_destroy(Context* ctx) {
    if (!ctx->flusher) return;
    {
        std::lock_guard<std::mutex> g(ctx->flusher->mu);
        ctx->flusher->stop = true;
        ctx->flusher->cv.notify_all();
    }
    if (ctx->flusher->th.joinable()) ctx->flusher->th.join();
    delete ctx->flusher;
    ctx->flusher = nullptr;
}

// The current record buffer cannot take the next batch: hand it to the drain thread (or, for a
// resident run, just count it) and continue in the other one once that is free.
// Device-side consumers of a record buffer whose statistics kernels have completed (stream order).
static int consume_records(Context* ctx, int buf, uint64_t n) {
    if (ctx->collect) {
        if (n == 0) return TWKB_OK;
        const size_t need = (size_t)(ctx->n_collected + n) * TWKB_RECORD_BYTES;
        if (need > ctx->d_collect.n) {  // grow: at least double, keep what is there
            DevBuf<uint8_t> bigger;
            CUDA_TRY(bigger.alloc(std::max(need, 2 * ctx->d_collect.n)));
            if (ctx->n_collected)
                CUDA_TRY(cudaMemcpyAsync(bigger.p, ctx->d_collect.p, (size_t)ctx->n_collected * TWKB_RECORD_BYTES, cudaMemcpyDeviceToDevice, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            std::swap(bigger.p, ctx->d_collect.p);
            std::swap(bigger.n, ctx->d_collect.n);
            bigger.release();
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->d_collect.p + (size_t)ctx->n_collected * TWKB_RECORD_BYTES, ctx->d_records[buf].p, (size_t)n * TWKB_RECORD_BYTES,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->n_collected += n;
        return TWKB_OK;
    }
    if (ctx->agg_pass && n) {
        const unsigned g = (unsigned)std::min<uint64_t>((n + 255) / 256, 148ull * 8);
        if (ctx->agg_pass == 1)
            agg_range_kernel<<<g, 256, 0, ctx->stream>>>(ctx->d_records[buf].p, n, ctx->agg_layout.n_contigs, ctx->d_agg_min.p, ctx->d_agg_max.p,
                                                        ctx->d_agg_base.p + ctx->agg_layout.n_contigs);
        else
            agg_bin_kernel<<<g, 256, 0, ctx->stream>>>(ctx->d_records[buf].p, n, ctx->agg_layout, ctx->d_agg_bins.p);
        CUDA_TRY(cudaGetLastError());
        ctx->stats.other_launches += 1;
        return TWKB_OK;
    }
    if (!ctx->decay_bins || n == 0) return TWKB_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148ull * 8);
    decay_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_records[buf].p, n, ctx->decay_width, ctx->decay_bins, ctx->d_decay_sum.p,
                                               ctx->d_decay_cnt.p);
    CUDA_TRY(cudaGetLastError());
    ctx->stats.other_launches += 1;
    return TWKB_OK;
}

static int rotate_records(Context* ctx, uint64_t fill, bool resident) {
    ctx->stats.records_out += fill;
    if (resident) {
        const int rc = consume_records(ctx, ctx->rec_cur, fill);
        if (rc) return rc;
    }
    if (!resident && fill) flusher_submit(ctx, ctx->rec_cur, fill);
    ctx->rec_cur ^= 1;
    if (!resident) {
        const int rc = flusher_wait(ctx, ctx->rec_cur);
        if (rc) return rc;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_counters.p + 1 + ctx->rec_cur, 0, sizeof(unsigned long long), ctx->stream));
    return TWKB_OK;
}

// Batch loop shared by the dense and the sparse phase of a pass: launches `launch(t, nb)` over
// the tile list in batches the candidate buffer can hold (retry with fewer tiles on overflow) and
// runs the statistics kernel on the survivors of every batch. ONE host round trip per batch (the
// candidate count sizes the statistics grid); the statistics kernel of batch b, the record drain
// and the count kernel of batch b + 1 are never waited for individually.
struct BatchPlan {
    size_t n_tiles = 0;
    uint64_t tile_pairs = 0;      // pairs per tile (worst-case candidates)
    bool no_screen = false;
    bool sparse = false;
    double* est_cand_per_tile = nullptr;
};
static int read_stats_timing(Context* ctx) {
    if (!ctx->stats_timing_pending) return TWKB_OK;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3));
    ctx->stats.ms_stats_kernel += ms;
    ctx->stats_timing_pending = false;
    return TWKB_OK;
}
template <typename LaunchFn, typename AccountFn>
static int run_batches(Context* ctx, const BatchPlan& bp, const DevParams& prm, LaunchFn launch, AccountFn account, bool resident,
                       std::vector<Candidate>* dump) {
    // Batch size: the candidate buffer must hold a whole batch. Start from the
    // worst case when nothing can be screened out, else optimistic and adapt.
    uint64_t batch = bp.no_screen ? std::max<uint64_t>(1, ctx->cand_cap / bp.tile_pairs)
                                  : std::max<uint64_t>(1, ctx->cand_cap / bp.tile_pairs * 64);
    if (!bp.no_screen && *bp.est_cand_per_tile >= 0.0)  // the previous run over this matrix measured the survivor rate
        batch = std::max<uint64_t>(batch, (uint64_t)(0.25 * ctx->cand_cap / std::max(1.0, *bp.est_cand_per_tile)));
    if (const char* e = tuning_env("TWKB_BATCH_TILES")) batch = std::max<uint64_t>(1, (uint64_t)atoll(e));
    size_t t = 0;
    while (t < bp.n_tiles) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(batch, bp.n_tiles - t);
        CUDA_TRY(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        CUDA_TRY(launch(t, nb));
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // also completes the statistics kernel of the previous batch
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (bp.sparse) { ctx->stats.ms_sparse_kernel += ms; ctx->stats.sparse_launches += 1; }
        else { ctx->stats.ms_count_kernel += ms; ctx->stats.count_launches += 1; }
        {
            const int rc = read_stats_timing(ctx);
            if (rc) return rc;
        }
        const uint64_t ncand = ctx->h_counters[0];
        if (ncand > ctx->cand_cap) {  // overflow: redo this batch with fewer tiles
            if (nb == 1) { ctx->err = "candidate buffer smaller than one tile"; return TWKB_ENOMEM; }
            batch = std::max<uint64_t>(1, nb / 2);
            if (!bp.no_screen) *bp.est_cand_per_tile = std::max(*bp.est_cand_per_tile, (double)ncand / nb);  // the counter runs past the capacity
            continue;
        }
        account(t, nb);
        ctx->stats.pairs_screened += ncand;
        t += nb;
        if (dump) {
            const size_t old = dump->size();
            dump->resize(old + ncand);
            CUDA_TRY(cudaMemcpy(dump->data() + old, ctx->d_cands.p, ncand * sizeof(Candidate), cudaMemcpyDeviceToHost));
            continue;
        }
        if (ncand) {
            const uint64_t fill = ctx->h_counters[1 + ctx->rec_cur];  // exact: every earlier statistics kernel has completed
            if (fill + ncand > ctx->rec_cap) {
                const int rc = rotate_records(ctx, fill, resident);
                if (rc) return rc;
            }
            CUDA_TRY(cudaEventRecord(ctx->ev2, ctx->stream));
            const unsigned grid = (unsigned)((ncand + STATS_PER_BLOCK - 1) / STATS_PER_BLOCK);
            uint8_t* recs = ctx->d_records[ctx->rec_cur].p;
            unsigned long long* rec_count = ctx->d_counters.p + 1 + ctx->rec_cur;
            static const int occ3 = [] { const char* e = tuning_env("TWKB_STATS_OCC"); return e ? atoi(e) : 3; }();
            if (prm.unphased)
                stats_kernel<true, 2><<<grid, STATS_THREADS, 0, ctx->stream>>>(ctx->d_cands.p, (uint32_t)ncand, ctx->d_meta.p, prm,
                                                                              ctx->d_lgamma.p, recs, ctx->rec_cap, rec_count);
            else if (occ3 == 3)
                stats_kernel<false, 3><<<grid, STATS_THREADS, 0, ctx->stream>>>(ctx->d_cands.p, (uint32_t)ncand, ctx->d_meta.p, prm,
                                                                               ctx->d_lgamma.p, recs, ctx->rec_cap, rec_count);
            else
                stats_kernel<false, 2><<<grid, STATS_THREADS, 0, ctx->stream>>>(ctx->d_cands.p, (uint32_t)ncand, ctx->d_meta.p, prm,
                                                                               ctx->d_lgamma.p, recs, ctx->rec_cap, rec_count);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(ctx->ev3, ctx->stream));
            ctx->stats_timing_pending = true;
            ctx->stats.stats_launches += 1;
        }
        // adapt: aim for a half-full candidate buffer
        if (!bp.no_screen && !tuning_env("TWKB_BATCH_TILES")) {
            const double per_tile = std::max(1.0, (double)ncand / nb);
            batch = std::max<uint64_t>(1, (uint64_t)(0.5 * ctx->cand_cap / per_tile));
            *bp.est_cand_per_tile = std::max(*bp.est_cand_per_tile, per_tile);
        }
    }
    return TWKB_OK;
}

// One pass over a sub-problem with the planes of `mode`.
// pair_filter: 0 = all pairs, 1 = only pairs with no missing variant, 2 = only pairs with one
// (auto mode passes; see compute()).
static int run_pass(Context* ctx, const Problem& pb_in, int mode, uint32_t pair_filter, bool resident, bool screen_off,
                    twkb_sink_fn sink, void* user, std::vector<Candidate>* dump) {
    // Rare-variant class active: the dense kernels see the [0, nD) x [0, nD) triangle only, the
    // sparse kernel every pair with a sparse member (second phase below).
    const bool sparse_phase = ctx->permuted && ctx->nS > 0;
    if (sparse_phase && (mode != MODE_PHASED_NOMISS || ctx->st.n_chunks != 1)) {
        ctx->err = "the resident matrix is ordered [dense | sparse] for the rare-variant path (sparse_max_words); "
                   "reload it to run unphased / chunked (-c) computations";
        return TWKB_ESTATE;
    }
    Problem pb = pb_in;
    if (sparse_phase) pb = Problem{0, ctx->nD, 0, ctx->nD, true};
    const uint32_t n_dense = sparse_phase ? ctx->nD : ctx->n_variants;
    int rc = ensure_planes(ctx, mode);
    if (rc) return rc;
    // Kernel choice. Tensor pipe (tcgen05) whenever the counts of the mode are a 0/1 contraction
    // whose fp32 accumulation is exact (2N < 2^24): 1 plane -> count_umma3_kernel<.,.,0>, masked
    // phased / unphased tables -> the planes variants (NP operand rows per variant). The LOP3+POPC
    // kernel serves explicit requests and 2N >= 2^24 with masks. The candidate dump of the tests
    // (twkb_debug_candidates) runs whatever kernel the settings select, so the raw counts of the
    // tcgen05 paths are checked directly, not only through the records that survive the screen.
    bool use_umma = false, use_fp4 = false;
    const bool planes_mode = mode != MODE_PHASED_NOMISS;
    // single mode (scalc): a handful of target rows against a neighbourhood -- the 128-row LOP3+POPC tiles, not 256 x 240 MMA tiles
    if (ctx->st.kernel != TWKB_KERNEL_POPC && !ctx->st.single && umma_supported()) {
        if (!planes_mode) use_umma = true;
        else use_umma = ctx->st.kernel != TWKB_KERNEL_UMMA && umma_fp4_possible(ctx->n_samples) && !tuning_env("TWKB_PLANES_POPC");
    }
    if ((ctx->st.kernel == TWKB_KERNEL_UMMA || ctx->st.kernel == TWKB_KERNEL_UMMA_FP4) && !use_umma && !ctx->st.single) {
        ctx->err = planes_mode ? "the int8 tensor-core kernel only serves phased data without missing genotypes (use AUTO or UMMA_FP4)"
                               : "tensor-core kernel requested but unavailable";
        return TWKB_EINVAL;
    }
    if (use_umma && !planes_mode) {
        // operand encoding: e2m1 (kind::mxf4, 2x the MAC rate of int8) whenever its fp32
        // accumulation is exact (2N < 2^24), int8 otherwise or on request
        use_fp4 = ctx->st.kernel != TWKB_KERNEL_UMMA && umma_fp4_possible(ctx->n_samples);
        if (const char* e = tuning_env("TWKB_UMMA_KIND")) {
            if (e[0] == 'i') use_fp4 = false;
        }
        if (ctx->st.kernel == TWKB_KERNEL_UMMA_FP4 && !use_fp4) {
            ctx->err = "TWKB_KERNEL_UMMA_FP4 needs 2N < 2^24 and the persistent kernel";
            return TWKB_EINVAL;
        }
    }
    if (use_umma && planes_mode) use_fp4 = true;
    uint32_t TI, TJ;
    if (use_umma && planes_mode) planes_tile(mode, TI, TJ);
    else tile_dims(mode, use_umma, use_fp4, TI, TJ);
    if (use_umma && n_dense >= 2) {
        if (planes_mode)
            rc = umma_prepare_planes(ctx->umma, mode, ctx->d_planes.p, ctx->K32, ctx->Mpad, ctx->n_variants, ctx->stream, ctx->err,
                                     &ctx->stats.other_launches);
        else  // only the dense variants are expanded: the operand is nD rows, not M
            rc = umma_prepare(ctx->umma, use_fp4, ctx->d_raw_data.p, ctx->raw_stride, n_dense, (n_dense + 255) / 256 * 256,
                              ctx->n_samples, ctx->stream, ctx->err, &ctx->stats.other_launches);
        if (rc) return rc;
    }
    // The tile plan depends only on the sub-problem, the tile shape, the window and the
    // partition: build it (and upload it) once and reuse it across runs.
    // (outside window mode the plan is a function of the ranges alone, so it survives a reload of a
    // matrix of the same shape: the end-to-end path does not rebuild and re-upload 3e5 tiles per call)
    // super-tile edge (in tiles) of the L2-friendly order. Short rows (C2: 2.5 KB of e2m1 per variant): 32 x 32 tiles of
    // operand rows are a 40 MB working set that stays in L2 (B200 sweep 16/24/32/48: 28.42 / 28.15 / 28.12 / 28.12 ms per C2
    // launch). Long rows (biobank scale: 500 KB per variant) never fit: what matters then is that the ~74 tiles that are
    // in flight TOGETHER (one per CTA pair) form a compact block, so that the K blocks they stream in near lock-step are
    // shared through L2 -- an 8 x 8 super-tile is about one such wave.
    uint32_t super = 32u;
    {
        const size_t row_bytes = ctx->umma.valid ? ctx->umma.Kbytes : (size_t)ctx->K32 * 4;
        if ((size_t)32 * (TI + TJ) * row_bytes > ((size_t)96 << 20)) super = 8u;  // B200, 1 M haplotypes: 32/16/12/9/8/6 -> 108.7 / 114.6 / 117.5 / 102.4 / 98.8 / 106.2 ms
    }
    if (const char* e = tuning_env("TWKB_SUPER")) super = (uint32_t)std::max(1, atoi(e));
    char keybuf[256];
    std::snprintf(keybuf, sizeof(keybuf), "%llu|%u-%u,%u-%u,%d|%ux%u/%u|w%d:%d:%d|p%d/%d|s%d",
                  (unsigned long long)(ctx->st.window ? ctx->matrix_epoch : 0ull),
                  pb.row_begin, pb.row_end, pb.col_begin, pb.col_end, (int)pb.diag, TI, TJ, super, (int)window_kind(ctx->st),
                  ctx->st.l_window, ctx->st.twk_block_size, ctx->st.part_index, ctx->st.part_count, ctx->st.shard_blocks);
    if (ctx->plan_key != keybuf) {
        if (n_dense >= 2) build_tiles(ctx, pb, TI, TJ, super, ctx->plan_tiles, &ctx->plan_pairs);
        else { ctx->plan_tiles.clear(); ctx->plan_pairs = 0; }
        CUDA_TRY(ctx->d_tiles.alloc(std::max<size_t>(ctx->plan_tiles.size(), 1)));
        if (!ctx->plan_tiles.empty())
            CUDA_TRY(cudaMemcpyAsync(ctx->d_tiles.p, ctx->plan_tiles.data(), ctx->plan_tiles.size() * sizeof(uint2),
                                     cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->plan_key = keybuf;
    }
    if (sparse_phase) {
        std::snprintf(keybuf, sizeof(keybuf), "%llu|w%d:%d:%d|p%d/%d|s%d", (unsigned long long)ctx->matrix_epoch, (int)window_kind(ctx->st),
                      ctx->st.l_window, ctx->st.twk_block_size, ctx->st.part_index, ctx->st.part_count, ctx->st.shard_blocks);
        if (ctx->sp_plan_key != keybuf) {
            build_sparse_tiles(ctx, ctx->sp_plan_tiles, &ctx->sp_plan_pairs);
            ctx->sp_tile_words_prefix.assign(ctx->sp_plan_tiles.size() + 1, 0);
            for (size_t i = 0; i < ctx->sp_plan_tiles.size(); ++i) {
                const uint2 tl = ctx->sp_plan_tiles[i];
                const uint32_t s0 = tl.x - ctx->nD, s1 = std::min<uint32_t>(s0 + SP_ROWS, ctx->nS);
                const uint64_t cols = std::min<uint32_t>(tl.y + SP_TJ, ctx->n_variants) - tl.y;
                ctx->sp_tile_words_prefix[i + 1] = ctx->sp_tile_words_prefix[i] + (uint64_t)(ctx->h_sp_off[s1] - ctx->h_sp_off[s0]) * cols;
            }
            CUDA_TRY(ctx->d_sp_tiles.alloc(std::max<size_t>(ctx->sp_plan_tiles.size(), 1)));
            if (!ctx->sp_plan_tiles.empty())
                CUDA_TRY(cudaMemcpyAsync(ctx->d_sp_tiles.p, ctx->sp_plan_tiles.data(), ctx->sp_plan_tiles.size() * sizeof(uint2),
                                         cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            ctx->sp_plan_key = keybuf;
        }
    }
    const std::vector<uint2>& tiles = ctx->plan_tiles;
    if (ctx->st.part_count > 1 && !ctx->st.window) ctx->stats.pairs_visited = ctx->plan_pairs + (sparse_phase ? ctx->sp_plan_pairs : 0);
    ctx->stats.kernel_used = use_umma ? (use_fp4 ? TWKB_KERNEL_UMMA_FP4 : TWKB_KERNEL_UMMA) : TWKB_KERNEL_POPC;
    ctx->stats.n_planes = ctx->np;
    ctx->stats.sparse_variants = sparse_phase ? ctx->nS : 0;
    const uint64_t tile_pairs = (uint64_t)TI * TJ;
    rc = ensure_work_buffers(ctx, std::max<uint64_t>(tile_pairs, (uint64_t)SP_ROWS * SP_TJ));
    if (rc) return rc;
    if (tiles.empty() && !(sparse_phase && !ctx->sp_plan_tiles.empty())) return TWKB_OK;

    DevParams prm = make_params(ctx, pb);
    prm.pair_filter = pair_filter;
    CountArgs args{};
    args.planes = ctx->d_planes.p;
    args.K32 = ctx->K32;
    args.Mpad = ctx->Mpad;
    args.meta = ctx->d_meta.p;
    args.plane_popc = ctx->d_plane_popc.p;
    args.blocks = DevBlocks{ctx->d_blk_of.p, ctx->d_blk_first.p, ctx->d_blk_last.p, ctx->d_blk_prune.p};
    args.cands = ctx->d_cands.p;
    args.cand_count = ctx->d_counters.p;
    args.cand_capacity = ctx->cand_cap;
    args.row_begin = pb.row_begin; args.row_end = pb.row_end;
    args.col_begin = pb.col_begin; args.col_end = pb.col_end;
    args.screen_off = screen_off ? 1u : 0u;
#ifdef TWKB_PROFILING
    if (const char* e = getenv("TWKB_DEBUG_FLAGS")) args.debug_flags = (uint32_t)atoi(e);
#endif

    const bool no_screen = screen_off || !(ctx->st.minR2 > 0.0);
    CUDA_TRY(cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
    ctx->rec_cur = 0;
    ctx->stats_timing_pending = false;
    if (!resident && !dump) flusher_begin(ctx, sink, user);

    // ---- phase 1: dense x dense tiles
    if (!tiles.empty()) {
        BatchPlan bp;
        bp.n_tiles = tiles.size();
        bp.tile_pairs = tile_pairs;
        bp.no_screen = no_screen;
        bp.est_cand_per_tile = &ctx->est_cand_per_tile;
        auto launch = [&](size_t t, uint32_t nb) -> cudaError_t {
            args.tiles = ctx->d_tiles.p + t;
            if (use_umma) return umma_launch(ctx->umma, args, prm, nb, ctx->stream);
            switch (mode) {
                case MODE_PHASED_NOMISS: return launch_popc<0>(ctx, args, prm, nb);
                case MODE_PHASED_MISS: return launch_popc<1>(ctx, args, prm, nb);
                case MODE_UNPHASED_NOMISS: return launch_popc<2>(ctx, args, prm, nb);
                default: return launch_popc<3>(ctx, args, prm, nb);
            }
        };
        auto account = [&](size_t, uint32_t nb) {
            if (use_umma) ctx->stats.mma_macs += (uint64_t)nb * (planes_mode ? 256ull * 240ull : tile_pairs) * ctx->umma.Kelems;
            else ctx->stats.word_ops += (uint64_t)nb * tile_pairs * ctx->K32 * ctx->np * ctx->np;
        };
        rc = run_batches(ctx, bp, prm, launch, account, resident, dump);
        if (rc) return rc;
    }
    // ---- phase 2: every pair with a sparse member (list kernel)
    if (sparse_phase && !ctx->sp_plan_tiles.empty()) {
        DevParams sprm = prm;
        sprm.diag = 0;  // the kernel applies its own "each pair once" rule
        CountArgs sargs = args;
        sargs.row_begin = 0; sargs.row_end = ctx->n_variants;
        sargs.col_begin = 0; sargs.col_end = ctx->n_variants;
        SparseArgs sp{ctx->d_sp_off.p, ctx->d_sp_ent.p, ctx->d_orig.p, ctx->nD, ctx->nS};
        BatchPlan bp;
        bp.n_tiles = ctx->sp_plan_tiles.size();
        bp.tile_pairs = (uint64_t)SP_ROWS * SP_TJ;
        bp.no_screen = no_screen;
        bp.sparse = true;
        bp.est_cand_per_tile = &ctx->est_cand_per_sp_tile;
        auto launch = [&](size_t t, uint32_t nb) -> cudaError_t {
            sargs.tiles = ctx->d_sp_tiles.p + t;
            if (no_screen) count_sparse_kernel<false><<<nb, SP_THREADS, 0, ctx->stream>>>(sargs, sp, sprm);
            else count_sparse_kernel<true><<<nb, SP_THREADS, 0, ctx->stream>>>(sargs, sp, sprm);
            return cudaGetLastError();
        };
        // word operations actually issued: entries of the tile's rows x columns of the tile
        std::vector<uint64_t>& pre = ctx->sp_tile_words_prefix;
        auto account = [&](size_t t, uint32_t nb) { ctx->stats.sparse_word_ops += pre[t + nb] - pre[t]; };
        rc = run_batches(ctx, bp, sprm, launch, account, resident, dump);
        if (rc) return rc;
    }
    if (!dump) {
        // the last statistics kernel: its record count, then the final drain
        CUDA_TRY(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        rc = read_stats_timing(ctx);
        if (rc) return rc;
        const uint64_t fill = ctx->h_counters[1 + ctx->rec_cur];
        ctx->stats.records_out += fill;
        if (resident) {
            rc = consume_records(ctx, ctx->rec_cur, fill);
            if (rc) return rc;
        }
        if (!resident) {
            if (fill) flusher_submit(ctx, ctx->rec_cur, fill);
            rc = flusher_wait(ctx, -1);
            ctx->stats.bytes_d2h += ctx->flusher->bytes_d2h;
            if (rc) return rc;
        }
    }
    return TWKB_OK;
}

static int compute_impl(Context* ctx, bool resident, twkb_sink_fn sink, void* user, bool screen_off,
                        std::vector<Candidate>* dump) {
    if (!ctx->loaded) { ctx->err = "twkb_compute before twkb_load_matrix"; return TWKB_ESTATE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const auto t0 = std::chrono::steady_clock::now();
    const double ms_h2d = ctx->stats.ms_h2d;
    const uint64_t b_h2d = ctx->stats.bytes_h2d;
    ctx->stats = twkb_stats{};
    ctx->stats.ms_h2d = ms_h2d;
    ctx->stats.bytes_h2d = b_h2d;
    ctx->stats.ms_decode_kernel = ctx->ms_decode;
    CUDA_TRY(cudaEventRecord(ctx->ev_begin, ctx->stream));
    Problem pb;
    int rc = select_problem(ctx, pb);
    if (rc) return rc;
    if (ctx->st.single) {
        // twk_ld_balancer::BuildSingleSite + CalculateSingle: block 0 (the targets) against itself (i < j) and against
        // every other block -- rows = targets, columns = everything, each pair once with the target first
        const uint32_t nT = (uint32_t)ctx->st.single_targets;
        if (nT == 0 || nT > ctx->n_variants) { ctx->err = "no data found for reference"; return TWKB_EINVAL; }
        if (nT == ctx->n_variants && nT < 2) { ctx->err = "no surrounding variants"; return TWKB_EINVAL; }
        pb = Problem{0, nT, 0, ctx->n_variants, true};
    }
    if (ctx->st.window) {
        rc = build_blocks(ctx);
        if (rc) return rc;
    }
    ctx->stats.pairs_visited = visited_pairs(ctx, pb);
    if (ctx->st.single) {
        const uint64_t nT = pb.row_end, M = ctx->n_variants;
        ctx->stats.pairs_visited = (nT * nT - nT) / 2 + nT * (M - nT);
    }
    if (ctx->st.part_count > 1) {
        // every part reports its share of the visited pairs (tiles are dealt round-robin)
        const uint64_t v = ctx->stats.pairs_visited, n = ctx->st.part_count, r = ctx->st.part_index;
        ctx->stats.pairs_visited = v / n + (r < v % n ? 1 : 0);
    }
    // (single mode always picks the comparator per pair, whatever -p / -u say: twk_ld_slave::Start, ld_engine.cpp:1826-1830)
    if (((ctx->st.force_phased || ctx->st.forced_unphased) && !ctx->st.single) || !ctx->any_missing) {
        // -p, -u, or auto mode on complete data (auto => phased for every pair, ld_engine.cpp:2775-2790)
        rc = run_pass(ctx, pb, wanted_mode(ctx, ctx->st.forced_unphased && !ctx->st.single), 0, resident, screen_off, sink, user, dump);
    } else {
        // Auto mode (ld_engine.cpp:2737-2838): a pair takes the unphased path iff either variant
        // has missing alleles (an != 0), else the phased path; gt_phase is never consulted (Q4).
        // Two passes over the same tile plan family: the rows of complete variants carry no mask
        // bits, so the no-missing phased planes (and the tensor-core kernel) are exact for pass 1.
        rc = run_pass(ctx, pb, MODE_PHASED_NOMISS, 1, resident, screen_off, sink, user, dump);
        if (rc == TWKB_OK) rc = run_pass(ctx, pb, MODE_UNPHASED_MISS, 2, resident, screen_off, sink, user, dump);
    }
    if (rc != TWKB_OK && ctx->flusher) {  // an aborted pass: let the drain thread release its buffers before returning
        const std::string keep = ctx->err;
        flusher_wait(ctx, -1);
        ctx->err = keep;
    }
    if (rc == TWKB_OK) {
        CUDA_TRY(cudaEventRecord(ctx->ev_end, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
        ctx->stats.ms_device_total = ms;
    }
    ctx->stats.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// Rare-variant classification (reference: twk_igt_list::Build keeps, per variant, the offsets of
// the registers that hold an alt allele, include/core.h:601-632, and PhasedListVector walks the
// list of the sparser variant, ld_engine.cpp:185-267). Variants whose row has <= T non-zero
// 32-bit words become the sparse class: the resident rows are re-ordered [dense | sparse] (file
// order inside each class) and the sparse rows get a CSR list of (word, value) entries.
static int classify_sparse(Context* ctx) {
    const uint32_t M = ctx->n_variants;
    ctx->permuted = false;
    ctx->nD = M;
    ctx->nS = 0;
    ctx->sparse_T = 0;
    ctx->sp_entries = 0;
    ctx->sp_plan_key.clear();
    ctx->est_cand_per_sp_tile = -1.0;
    if (ctx->h_orig.size() != M || !ctx->h_orig_identity) {  // (kept across loads of the same shape: 566,000 stores otherwise)
        ctx->h_orig.resize(M);
        for (uint32_t x = 0; x < M; ++x) ctx->h_orig[x] = x;
        ctx->h_orig_identity = true;
    }
    const twkb_settings& st = ctx->st;
    const uint32_t n_bits = 2 * ctx->n_samples;
    const uint32_t K32raw = (n_bits + 31) / 32;
    int64_t T = 0;
    if (st.sparse_max_words > 0) T = st.sparse_max_words;
    else if (st.sparse_max_words == 0 && st.kernel == TWKB_KERNEL_AUTO && n_bits >= 32768u) T = K32raw / 64;
    if (st.sparse_max_words >= 0) {
        if (const char* e = tuning_env("TWKB_SPARSE_T")) T = atoll(e);
    }
    if (T <= 0 || ctx->any_missing || st.forced_unphased || st.n_chunks != 1 || M < 2 || st.single) return TWKB_OK;
    DevBuf<uint32_t> d_nnz;
    CUDA_TRY(d_nnz.alloc(M));
    row_nnz32_kernel<<<(M + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_raw_data.p, ctx->raw_stride, M, n_bits, d_nnz.p);
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> nnz(M);
    CUDA_TRY(cudaMemcpyAsync(nnz.data(), d_nnz.p, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    d_nnz.release();
    ctx->stats.other_launches += 1;
    uint32_t nS = 0;
    for (uint32_t v = 0; v < M; ++v) nS += (nnz[v] <= (uint64_t)T) ? 1u : 0u;
    // an automatic threshold only pays when a real share of the variants is rare
    if (nS == 0 || (st.sparse_max_words == 0 && !tuning_env("TWKB_SPARSE_T") && nS < std::max<uint32_t>(256u, M / 20))) return TWKB_OK;
    const uint32_t nD = M - nS;
    ctx->h_orig_identity = false;
    uint32_t d = 0, sidx = nD;
    ctx->h_sp_off.assign(nS + 1, 0);
    for (uint32_t v = 0; v < M; ++v) {
        if (nnz[v] <= (uint64_t)T) {
            ctx->h_sp_off[sidx - nD + 1] = ctx->h_sp_off[sidx - nD] + nnz[v];
            ctx->h_orig[sidx++] = v;
        } else {
            ctx->h_orig[d++] = v;
        }
    }
    ctx->sp_entries = ctx->h_sp_off[nS];
    // resident order -> device, rows gathered into the new order
    std::vector<uint32_t> orig_pad(ctx->Mpad);
    for (uint32_t x = 0; x < ctx->Mpad; ++x) orig_pad[x] = x < M ? ctx->h_orig[x] : x;
    CUDA_TRY(ctx->d_orig.alloc(ctx->Mpad));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_orig.p, orig_pad.data(), (size_t)ctx->Mpad * 4, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<uint64_t> gathered;
    CUDA_TRY(gathered.alloc((size_t)M * ctx->raw_stride));
    gather_rows_kernel<<<M, 128, 0, ctx->stream>>>(ctx->d_raw_data.p, gathered.p, ctx->d_orig.p, ctx->raw_stride, M);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->d_raw_data.p, gathered.p);
    std::swap(ctx->d_raw_data.n, gathered.n);
    gathered.release();
    // CSR entries
    CUDA_TRY(ctx->d_sp_off.alloc(nS + 1));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_sp_off.p, ctx->h_sp_off.data(), (size_t)(nS + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx->d_sp_ent.alloc(std::max<uint64_t>(ctx->sp_entries, 1)));
    build_sparse_entries_kernel<<<(nS + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_raw_data.p, ctx->raw_stride, nD, nS, n_bits, ctx->d_sp_off.p,
                                                                        ctx->d_sp_ent.p);
    CUDA_TRY(cudaGetLastError());
    ctx->stats.other_launches += 2;
    ctx->permuted = true;
    ctx->nD = nD;
    ctx->nS = nS;
    ctx->sparse_T = (uint32_t)T;
    return TWKB_OK;
}

// First half of every load: shape, file-order metadata, missing-data flag.
static int load_begin_shapes(Context* ctx, uint32_t n_samples, uint32_t n_variants, size_t stride, const twkb_variant* meta) {
    if (!meta || n_samples == 0 || n_variants == 0) { ctx->err = "null/empty matrix"; return TWKB_EINVAL; }
    if (stride * 64 < 2 * (size_t)n_samples) { ctx->err = "row_stride_words too small for n_samples"; return TWKB_EINVAL; }
    if (2 * (uint64_t)n_samples >= (1ull << 31)) { ctx->err = "n_samples too large"; return TWKB_EINVAL; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->loaded = false;
    ctx->mode = -1;
    ctx->umma.valid = false;
    ctx->matrix_epoch += 1;
    // the survivor rate the previous run measured sizes the first batch of the next one; it is kept across a reload of a
    // matrix of the same shape (a file processed chunk by chunk, the end-to-end bench): a guess that turns out too low
    // costs one retry of the batch (run_batches), a reset costs an extra batch boundary on every run
    if (n_samples != ctx->n_samples || n_variants != ctx->n_variants) ctx->est_cand_per_tile = -1.0;
    ctx->n_samples = n_samples;
    ctx->n_variants = n_variants;
    ctx->Mpad = (n_variants + 255) / 256 * 256;
    ctx->raw_stride = stride;
    ctx->h_meta_orig.clear();
    ctx->permuted = false;
    ctx->file_blocks.clear();
    return TWKB_OK;
}
// Host side of a load, per variant: (a) the copy of the caller's metadata (file order; the scheduler consults it for window
// rules, chunks and the rare-variant class) and (b) the 16-byte device records, staged in pinned memory. At 566,000
// variants (8-GPU weak scaling) that is 18 MB read twice and 27 MB written -- 8 ms when done serially in front of the
// upload, with 8 ranks competing for the host's memory bandwidth. The matrix loads therefore enqueue the row upload (and
// the all-gather) FIRST, then run (a) and (b) behind the transfer (host_meta_work).
static void copy_meta_range(HostVar* dst, const twkb_variant* meta, uint32_t v0, uint32_t v1) {
    for (uint32_t v = v0; v < v1; ++v) dst[v] = HostVar{meta[v].rid, meta[v].pos};
}
static void copy_meta(Context* ctx, const twkb_variant* meta) {
    ctx->h_meta.resize(ctx->n_variants);
    copy_meta_range(ctx->h_meta.data(), meta, 0, ctx->n_variants);
}
static int ensure_dm(Context* ctx) {
    if (ctx->h_dm_cap < ctx->Mpad) {
        if (ctx->h_dm) cudaFreeHost(ctx->h_dm);
        ctx->h_dm = nullptr;
        ctx->h_dm_cap = 0;
        CUDA_TRY(cudaMallocHost((void**)&ctx->h_dm, (size_t)ctx->Mpad * sizeof(DevVariant)));
        ctx->h_dm_cap = ctx->Mpad;
    }
    return TWKB_OK;
}
// device records of resident variants [v0, v1) from the caller's (file-order) metadata; perm (nullable) = file index of a
// resident variant. Returns whether any of them has missing genotypes.
static bool fill_dm_range(DevVariant* dm, const twkb_variant* meta, uint32_t v0, uint32_t v1, const uint32_t* perm = nullptr) {
    bool miss = false;
    for (uint32_t v = v0; v < v1; ++v) {
        const twkb_variant& m = meta[perm ? perm[v] : v];
        dm[v].pos = m.pos;
        dm[v].ac = m.ac;
        dm[v].rid = m.rid;
        dm[v].flags = (m.an ? VF_HAS_MISSING : 0u) | (m.hwe < 1e-4 ? VF_BAD_HWE : 0u) | (m.gt_missing ? VF_GT_MISSING : 0u);
        miss = miss || m.gt_missing || m.an;
    }
    return miss;
}
static int fill_dm(Context* ctx, const twkb_variant* meta, bool* any_missing, const uint32_t* perm = nullptr) {
    const uint32_t n_variants = ctx->n_variants;
    const int rc = ensure_dm(ctx);
    if (rc) return rc;
    const bool miss = fill_dm_range(ctx->h_dm, meta, 0, n_variants, perm);
    std::memset(ctx->h_dm + n_variants, 0, (size_t)(ctx->Mpad - n_variants) * sizeof(DevVariant));
    if (any_missing) *any_missing = miss;
    return TWKB_OK;
}
// (a) + (b) of the matrix loads, behind the transfer. Up to META_TEAM_MIN variants: (a) on a helper thread, (b) on the
// calling thread -- the transfer (2.4 ms for the 131 MB of C2) outlasts both. Beyond it the passes are what the load waits
// for: at 8 ranks the weak-scaled matrix has 566,000 variants but a rank uploads only 1/8 of the rows (46 MB, < 1 ms) and
// the load still took 5.8 ms. There one pass per chunk of variants runs on a small team of threads (the calling thread is
// one of them; the ranks of a node share the host's cores).
constexpr uint32_t META_TEAM_MIN = 400000;
static int host_meta_work(Context* ctx, const twkb_variant* meta, bool* any_missing) {
    const uint32_t M = ctx->n_variants;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const unsigned team = std::min(4u, std::max(1u, hw / (unsigned)std::max(1, ctx->comm_size)));
    if (M < META_TEAM_MIN || team < 3) {
        std::thread helper(copy_meta, ctx, meta);
        const int rc = fill_dm(ctx, meta, any_missing);
        helper.join();
        return rc;
    }
    const int rc = ensure_dm(ctx);
    if (rc) return rc;
    ctx->h_meta.resize(M);
    std::vector<uint8_t> miss(team, 0);
    auto work = [&](unsigned k) {
        const uint32_t v0 = (uint32_t)((uint64_t)M * k / team), v1 = (uint32_t)((uint64_t)M * (k + 1) / team);
        copy_meta_range(ctx->h_meta.data(), meta, v0, v1);
        miss[k] = fill_dm_range(ctx->h_dm, meta, v0, v1) ? 1 : 0;
    };
    std::vector<std::thread> helpers;
    for (unsigned k = 1; k < team; ++k) helpers.emplace_back(work, k);
    work(0);
    for (std::thread& t : helpers) t.join();
    std::memset(ctx->h_dm + M, 0, (size_t)(ctx->Mpad - M) * sizeof(DevVariant));
    bool any = false;
    for (uint8_t m : miss) any = any || m;
    if (any_missing) *any_missing = any;
    return TWKB_OK;
}
static void load_meta(Context* ctx, const twkb_variant* meta) {
    const uint32_t n_variants = ctx->n_variants;
    copy_meta(ctx, meta);
    ctx->any_missing = false;
    for (uint32_t v = 0; v < n_variants; ++v)
        if (meta[v].gt_missing || meta[v].an) { ctx->any_missing = true; break; }
}
static int load_begin(Context* ctx, uint32_t n_samples, uint32_t n_variants, size_t stride, const twkb_variant* meta) {
    const int rc = load_begin_shapes(ctx, n_samples, n_variants, stride, meta);
    if (rc) return rc;
    load_meta(ctx, meta);
    return TWKB_OK;
}

// Second half: d_raw_data (+ d_raw_mask) hold the reference-layout rows, however they got there
// (ev0 was recorded before the upload started).
static int load_finish(Context* ctx, const twkb_variant* meta, bool dm_ready = false) {
    const uint32_t n_samples = ctx->n_samples, n_variants = ctx->n_variants;
    // rare-variant class: may re-order the resident rows [dense | sparse]
    {
        const int rc_sp = classify_sparse(ctx);
        if (rc_sp) return rc_sp;
    }
    if (ctx->permuted) {  // resident order != file order: keep both
        ctx->h_meta_orig = ctx->h_meta;
        for (uint32_t x = 0; x < n_variants; ++x) ctx->h_meta[x] = ctx->h_meta_orig[ctx->h_orig[x]];
    }
    // device metadata, staged in pinned memory so that the copy is asynchronous (already filled by the matrix loads
    // while the rows were in flight, unless the rare-variant class re-ordered the variants: then in resident order,
    // from the caller's file-order metadata through h_orig)
    if (!dm_ready || ctx->permuted) {
        const int rc_dm = fill_dm(ctx, meta, nullptr, ctx->permuted ? ctx->h_orig.data() : nullptr);
        if (rc_dm) return rc_dm;
    }
    DevVariant* dm = ctx->h_dm;
    CUDA_TRY(ctx->d_meta.alloc(ctx->Mpad));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_meta.p, dm, (size_t)ctx->Mpad * sizeof(DevVariant), cudaMemcpyHostToDevice, ctx->stream));
    // log-factorial table from the host libm: lg[n] = lgamma(n+1), the exact values the
    // reference's lbinom() (lib/fisher_math.cpp:183-187) obtains from glibc. It depends on the sample count only:
    // kept across loads (1 M haplotypes = 1 M lgamma calls, ~50 ms of host time).
    ctx->lgamma_len = 2 * n_samples + 64;
    size_t lg_bytes = 0;
    if (ctx->lgamma_ready != ctx->lgamma_len) {
        std::vector<double> lg(ctx->lgamma_len);
        for (uint32_t n = 0; n < ctx->lgamma_len; ++n) lg[n] = lgamma((double)n + 1.0);
        CUDA_TRY(ctx->d_lgamma.alloc(ctx->lgamma_len));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_lgamma.p, lg.data(), lg.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // lg is a local
        ctx->lgamma_ready = ctx->lgamma_len;
        lg_bytes = lg.size() * 8;
    }
    ctx->stats.bytes_h2d += (size_t)ctx->Mpad * sizeof(DevVariant) + lg_bytes;
    ctx->loaded = true;
    // build the planes the configured mode needs right away so that the upload cost
    // (H2D + transpose) is accounted to the load, not to the first compute
    const bool unph = ctx->st.forced_unphased;
    int rc = ensure_planes(ctx, wanted_mode(ctx, unph));
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.ms_h2d = ms;
    ctx->stats.ms_decode_kernel = ctx->ms_decode;
    return TWKB_OK;
}

static int load_common(Context* ctx, uint32_t n_samples, uint32_t n_variants, const uint64_t* data, const uint64_t* mask,
                       size_t stride, const twkb_variant* meta, bool device_src) {
    if (!data) { ctx->err = "null/empty matrix"; return TWKB_EINVAL; }
    int rc = load_begin_shapes(ctx, n_samples, n_variants, stride, meta);
    if (rc) return rc;
    ctx->ms_decode = 0.0;
    const size_t words = (size_t)n_variants * stride;
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    CUDA_TRY(ctx->d_raw_data.alloc(words));
    const cudaMemcpyKind kind = device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_raw_data.p, data, words * 8, kind, ctx->stream));
    ctx->stats.bytes_h2d = device_src ? 0 : words * 8;
    // host metadata work behind the transfer: copy on a helper thread, device records (+ the missing-data flag) here
    bool miss = false;
    rc = host_meta_work(ctx, meta, &miss);
    if (rc) return rc;
    ctx->any_missing = miss;
    if (miss && !mask) { ctx->err = "variants flagged missing but mask_bits is NULL"; return TWKB_EINVAL; }
    if (miss) {
        CUDA_TRY(ctx->d_raw_mask.alloc(words));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_raw_mask.p, mask, words * 8, kind, ctx->stream));
        if (!device_src) ctx->stats.bytes_h2d += words * 8;
    }
    return load_finish(ctx, meta, true);
}

// twkb_load_runs: run-length records -> resident rows, decoded by decode_runs_kernel.
static int load_runs(Context* ctx, uint32_t n_samples, uint32_t n_variants, const uint8_t* bytes, size_t n_bytes,
                     const twkb_run_desc* desc, const twkb_variant* meta) {
    if (!bytes || !desc) { ctx->err = "null run buffer"; return TWKB_EINVAL; }
    const uint64_t H = 2ull * n_samples;
    const size_t stride = ((H + 63) / 64 + 1) / 2 * 2;  // 128-bit aligned rows, as twk_igt_vec (include/core.h:52-60)
    int rc = load_begin(ctx, n_samples, n_variants, stride, meta);
    if (rc) return rc;
    for (uint32_t v = 0; v < n_variants; ++v) {
        const twkb_run_desc& d = desc[v];
        if ((d.width != 1 && d.width != 2 && d.width != 4) || d.miss > 1 || d.offset > n_bytes ||
            (uint64_t)d.n_runs * d.width > n_bytes - d.offset) {
            ctx->err = "illegal gt primitive type / truncated runs (variant " + std::to_string(v) + ")";
            return TWKB_EINVAL;
        }
    }
    const size_t words = (size_t)n_variants * stride;
    const bool trace = tuning_env("TWKB_TRACE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        std::fprintf(stderr, "[twkb trace] load_runs %-14s %8.3f ms\n", what,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count());
    };
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    DevBuf<uint8_t> d_bytes;
    DevBuf<twkb_run_desc> d_desc;
    DevBuf<uint32_t> d_err;
    CUDA_TRY(d_bytes.alloc(n_bytes + 16));
    CUDA_TRY(d_desc.alloc(n_variants));
    CUDA_TRY(d_err.alloc(2));
    CUDA_TRY(ctx->d_raw_data.alloc(words));
    if (ctx->any_missing) CUDA_TRY(ctx->d_raw_mask.alloc(words));
    mark("alloc");
    CUDA_TRY(cudaMemcpyAsync(d_bytes.p, bytes, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    mark("h2d runs");
    CUDA_TRY(cudaMemcpyAsync(d_desc.p, desc, (size_t)n_variants * sizeof(twkb_run_desc), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_raw_data.p, 0, words * 8, ctx->stream));
    if (ctx->any_missing) CUDA_TRY(cudaMemsetAsync(ctx->d_raw_mask.p, 0, words * 8, ctx->stream));
    const uint32_t h_err0[2] = {0u, 0xffffffffu};
    CUDA_TRY(cudaMemcpyAsync(d_err.p, h_err0, sizeof(h_err0), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev2, ctx->stream));
    decode_runs_kernel<<<(n_variants + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, 0, ctx->stream>>>(
        d_bytes.p, d_desc.p, n_variants, (uint32_t)H, reinterpret_cast<uint32_t*>(ctx->d_raw_data.p),
        ctx->any_missing ? reinterpret_cast<uint32_t*>(ctx->d_raw_mask.p) : nullptr, stride * 2, d_err.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev3, ctx->stream));
    mark("decode");
    uint32_t h_err[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h_err, d_err.p, sizeof(h_err), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    d_bytes.release();
    d_desc.release();
    d_err.release();
    mark("free");
    {
        float ms_dec = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms_dec, ctx->ev2, ctx->ev3));
        ctx->ms_decode = ms_dec;
    }
    ctx->stats.other_launches += 1;
    ctx->stats.bytes_h2d = n_bytes + (size_t)n_variants * sizeof(twkb_run_desc);
    if (h_err[0]) {
        ctx->err = "run lengths do not cover all samples (variant " + std::to_string(h_err[1]) + ")";
        return TWKB_EINVAL;
    }
    rc = load_finish(ctx, meta);
    mark("finish");
    return rc;
}

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            ctx->err = std::string(#expr) + ": " + nccl_api().GetErrorString(_r);                   \
            return TWKB_ECUDA;                                                                      \
        }                                                                                           \
    } while (0)

// In-place all-gather of the ranks' row slices: slice k = rows [k S, (k + 1) S), S = ceil(M / N); d_rows holds N S rows
// (the rows past M are padding). One collective over NVLink / NVSwitch; every rank issues the same call.
// (A first version exchanged chunks with grouped ncclBroadcast calls behind the upload: 362 MB took 10 ms at 8 ranks.)
static int exchange_slices(Context* ctx, uint64_t* d_rows, size_t stride, uint32_t M) {
    const NcclApi& nc = nccl_api();
    const uint64_t S = ((uint64_t)M + ctx->comm_size - 1) / ctx->comm_size;
    NCCL_TRY(nc.AllGather(d_rows + (size_t)ctx->comm_rank * S * stride, d_rows, (size_t)S * stride, ncclUint64, ctx->comm, ctx->stream));
    return TWKB_OK;
}

static int ensure_chunk_events(Context* ctx, int n) {
    while ((int)ctx->chunk_events.size() < n) {
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->chunk_events.push_back(ev);
    }
    return TWKB_OK;
}

// twkb_load_matrix_sliced: this rank's rows go up over its own PCIe link (copy_stream), then one in-place
// all-gather over NVLink completes the matrix on every GPU.
static int load_matrix_sliced(Context* ctx, uint32_t n_samples, uint32_t n_variants, const uint64_t* slice_data, const uint64_t* slice_mask,
                              size_t stride, const twkb_variant* meta) {
    if (!ctx->comm || ctx->comm_size <= 1) return load_common(ctx, n_samples, n_variants, slice_data, slice_mask, stride, meta, false);
    int rc = load_begin_shapes(ctx, n_samples, n_variants, stride, meta);
    if (rc) return rc;
    ctx->ms_decode = 0.0;
    uint32_t b, e;
    comm_slice(n_variants, ctx->comm_rank, ctx->comm_size, b, e);
    const uint64_t rows = e - b;
    if (rows && !slice_data) { ctx->err = "null/empty matrix slice"; return TWKB_EINVAL; }
    const uint64_t S = ((uint64_t)n_variants + ctx->comm_size - 1) / ctx->comm_size;
    const size_t words = (size_t)S * ctx->comm_size * stride;  // padded to N equal slices
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    CUDA_TRY(ctx->d_raw_data.alloc(words));
    rc = ensure_chunk_events(ctx, 2);
    if (rc) return rc;
    // own slice over this GPU's PCIe link (copy_stream), padding rows of a short slice zeroed, then the collective
    auto upload = [&](uint64_t* d_rows, const uint64_t* h_rows) -> int {
        const size_t base = (size_t)S * ctx->comm_rank;  // == b unless the slice is empty
        if (rows) CUDA_TRY(cudaMemcpyAsync(d_rows + base * stride, h_rows, rows * stride * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (rows < S) CUDA_TRY(cudaMemsetAsync(d_rows + (base + rows) * stride, 0, (S - rows) * stride * 8, ctx->copy_stream));
        return TWKB_OK;
    };
    const bool trace = tuning_env("TWKB_TRACE") != nullptr;  // development aid: synchronous phase times on stderr
    const auto tr0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        std::fprintf(stderr, "[twkb trace] rank %d load_sliced %-12s %8.3f ms\n", ctx->comm_rank, what,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count());
    };
    rc = upload(ctx->d_raw_data.p, slice_data);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->chunk_events[0], ctx->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->chunk_events[0], 0));
    rc = exchange_slices(ctx, ctx->d_raw_data.p, stride, n_variants);
    if (rc) return rc;
    // host metadata work behind the transfer + collective: copy on a helper thread, device records here. The
    // missing-data flag comes out of the same pass; it is identical on every rank (same metadata), so all ranks
    // agree on whether a mask exchange follows.
    bool miss = false;
    rc = host_meta_work(ctx, meta, &miss);
    if (rc) return rc;
    ctx->any_missing = miss;
    if (miss && rows && !slice_mask) { ctx->err = "variants flagged missing but mask_bits is NULL"; return TWKB_EINVAL; }
    mark("data + meta");
    if (miss) {
        CUDA_TRY(ctx->d_raw_mask.alloc(words));
        rc = upload(ctx->d_raw_mask.p, slice_mask);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(ctx->chunk_events[1], ctx->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->chunk_events[1], 0));
        rc = exchange_slices(ctx, ctx->d_raw_mask.p, stride, n_variants);
        if (rc) return rc;
    }
    ctx->stats.bytes_h2d = rows * stride * 8 * (ctx->any_missing ? 2 : 1);
    rc = load_finish(ctx, meta, true);
    mark("finish");
    return rc;
}

// twkb_load_runs_sliced: this rank uploads the run words of ITS variants only, decodes them on the device
// into its rows, then the rows are exchanged. The coverage check of the decoder is exchanged too, so that
// every rank fails (or succeeds) together.
static int load_runs_sliced(Context* ctx, uint32_t n_samples, uint32_t n_variants, const uint8_t* bytes, size_t n_bytes,
                            const twkb_run_desc* desc, const twkb_variant* meta) {
    if (!ctx->comm || ctx->comm_size <= 1) return load_runs(ctx, n_samples, n_variants, bytes, n_bytes, desc, meta);
    if (!bytes || !desc) { ctx->err = "null run buffer"; return TWKB_EINVAL; }
    const uint64_t H = 2ull * n_samples;
    const size_t stride = ((H + 63) / 64 + 1) / 2 * 2;
    int rc = load_begin(ctx, n_samples, n_variants, stride, meta);
    if (rc) return rc;
    uint32_t b, e;
    comm_slice(n_variants, ctx->comm_rank, ctx->comm_size, b, e);
    const uint32_t rows = e - b;
    // byte range of the slice's run words (validated like twkb_load_runs; a bad descriptor fails on every rank
    // because every rank checks ALL descriptors)
    for (uint32_t v = 0; v < n_variants; ++v) {
        const twkb_run_desc& d = desc[v];
        if ((d.width != 1 && d.width != 2 && d.width != 4) || d.miss > 1 || d.offset > n_bytes ||
            (uint64_t)d.n_runs * d.width > n_bytes - d.offset) {
            ctx->err = "illegal gt primitive type / truncated runs (variant " + std::to_string(v) + ")";
            return TWKB_EINVAL;
        }
    }
    uint64_t lo = n_bytes, hi = 0;
    for (uint32_t v = b; v < e; ++v) {
        lo = std::min<uint64_t>(lo, desc[v].offset);
        hi = std::max<uint64_t>(hi, desc[v].offset + (uint64_t)desc[v].n_runs * desc[v].width);
    }
    if (hi < lo) lo = hi = 0;
    std::vector<twkb_run_desc> local(desc + b, desc + e);
    for (twkb_run_desc& d : local) d.offset -= lo;
    const uint64_t S = ((uint64_t)n_variants + ctx->comm_size - 1) / ctx->comm_size;
    const size_t words = (size_t)S * ctx->comm_size * stride;  // padded to N equal slices
    const size_t base = (size_t)S * ctx->comm_rank;            // == b unless the slice is empty
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    DevBuf<uint8_t> d_bytes;
    DevBuf<twkb_run_desc> d_desc;
    DevBuf<uint32_t> d_err, d_status;
    CUDA_TRY(d_bytes.alloc(hi - lo + 16));
    CUDA_TRY(d_desc.alloc(std::max<uint32_t>(rows, 1)));
    CUDA_TRY(d_err.alloc(2));
    CUDA_TRY(d_status.alloc(2 * (size_t)ctx->comm_size));
    CUDA_TRY(ctx->d_raw_data.alloc(words));
    if (ctx->any_missing) CUDA_TRY(ctx->d_raw_mask.alloc(words));
    const uint32_t h_err0[2] = {0u, 0xffffffffu};
    CUDA_TRY(cudaMemcpyAsync(d_err.p, h_err0, sizeof(h_err0), cudaMemcpyHostToDevice, ctx->stream));
    if (rows) {
        CUDA_TRY(cudaMemcpyAsync(d_bytes.p, bytes + lo, hi - lo, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_desc.p, local.data(), (size_t)rows * sizeof(twkb_run_desc), cudaMemcpyHostToDevice, ctx->stream));
    }
    // the whole padded slice starts from zero (the decoder ORs runs into the rows)
    CUDA_TRY(cudaMemsetAsync(ctx->d_raw_data.p + base * stride, 0, (size_t)S * stride * 8, ctx->stream));
    if (ctx->any_missing) CUDA_TRY(cudaMemsetAsync(ctx->d_raw_mask.p + base * stride, 0, (size_t)S * stride * 8, ctx->stream));
    if (rows) {
        CUDA_TRY(cudaEventRecord(ctx->ev2, ctx->stream));
        decode_runs_kernel<<<(rows + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, 0, ctx->stream>>>(
            d_bytes.p, d_desc.p, rows, (uint32_t)H, reinterpret_cast<uint32_t*>(ctx->d_raw_data.p + base * stride),
            ctx->any_missing ? reinterpret_cast<uint32_t*>(ctx->d_raw_mask.p + base * stride) : nullptr, stride * 2, d_err.p);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ctx->ev3, ctx->stream));
        ctx->stats.other_launches += 1;
    }
    // status of every rank: {error flag, slice-local variant}
    CUDA_TRY(cudaMemcpyAsync(d_status.p + 2 * ctx->comm_rank, d_err.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    {
        const NcclApi& nc = nccl_api();
        NCCL_TRY(nc.GroupStart());
        for (int k = 0; k < ctx->comm_size; ++k)
            NCCL_TRY(nc.Broadcast(d_status.p + 2 * k, d_status.p + 2 * k, 2, ncclUint32, k, ctx->comm, ctx->stream));
        NCCL_TRY(nc.GroupEnd());
    }
    rc = exchange_slices(ctx, ctx->d_raw_data.p, stride, n_variants);
    if (rc) return rc;
    if (ctx->any_missing) {
        rc = exchange_slices(ctx, ctx->d_raw_mask.p, stride, n_variants);
        if (rc) return rc;
    }
    std::vector<uint32_t> h_status(2 * (size_t)ctx->comm_size, 0);
    CUDA_TRY(cudaMemcpyAsync(h_status.data(), d_status.p, h_status.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->ms_decode = 0.0;
    if (rows) {
        float ms_dec = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms_dec, ctx->ev2, ctx->ev3));
        ctx->ms_decode = ms_dec;
    }
    ctx->stats.bytes_h2d = (hi - lo) + (size_t)rows * sizeof(twkb_run_desc);
    for (int k = 0; k < ctx->comm_size; ++k)
        if (h_status[2 * k]) {
            uint32_t kb, ke;
            comm_slice(n_variants, k, ctx->comm_size, kb, ke);
            ctx->err = "run lengths do not cover all samples (variant " + std::to_string(kb + h_status[2 * k + 1]) + ")";
            return TWKB_EINVAL;
        }
    return load_finish(ctx, meta);
}

}  // namespace twkb

using namespace twkb;

// No C++ exception may cross the C ABI (a corrupt file that drives an allocation to 2^62 bytes would
// otherwise reach the caller as std::terminate): every entry point that allocates or parses runs
// inside one of these guards and reports TWKB_ENOMEM / TWKB_EIO instead.
template <typename F>
static int guarded_buf(char* errbuf, size_t errbuf_len, F&& body) {
    auto put = [&](const char* m, int code) {
        if (errbuf && errbuf_len) std::snprintf(errbuf, errbuf_len, "%s", m);
        return code;
    };
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return put("out of host memory", TWKB_ENOMEM);
    } catch (const std::exception& e) {
        return put((std::string("internal error: ") + e.what()).c_str(), TWKB_EIO);
    } catch (...) {
        return put("internal error", TWKB_EIO);
    }
}
template <typename F>
static int guarded_ctx(void* c, F&& body) {
    Context* ctx = static_cast<Context*>(c);
    try {
        return body();
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "out of host memory";
        return TWKB_ENOMEM;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = std::string("internal error: ") + e.what();
        return TWKB_ECUDA;
    } catch (...) {
        if (ctx) ctx->err = "internal error";
        return TWKB_ECUDA;
    }
}

extern "C" {

void twkb_settings_init(twkb_settings* s) {
    if (s) settings_defaults(s);
}

int twkb_version(void) { return 100; }

const char* twkb_last_error(void* c) {
    if (!c) return g_create_error.c_str();
    return static_cast<Context*>(c)->err.c_str();
}

static int twkb_create_impl(const twkb_settings* s, void** out) {
    std::lock_guard<std::mutex> lock(g_create_mutex);
    if (!s || !out) { g_create_error = "null argument"; return TWKB_EINVAL; }
    std::string why;
    int rc = validate_settings(s, why);
    if (rc) { g_create_error = why; return rc; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
        return TWKB_ENODEVICE;
    }
    if (s->device < 0 || s->device >= n_dev) { g_create_error = "device ordinal out of range"; return TWKB_ENODEVICE; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, s->device);
    if (prop.major != 10) {
        g_create_error = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                         "; libtwkb is built for sm_100a (B200) only";
        return TWKB_ENODEVICE;
    }
    Context* ctx = new Context();
    ctx->st = *s;
    if (ctx->st.part_count <= 0) { ctx->st.part_count = 1; ctx->st.part_index = 0; }
    ctx->device = s->device;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev2) != cudaSuccess || cudaEventCreate(&ctx->ev3) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_begin) != cudaSuccess || cudaEventCreate(&ctx->ev_end) != cudaSuccess) {
        g_create_error = std::string("CUDA init failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return TWKB_ECUDA;
    }
    *out = ctx;
    return TWKB_OK;
}

void twkb_destroy(void* c) {
    if (!c) return;
    Context* ctx = static_cast<Context*>(c);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->d_raw_data.release(); ctx->d_raw_mask.release(); ctx->d_planes.release(); ctx->d_plane_popc.release();
    ctx->d_meta.release(); ctx->d_lgamma.release(); ctx->d_blk_of.release(); ctx->d_blk_first.release();
    ctx->d_blk_last.release(); ctx->d_blk_prune.release(); ctx->d_tiles.release(); ctx->d_cands.release();
    flusher_destroy(ctx);
    if (ctx->comm && nccl_api().ok) nccl_api().CommDestroy(ctx->comm);
    for (cudaEvent_t ev : ctx->chunk_events) cudaEventDestroy(ev);
    ctx->d_counters.release(); ctx->d_records[0].release(); ctx->d_records[1].release();
    ctx->d_decay_sum.release(); ctx->d_decay_cnt.release();
    ctx->d_agg_min.release(); ctx->d_agg_max.release(); ctx->d_agg_base.release(); ctx->d_agg_bins.release();
    ctx->d_collect.release();
    ctx->d_orig.release(); ctx->d_sp_off.release(); ctx->d_sp_ent.release(); ctx->d_sp_tiles.release();
    umma_release(ctx->umma);
    if (ctx->h_stage[0]) cudaFreeHost(ctx->h_stage[0]);
    if (ctx->h_stage[1]) cudaFreeHost(ctx->h_stage[1]);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_dm) cudaFreeHost(ctx->h_dm);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int twkb_update_settings(void* c, const twkb_settings* s) {
    if (!c || !s) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    std::string why;
    int rc = validate_settings(s, why);
    if (rc) { ctx->err = why; return rc; }
    if (s->device != ctx->device) { ctx->err = "cannot move a context to another device"; return TWKB_EINVAL; }
    ctx->st = *s;
    if (ctx->st.part_count <= 0) { ctx->st.part_count = 1; ctx->st.part_index = 0; }
    ctx->est_cand_per_tile = -1.0;
    return TWKB_OK;
}

int twkb_load_matrix(void* c, uint32_t n_samples, uint32_t n_variants, const uint64_t* data_bits, const uint64_t* mask_bits,
                     size_t row_stride_words, const twkb_variant* meta) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return load_common(static_cast<Context*>(c), n_samples, n_variants, data_bits, mask_bits, row_stride_words, meta, false); });
}

int twkb_load_matrix_device(void* c, uint32_t n_samples, uint32_t n_variants, const uint64_t* d_data_bits,
                            const uint64_t* d_mask_bits, size_t row_stride_words, const twkb_variant* meta) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return load_common(static_cast<Context*>(c), n_samples, n_variants, d_data_bits, d_mask_bits, row_stride_words, meta, true); });
}

int twkb_load_runs(void* c, uint32_t n_samples, uint32_t n_variants, const uint8_t* run_bytes, size_t n_run_bytes,
                   const twkb_run_desc* desc, const twkb_variant* meta) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return load_runs(static_cast<Context*>(c), n_samples, n_variants, run_bytes, n_run_bytes, desc, meta); });
}

int twkb_set_blocks(void* c, const uint32_t* block_first, uint32_t n_blocks) {
    if (!c || (!block_first && n_blocks)) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    if (!ctx->loaded) { ctx->err = "twkb_set_blocks before a load"; return TWKB_ESTATE; }
    for (uint32_t b = 0; b < n_blocks; ++b)
        if ((b == 0 && block_first[0] != 0) || (b && block_first[b] <= block_first[b - 1]) || block_first[b] >= ctx->n_variants) {
            ctx->err = "block_first must start at 0 and increase strictly below n_variants";
            return TWKB_EINVAL;
        }
    ctx->file_blocks.assign(block_first, block_first + n_blocks);
    ctx->plan_key.clear();
    ctx->sp_plan_key.clear();
    return TWKB_OK;
}

int twkb_twk_blocks(void* handle, const uint32_t** block_first, uint32_t* n_blocks) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (block_first) *block_first = f->block_first.empty() ? nullptr : f->block_first.data();
    if (n_blocks) *n_blocks = f->block_first.empty() ? 0u : (uint32_t)f->block_first.size() - 1u;
    return TWKB_OK;
}

int twkb_comm_unique_id(uint8_t* id) {
    if (!id) return TWKB_EINVAL;
    const NcclApi& nc = nccl_api();
    if (!nc.ok) { std::lock_guard<std::mutex> lock(g_create_mutex); g_create_error = nc.why; return TWKB_ENODEVICE; }
    ncclUniqueId uid;
    if (nc.GetUniqueId(&uid) != ncclSuccess) return TWKB_ECUDA;
    static_assert(sizeof(uid) == TWKB_COMM_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id, &uid, sizeof(uid));
    return TWKB_OK;
}

int twkb_comm_init(void* c, const uint8_t* id, int32_t rank, int32_t n_ranks) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    const NcclApi& nc = nccl_api();
    if (!nc.ok) { ctx->err = nc.why; return TWKB_ENODEVICE; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->comm) { nc.CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(nc.CommInitRank(&ctx->comm, n_ranks, uid, rank));
    ctx->comm_rank = rank;
    ctx->comm_size = n_ranks;
    return TWKB_OK;
}

int twkb_comm_slice(uint32_t n_variants, int32_t rank, int32_t n_ranks, uint32_t* row_begin, uint32_t* row_end) {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !row_begin || !row_end) return TWKB_EINVAL;
    comm_slice(n_variants, rank, n_ranks, *row_begin, *row_end);
    return TWKB_OK;
}

int twkb_load_matrix_sliced(void* c, uint32_t n_samples, uint32_t n_variants, const uint64_t* slice_data_bits,
                            const uint64_t* slice_mask_bits, size_t row_stride_words, const twkb_variant* meta) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return load_matrix_sliced(static_cast<Context*>(c), n_samples, n_variants, slice_data_bits, slice_mask_bits, row_stride_words, meta); });
}

int twkb_load_runs_sliced(void* c, uint32_t n_samples, uint32_t n_variants, const uint8_t* run_bytes, size_t n_run_bytes,
                          const twkb_run_desc* desc, const twkb_variant* meta) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return load_runs_sliced(static_cast<Context*>(c), n_samples, n_variants, run_bytes, n_run_bytes, desc, meta); });
}

int twkb_debug_rows(void* c, uint64_t* data_bits, uint64_t* mask_bits, size_t row_stride_words) {
    if (!c || !data_bits) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    if (!ctx->loaded) { ctx->err = "twkb_debug_rows before a load"; return TWKB_ESTATE; }
    if (row_stride_words < ctx->raw_stride) { ctx->err = "row_stride_words too small"; return TWKB_EINVAL; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t w = ctx->raw_stride * 8;
    CUDA_TRY(cudaMemcpy2D(data_bits, row_stride_words * 8, ctx->d_raw_data.p, w, w, ctx->n_variants, cudaMemcpyDeviceToHost));
    if (mask_bits) {
        if (ctx->any_missing)
            CUDA_TRY(cudaMemcpy2D(mask_bits, row_stride_words * 8, ctx->d_raw_mask.p, w, w, ctx->n_variants, cudaMemcpyDeviceToHost));
        else
            for (uint32_t v = 0; v < ctx->n_variants; ++v) std::memset(mask_bits + (size_t)v * row_stride_words, 0, w);
    }
    return TWKB_OK;
}

int twkb_compute(void* c, twkb_sink_fn sink, void* user) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return compute_impl(static_cast<Context*>(c), false, sink, user, false, nullptr); });
}

int twkb_compute_resident(void* c) {
    if (!c) return TWKB_EINVAL;
    return guarded_ctx(c, [&] { return compute_impl(static_cast<Context*>(c), true, nullptr, nullptr, false, nullptr); });
}

int twkb_compute_decay(void* c, int64_t window_bp, int32_t n_bins, double* sum_r2, uint64_t* count) {
    if (!c || !sum_r2 || !count) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    // two_reader::Decay, lib/two_reader.cpp:425-433
    if (window_bp <= 0) { ctx->err = "Window size cannot be <= 0 (provided " + std::to_string(window_bp) + ")..."; return TWKB_EINVAL; }
    if (n_bins <= 0) { ctx->err = "Number of bins cannot be <= 0 (provided " + std::to_string(n_bins) + ")..."; return TWKB_EINVAL; }
    if (window_bp / n_bins <= 0 || window_bp / n_bins > 0xffffffffll) { ctx->err = "window / bins must be a positive 32-bit bin width"; return TWKB_EINVAL; }
    return guarded_ctx(c, [&]() -> int {
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(ctx->d_decay_sum.alloc((size_t)n_bins));
        CUDA_TRY(ctx->d_decay_cnt.alloc((size_t)n_bins));
        CUDA_TRY(cudaMemsetAsync(ctx->d_decay_sum.p, 0, (size_t)n_bins * 8, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->d_decay_cnt.p, 0, (size_t)n_bins * 8, ctx->stream));
        ctx->decay_bins = (uint32_t)n_bins;
        ctx->decay_width = (uint32_t)(window_bp / n_bins);
        int rc = compute_impl(ctx, true, nullptr, nullptr, false, nullptr);
        ctx->decay_bins = 0;
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(sum_r2, ctx->d_decay_sum.p, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(count, ctx->d_decay_cnt.p, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return TWKB_OK;
    });
}

// twkb_compute_sorted: LD computation with the records kept on the device, device radix sort of both orientations
// (sort.cuh), sorted records streamed to the sink through two pinned staging buffers.
static int compute_sorted_impl(Context* ctx, twkb_sink_fn sink, void* user) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->collect = true;
    ctx->n_collected = 0;
    int rc = compute_impl(ctx, true, nullptr, nullptr, false, nullptr);
    ctx->collect = false;
    if (rc) return rc;
    const uint64_t n_fwd = ctx->n_collected, n_items = 2 * n_fwd;
    if (n_items == 0) return TWKB_OK;
    if (n_items > 0xffffffffull) { ctx->err = "too many records for the device sorter (2^31 forward records)"; return TWKB_ENOMEM; }
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t st = ctx->stream;
    const uint32_t n_tiles = (uint32_t)((n_items + SORT_TILE - 1) / SORT_TILE);
    DevBuf<unsigned long long> k_hi[2], k_lo[2], or_and;
    DevBuf<uint32_t> ref[2], counts;
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(k_hi[b].alloc(n_items));
        CUDA_TRY(k_lo[b].alloc(n_items));
        CUDA_TRY(ref[b].alloc(n_items));
    }
    CUDA_TRY(counts.alloc((size_t)256 * n_tiles));
    CUDA_TRY(or_and.alloc(4));
    const unsigned long long init[4] = {0ull, 0ull, ~0ull, ~0ull};
    CUDA_TRY(cudaMemcpyAsync(or_and.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const unsigned g_keys = (unsigned)std::min<uint64_t>((n_items + SORT_THREADS - 1) / SORT_THREADS, 148ull * 16);
    sort_keys_kernel<<<g_keys, SORT_THREADS, 0, st>>>(ctx->d_collect.p, n_items, k_hi[0].p, k_lo[0].p, ref[0].p, or_and.p);
    CUDA_TRY(cudaGetLastError());
    unsigned long long oa[4];
    CUDA_TRY(cudaMemcpyAsync(oa, or_and.p, sizeof(oa), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned long long diff_lo = oa[1] ^ oa[3], diff_hi = oa[0] ^ oa[2];  // bits that differ between some two keys
    const unsigned g_tiles = (n_tiles + SORT_WARPS - 1) / SORT_WARPS;
    int cur = 0, passes = 0;
    for (int pass = 0; pass < 16; ++pass) {
        const unsigned long long diff = pass < 8 ? diff_lo : diff_hi;
        if (((diff >> ((pass & 7) * 8)) & 0xffull) == 0) continue;  // every key has the same digit here
        radix_hist_kernel<<<g_tiles, SORT_THREADS, 0, st>>>(k_hi[cur].p, k_lo[cur].p, n_items, pass, n_tiles, counts.p);
        exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts.p, (unsigned long long)256 * n_tiles);
        radix_scatter_kernel<<<g_tiles, SORT_THREADS, 0, st>>>(k_hi[cur].p, k_lo[cur].p, ref[cur].p, n_items, pass, n_tiles, counts.p,
                                                              k_hi[cur ^ 1].p, k_lo[cur ^ 1].p, ref[cur ^ 1].p);
        CUDA_TRY(cudaGetLastError());
        cur ^= 1;
        ++passes;
    }
    ctx->stats.other_launches += 1 + 3 * (uint64_t)passes;
    // ---- gather in file order and stream out: chunk c is gathered + copied while the sink consumes chunk c - 1
    const uint64_t chunk = 1u << 17;  // records per chunk (13.9 MB)
    DevBuf<uint8_t> d_out[2];
    uint8_t* h_out[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            if (h_out[b]) cudaFreeHost(h_out[b]);
            if (ev[b]) cudaEventDestroy(ev[b]);
        }
    };
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = d_out[b].alloc((size_t)chunk * TWKB_RECORD_BYTES);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&h_out[b], (size_t)chunk * TWKB_RECORD_BYTES, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
        if (e != cudaSuccess) { cleanup(); ctx->err = std::string("device sorter staging: ") + cudaGetErrorString(e); return TWKB_ENOMEM; }
    }
    const uint64_t n_chunks = (n_items + chunk - 1) / chunk;
    for (uint64_t c = 0; c <= n_chunks && rc == TWKB_OK; ++c) {
        if (c < n_chunks) {
            const int b = (int)(c & 1);
            const uint64_t first = c * chunk, n = std::min(chunk, n_items - first);
            const unsigned long long words = n * 53ull;
            sort_gather_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(ctx->d_collect.p, ref[cur].p, first, n, d_out[b].p);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_out[b], d_out[b].p, (size_t)n * TWKB_RECORD_BYTES, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
            if (e != cudaSuccess) { cleanup(); return ctx->fail(std::string("device sorter: ") + cudaGetErrorString(e)); }
            ctx->stats.other_launches += 1;
            ctx->stats.bytes_d2h += n * TWKB_RECORD_BYTES;
        }
        if (c > 0) {
            const int b = (int)((c - 1) & 1);
            const uint64_t first = (c - 1) * chunk, n = std::min(chunk, n_items - first);
            cudaError_t e = cudaEventSynchronize(ev[b]);
            if (e != cudaSuccess) { cleanup(); return ctx->fail(std::string("device sorter: ") + cudaGetErrorString(e)); }
            if (sink && sink(user, h_out[b], n) != 0) { rc = TWKB_ESINK; ctx->err = "the record sink returned non-zero"; }
        }
    }
    cudaStreamSynchronize(st);
    cleanup();
    ctx->stats.seconds_total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int twkb_compute_sorted(void* c, twkb_sink_fn sink, void* user) {
    if (!c || !sink) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    if (ctx->st.part_count > 1) { ctx->err = "sorted output needs every record on one device (part_count = 1)"; return TWKB_EINVAL; }
    return guarded_ctx(c, [&] { return compute_sorted_impl(ctx, sink, user); });
}

int twkb_compute_aggregate(void* c, int32_t field, int32_t xbins, int32_t ybins, const int64_t* contig_n_bases, uint32_t n_contigs,
                           twkb_agg_bin* bins, twkb_agg_layout* layout, uint64_t* contig_offset, uint32_t* contig_min, uint32_t* contig_max) {
    if (!c || !bins || !contig_n_bases || n_contigs == 0 || n_contigs > (1u << 24)) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    static_assert(sizeof(twkb_agg_bin) == sizeof(AggBin), "twk_sstats layout");
    // two_reader::Aggregate, lib/two_reader.cpp:620-628
    if (field < TWKB_AGG_R2 || field > TWKB_AGG_ALTS) { ctx->err = "Unknown aggregation function..."; return TWKB_EINVAL; }
    if (xbins < 5) { ctx->err = "Number of x-bins cannot be < 5!"; return TWKB_EINVAL; }
    if (ybins < 5) { ctx->err = "Number of y-bins cannot be < 5!"; return TWKB_EINVAL; }
    if ((uint64_t)xbins * (uint64_t)ybins > (1ull << 28)) { ctx->err = "raster too large"; return TWKB_EINVAL; }
    return guarded_ctx(c, [&]() -> int {
        CUDA_TRY(cudaSetDevice(ctx->device));
        const size_t nb = (size_t)xbins * ybins;
        // ---- pass 1: which contigs occur, and their position ranges (FindRangesUnsorted, aggregation.h:127-148)
        CUDA_TRY(ctx->d_agg_min.alloc(n_contigs));
        CUDA_TRY(ctx->d_agg_max.alloc(n_contigs));
        CUDA_TRY(ctx->d_agg_base.alloc((size_t)n_contigs + 1));
        CUDA_TRY(cudaMemsetAsync(ctx->d_agg_min.p, 0xff, (size_t)n_contigs * 4, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->d_agg_max.p, 0, (size_t)n_contigs * 4, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->d_agg_base.p, 0, ((size_t)n_contigs + 1) * 8, ctx->stream));
        ctx->agg_layout = AggLayout{};
        ctx->agg_layout.n_contigs = n_contigs;
        ctx->agg_pass = 1;
        int rc = compute_impl(ctx, true, nullptr, nullptr, false, nullptr);
        ctx->agg_pass = 0;
        if (rc) return rc;
        std::vector<uint32_t> cmin(n_contigs), cmax(n_contigs);
        unsigned long long bad = 0;
        CUDA_TRY(cudaMemcpyAsync(cmin.data(), ctx->d_agg_min.p, (size_t)n_contigs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(cmax.data(), ctx->d_agg_max.p, (size_t)n_contigs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&bad, ctx->d_agg_base.p + n_contigs, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (bad) { ctx->err = "a record names a contig beyond n_contigs"; return TWKB_EINVAL; }
        const uint64_t n_fwd = ctx->stats.records_out;
        // ---- coordinate system (two_reader.cpp:734-797)
        std::vector<uint8_t> set(n_contigs);
        uint32_t n_set = 0, n_set_ref = 0;
        for (uint32_t i = 0; i < n_contigs; ++i) { set[i] = cmin[i] != 0xffffffffu; n_set += set[i]; }
        n_set_ref = n_set + (ctx->st.emulate_quirks && set[0] ? 1u : 0u);  // :737-740 starts the sum AT contig_avail[0].set and adds it again
        std::vector<uint64_t> cum(n_contigs, 0);
        std::vector<uint32_t> omin(n_contigs), omax(n_contigs);
        uint64_t range = 0;
        for (uint32_t i = 0; i < n_contigs; ++i) {
            uint64_t span;
            if (n_set_ref == 1) { omin[i] = cmin[i]; omax[i] = cmax[i]; span = set[i] ? (uint64_t)(cmax[i] - cmin[i]) + 1 : 0; }
            else { omin[i] = 0; omax[i] = (uint32_t)contig_n_bases[i]; span = set[i] ? (uint64_t)contig_n_bases[i] : 0; }
            range += span;
            cum[i] = range;
        }
        std::memset(bins, 0, nb * sizeof(twkb_agg_bin));
        if (layout) *layout = twkb_agg_layout{range, 0, 0, n_set, 2 * n_fwd};
        for (uint32_t i = 0; i < n_contigs; ++i) {
            if (contig_offset) contig_offset[i] = cum[i];
            if (contig_min) contig_min[i] = omin[i];
            if (contig_max) contig_max[i] = omax[i];
        }
        if (n_fwd == 0 || range == 0) return TWKB_OK;  // "Cannot aggregate empty file...": an all-zero raster
        const uint32_t xrange = (uint32_t)std::ceil((float)range / xbins), yrange = (uint32_t)std::ceil((float)range / ybins);  // :801-802
        if (layout) { layout->bpx = xrange; layout->bpy = yrange; }
        // ---- pass 2: the raster (BuildMatrix, aggregation.h:150-172)
        std::vector<unsigned long long> base(n_contigs);
        for (uint32_t i = 0; i < n_contigs; ++i) base[i] = cum[i] - (uint64_t)(uint32_t)(omax[i] - omin[i]);
        CUDA_TRY(cudaMemcpyAsync(ctx->d_agg_base.p, base.data(), (size_t)n_contigs * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_agg_min.p, omin.data(), (size_t)n_contigs * 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx->d_agg_bins.alloc(nb));
        CUDA_TRY(cudaMemsetAsync(ctx->d_agg_bins.p, 0, nb * sizeof(AggBin), ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // base / omin are stack vectors
        ctx->agg_layout = AggLayout{ctx->d_agg_base.p, ctx->d_agg_min.p, n_contigs, xrange, yrange, (uint32_t)xbins, (uint32_t)ybins, field};
        ctx->agg_pass = 2;
        rc = compute_impl(ctx, true, nullptr, nullptr, false, nullptr);
        ctx->agg_pass = 0;
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(bins, ctx->d_agg_bins.p, nb * sizeof(AggBin), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return TWKB_OK;
    });
}

int twkb_twk_contigs(void* handle, int64_t* n_bases, uint32_t capacity, uint32_t* n_contigs) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (n_contigs) *n_contigs = (uint32_t)f->contig_n_bases.size();
    if (n_bases) {
        if (capacity < f->contig_n_bases.size()) return TWKB_ENOMEM;
        std::copy(f->contig_n_bases.begin(), f->contig_n_bases.end(), n_bases);
    }
    return TWKB_OK;
}

int twkb_get_stats(void* c, twkb_stats* out) {
    if (!c || !out) return TWKB_EINVAL;
    *out = static_cast<Context*>(c)->stats;
    return TWKB_OK;
}

static int twkb_debug_candidates_impl(void* c, int screen_off, uint32_t* out, uint64_t capacity, uint64_t* n_out) {
    if (!c || !n_out) return TWKB_EINVAL;
    Context* ctx = static_cast<Context*>(c);
    std::vector<Candidate> dump;
    int rc = compute_impl(ctx, true, nullptr, nullptr, screen_off != 0, &dump);
    if (rc) return rc;
    if (ctx->permuted)  // candidates carry resident indices; callers see the file order
        for (Candidate& cd : dump) { cd.i = ctx->h_orig[cd.i]; cd.j = ctx->h_orig[cd.j]; }
    *n_out = dump.size();
    if (out) {
        if (dump.size() > capacity) { ctx->err = "debug buffer too small"; return TWKB_ENOMEM; }
        std::memcpy(out, dump.data(), dump.size() * sizeof(Candidate));
    }
    return TWKB_OK;
}

static int writer_sink(void* user, const uint8_t* recs, uint64_t n) {
    return static_cast<TwoWriter*>(user)->add(recs, n);
}

int twkb_calc_file(const twkb_settings* s, const char* in_path, const char* out_path, twkb_stats* stats_out, char* errbuf,
                   size_t errbuf_len) {
    return twkb_calc_file_intervals(s, in_path, out_path, nullptr, 0, stats_out, errbuf, errbuf_len);
}

// `tomahawk calc`: twk_ld::Compute (reference lib/ld/ld.cpp:477-671) end to end.
static int twkb_calc_file_intervals_impl(const twkb_settings* s, const char* in_path, const char* out_path, const char* const* intervals,
                             int32_t n_intervals, twkb_stats* stats_out, char* errbuf, size_t errbuf_len) {
    auto fail = [&](int code, const std::string& m) {
        if (errbuf && errbuf_len) {
            std::snprintf(errbuf, errbuf_len, "%s", m.c_str());
        }
        return code;
    };
    if (!s || !in_path || !out_path) return fail(TWKB_EINVAL, "null argument");
    if (std::strlen(in_path) == 0) return fail(TWKB_EINVAL, "No file-name provided...");  // ld.cpp:480
    std::string err;
    TwkFile twk;
    std::vector<std::string> ivals;
    for (int32_t i = 0; i < n_intervals; ++i) {
        if (!intervals || !intervals[i]) return fail(TWKB_EINVAL, "null interval string");
        ivals.emplace_back(intervals[i]);
    }
    const bool runs = s->host_unpack == 0;  // default: the device decodes the run-length records
    const auto t_open = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    twkb_settings settings = *s;
    if (s->single) {  // scalc: twk_ld::ComputeSingle, lib/ld/ld.cpp:673-697
        if (s->n_chunks != 1) return fail(TWKB_EINVAL, "Cannot use chunking in single mode!");
        if (s->window) return fail(TWKB_EINVAL, "Cannot use window in single mode!");
        if (ivals.empty()) return fail(TWKB_EINVAL, "An interval has to be provided in single mode!");
        if (ivals.size() != 1) return fail(TWKB_EINVAL, "Only a single interval can be provided in single mode!");
    }
    int rc = read_twk(in_path, std::max(1, s->n_threads), twk, err, ivals.empty() ? nullptr : &ivals, s->emulate_quirks != 0, runs,
                      s->single ? std::max(0, s->l_surrounding) : -1);
    if (rc) return fail(rc, err);
    if (s->single) settings.single_targets = (int32_t)twk.n_targets;
    s = &settings;
    const double sec_read = since(t_open);
    void* c = nullptr;
    rc = twkb_create(s, &c);
    if (rc) return fail(rc, twkb_last_error(nullptr));
    Context* ctx = static_cast<Context*>(c);
    const auto t_load = std::chrono::steady_clock::now();
    if (runs)
        rc = twkb_load_runs(c, twk.n_samples, twk.n_variants, twk.raw.data(), twk.raw.size(), twk.run_desc.data(), twk.meta.data());
    else
        rc = twkb_load_matrix(c, twk.n_samples, twk.n_variants, twk.data.data(), twk.any_missing ? twk.mask.data() : nullptr,
                              twk.stride, twk.meta.data());
    if (rc == TWKB_OK && twk.block_first.size() > 1)  // the file's own block structure (window rules, -c chunks)
        rc = twkb_set_blocks(c, twk.block_first.data(), (uint32_t)twk.block_first.size() - 1);
    if (rc == TWKB_OK && runs) twk.raw.release();  // the runs now live on the device
    const double sec_load = since(t_load);
    if (rc) { err = ctx->err; twkb_destroy(c); return fail(rc, err); }
    // output name: the reference forces a ".two" suffix (ld.cpp:589-598)
    std::string out = out_path;
    {
        const size_t slash = out.find_last_of('/');
        const size_t dot = out.find_last_of('.');
        const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
        std::string ext = has_ext ? out.substr(dot + 1) : "";
        for (auto& ch : ext) ch = (char)std::tolower(ch);
        if (ext != "two") out = (has_ext ? out.substr(0, dot) : out) + ".two";
        if (std::strcmp(out_path, "-") == 0) out = "-";  // stream to stdout (the reference's default, ld.cpp:585-588)
    }
    std::string cmd = std::string("tomahawk_b200 calc -i ") + in_path + " -o " + out_path;
    if (s->sorted_output) {  // records ordered on the device, written as a sorted .two (what calc + sort produce in the reference)
        if (out == "-") { twkb_destroy(c); return fail(TWKB_EINVAL, "sorted output needs a file (-o), not stdout"); }
        SortedTwoWriter sw;
        rc = sw.open(out, twk, cmd, s->c_level, std::max(1, s->n_threads), err);
        if (rc) { twkb_destroy(c); return fail(rc, err); }
        auto ssink = [](void* user, const uint8_t* recs, uint64_t n) -> int { return static_cast<SortedTwoWriter*>(user)->add(recs, n); };
        rc = twkb_compute_sorted(c, ssink, &sw);
        if (rc) { err = ctx->err.empty() || rc == TWKB_ESINK ? sw.error() : ctx->err; twkb_destroy(c); return fail(rc, err); }
        rc = sw.finish();
        if (rc) { err = sw.error(); twkb_destroy(c); return fail(rc, err); }
        if (stats_out) {
            *stats_out = ctx->stats;
            stats_out->seconds_file_read = sec_read;
            stats_out->seconds_file_load = sec_load;
            stats_out->seconds_file_total = since(t_open);
        }
        twkb_destroy(c);
        return TWKB_OK;
    }
    TwoWriter writer;
    rc = writer.open(out, twk, cmd, s->c_level, s->b_size, err);
    if (rc) { twkb_destroy(c); return fail(rc, err); }
    writer.set_threads(std::max(1, s->n_threads));
    rc = twkb_compute(c, writer_sink, &writer);
    if (rc) { err = ctx->err.empty() ? writer.error() : ctx->err; twkb_destroy(c); return fail(rc, err); }
    rc = writer.finish();
    if (rc) { err = writer.error(); twkb_destroy(c); return fail(rc, err); }
    if (stats_out) {
        *stats_out = ctx->stats;
        stats_out->seconds_file_read = sec_read;
        stats_out->seconds_file_load = sec_load;
        stats_out->seconds_file_total = since(t_open);
    }
    twkb_destroy(c);
    return TWKB_OK;
}

// ------------------------------------------------------------------ host-only hooks
static int copy_err(char* errbuf, size_t n, const std::string& m, int code) {
    if (errbuf && n) std::snprintf(errbuf, n, "%s", m.c_str());
    return code;
}

int twkb_twk_open(const char* path, int n_threads, void** handle, char* errbuf, size_t errbuf_len) {
    return twkb_twk_open_intervals(path, n_threads, nullptr, 0, 1, handle, errbuf, errbuf_len);
}

static int twkb_twk_open_intervals_impl(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                            int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len) {
    if (!path || !handle) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
    std::vector<std::string> ivals;
    for (int32_t i = 0; i < n_intervals; ++i) {
        if (!intervals || !intervals[i]) return copy_err(errbuf, errbuf_len, "null interval string", TWKB_EINVAL);
        ivals.emplace_back(intervals[i]);
    }
    TwkFile* f = new TwkFile();
    std::string err;
    const int rc = read_twk(path, std::max(1, n_threads), *f, err, ivals.empty() ? nullptr : &ivals, emulate_quirks != 0);
    if (rc) { delete f; return copy_err(errbuf, errbuf_len, err, rc); }
    *handle = f;
    return TWKB_OK;
}

static int twkb_twk_open_runs_impl(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                       int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len) {
    if (!path || !handle) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
    std::vector<std::string> ivals;
    for (int32_t i = 0; i < n_intervals; ++i) {
        if (!intervals || !intervals[i]) return copy_err(errbuf, errbuf_len, "null interval string", TWKB_EINVAL);
        ivals.emplace_back(intervals[i]);
    }
    TwkFile* f = new TwkFile();
    std::string err;
    const int rc = read_twk(path, std::max(1, n_threads), *f, err, ivals.empty() ? nullptr : &ivals, emulate_quirks != 0, true);
    if (rc) { delete f; return copy_err(errbuf, errbuf_len, err, rc); }
    *handle = f;
    return TWKB_OK;
}

int twkb_twk_open_single(const char* path, int n_threads, const char* interval, int32_t l_surrounding, int32_t emulate_quirks,
                         int32_t runs_mode, void** handle, uint32_t* n_targets, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&]() -> int {
        if (!path || !handle || !interval) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
        std::vector<std::string> ivals{std::string(interval)};
        TwkFile* f = new TwkFile();
        std::string err;
        const int rc = read_twk(path, std::max(1, n_threads), *f, err, &ivals, emulate_quirks != 0, runs_mode != 0, std::max(0, l_surrounding));
        if (rc) { delete f; return copy_err(errbuf, errbuf_len, err, rc); }
        if (n_targets) *n_targets = f->n_targets;
        *handle = f;
        return TWKB_OK;
    });
}

int twkb_twk_runs_view(void* handle, const uint8_t** run_bytes, size_t* n_run_bytes, const twkb_run_desc** desc,
                       const twkb_variant** meta) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (!f->runs_mode) return TWKB_ESTATE;
    if (run_bytes) *run_bytes = f->raw.data();
    if (n_run_bytes) *n_run_bytes = f->raw.size();
    if (desc) *desc = f->run_desc.data();
    if (meta) *meta = f->meta.data();
    return TWKB_OK;
}

int twkb_twk_dims(void* handle, uint32_t* n_samples, uint32_t* n_variants, size_t* row_stride_words, int32_t* any_missing,
                  uint32_t* n_blocks) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (n_samples) *n_samples = f->n_samples;
    if (n_variants) *n_variants = f->n_variants;
    if (row_stride_words) *row_stride_words = f->stride;
    if (any_missing) *any_missing = f->any_missing ? 1 : 0;
    if (n_blocks) *n_blocks = f->n_blocks;
    return TWKB_OK;
}

int twkb_twk_copy(void* handle, uint64_t* data_bits, uint64_t* mask_bits, twkb_variant* meta) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (f->runs_mode) return TWKB_ESTATE;
    if (data_bits) std::memcpy(data_bits, f->data.data(), f->data.size() * 8);
    if (mask_bits) {
        if (f->any_missing) std::memcpy(mask_bits, f->mask.data(), f->mask.size() * 8);
        else std::memset(mask_bits, 0, f->data.size() * 8);
    }
    if (meta) std::memcpy(meta, f->meta.data(), f->meta.size() * sizeof(twkb_variant));
    return TWKB_OK;
}

int twkb_twk_view(void* handle, const uint64_t** data_bits, const uint64_t** mask_bits, const twkb_variant** meta) {
    if (!handle) return TWKB_EINVAL;
    const TwkFile* f = static_cast<TwkFile*>(handle);
    if (f->runs_mode) return TWKB_ESTATE;
    if (data_bits) *data_bits = f->data.data();
    if (mask_bits) *mask_bits = f->any_missing ? f->mask.data() : nullptr;
    if (meta) *meta = f->meta.data();
    return TWKB_OK;
}

void twkb_twk_close(void* handle) { delete static_cast<TwkFile*>(handle); }

static int twkb_two_open_impl(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t b_size,
                  void** writer, char* errbuf, size_t errbuf_len) {
    if (!path || !twk_handle || !writer) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
    TwoWriter* w = new TwoWriter();
    std::string err;
    const int rc = w->open(path, *static_cast<TwkFile*>(twk_handle), command_line ? command_line : "", c_level, b_size, err);
    if (rc) { delete w; return copy_err(errbuf, errbuf_len, err, rc); }
    *writer = w;
    return TWKB_OK;
}

int twkb_two_set_threads(void* writer, int32_t n_threads) {
    if (!writer) return TWKB_EINVAL;
    static_cast<TwoWriter*>(writer)->set_threads(n_threads);
    return TWKB_OK;
}

static int twkb_two_add_impl(void* writer, const uint8_t* records, uint64_t n) {
    if (!writer || (!records && n)) return TWKB_EINVAL;
    return static_cast<TwoWriter*>(writer)->add(records, n);
}

static int twkb_two_close_impl(void* writer) {
    if (!writer) return TWKB_EINVAL;
    TwoWriter* w = static_cast<TwoWriter*>(writer);
    const int rc = w->finish();
    delete w;
    return rc;
}

static int twkb_two_open_sorted_impl(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t n_threads,
                                     void** writer, char* errbuf, size_t errbuf_len) {
    if (!path || !twk_handle || !writer) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
    SortedTwoWriter* w = new SortedTwoWriter();
    std::string err;
    const int rc = w->open(path, *static_cast<TwkFile*>(twk_handle), command_line ? command_line : "", c_level, n_threads, err);
    if (rc) { delete w; return copy_err(errbuf, errbuf_len, err, rc); }
    *writer = w;
    return TWKB_OK;
}

static int twkb_two_sort_impl(const char* in_path, const char* out_path, int32_t c_level, int32_t n_threads, uint64_t* n_records,
                  char* errbuf, size_t errbuf_len) {
    return twkb_two_sort_mem(in_path, out_path, c_level, n_threads, 0, n_records, errbuf, errbuf_len);
}

int twkb_two_sort_mem(const char* in_path, const char* out_path, int32_t c_level, int32_t n_threads, uint64_t memory_limit_bytes,
                      uint64_t* n_records, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&]() -> int {
        if (!in_path || !out_path) return copy_err(errbuf, errbuf_len, "null argument", TWKB_EINVAL);
        if (std::strlen(in_path) == 0) return copy_err(errbuf, errbuf_len, "No input value specified...", TWKB_EINVAL);  // two_reader.cpp:169
        std::string err;
        const int rc = sort_two(in_path, out_path, c_level, n_threads, err, n_records, memory_limit_bytes);
        if (rc) return copy_err(errbuf, errbuf_len, err, rc);
        return TWKB_OK;
    });
}

static int twkb_plan_tiles_impl(const twkb_settings* s, uint32_t n_variants, const twkb_variant* meta, uint32_t tile_i, uint32_t tile_j,
                    uint32_t* out_ij, uint64_t capacity, uint64_t* n_tiles, uint64_t* n_pairs) {
    if (!s || !meta || !n_tiles || n_variants == 0 || tile_i == 0 || tile_j == 0) return TWKB_EINVAL;
    std::string why;
    int rc = validate_settings(s, why);
    if (rc) return rc;
    Context ctx;  // host-only use: no CUDA call is made on this path
    ctx.st = *s;
    if (ctx.st.part_count <= 0) { ctx.st.part_count = 1; ctx.st.part_index = 0; }
    ctx.n_variants = n_variants;
    copy_meta(&ctx, meta);
    Problem pb;
    rc = select_problem(&ctx, pb);
    if (rc) return rc;
    ctx.h_orig.resize(n_variants);
    for (uint32_t x = 0; x < n_variants; ++x) ctx.h_orig[x] = x;
    if (ctx.st.window) {
        std::vector<uint32_t> blk_of_orig;
        build_blocks_host(&ctx, blk_of_orig);
    }
    std::vector<uint2> tiles;
    uint64_t pairs = 0;
    build_tiles(&ctx, pb, tile_i, tile_j, 32, tiles, &pairs);
    *n_tiles = tiles.size();
    if (n_pairs) *n_pairs = pairs;
    if (out_ij) {
        if (tiles.size() > capacity) return TWKB_ENOMEM;
        for (size_t t = 0; t < tiles.size(); ++t) { out_ij[2 * t] = tiles[t].x; out_ij[2 * t + 1] = tiles[t].y; }
    }
    return TWKB_OK;
}

// Position shards of a -w run (twkb.h). The reach of a block row is the balancer's row prune (ld_balancing.h:189-196,
// the same loop as build_blocks_host): block bj is reachable from bi until the first bj whose first position is more
// than the window past bi's last position -- positions only, uint32 wrap-around included, no contig test.
static int twkb_plan_shards_impl(const uint32_t* block_first, uint32_t n_blocks, const twkb_variant* meta, uint32_t n_variants,
                                 int32_t l_window, int32_t n_shards, uint32_t* own_begin, uint32_t* halo_end) {
    if (!block_first || !meta || !own_begin || !halo_end || n_blocks == 0 || n_shards < 1 || l_window < 0) return TWKB_EINVAL;
    if (block_first[0] != 0 || block_first[n_blocks] != n_variants) return TWKB_EINVAL;
    for (uint32_t b = 0; b < n_blocks; ++b)
        if (block_first[b + 1] <= block_first[b]) return TWKB_EINVAL;
    const uint32_t w = (uint32_t)l_window;
    std::vector<uint32_t> prune(n_blocks, n_blocks);
    std::vector<double> work(n_blocks + 1, 0.0);  // prefix sums of the pairs a block row visits
    for (uint32_t bi = 0; bi < n_blocks; ++bi) {
        const uint32_t last_pos = meta[block_first[bi + 1] - 1].pos;
        for (uint32_t bj = bi + 1; bj < n_blocks; ++bj)
            if ((uint32_t)(meta[block_first[bj]].pos - last_pos) > w) { prune[bi] = bj; break; }
        const double ni = (double)(block_first[bi + 1] - block_first[bi]);
        const double reach = (double)(block_first[prune[bi]] - block_first[bi + 1]);
        work[bi + 1] = work[bi] + ni * (ni - 1.0) / 2.0 + ni * reach;
    }
    const uint32_t shards = std::min<uint32_t>((uint32_t)n_shards, n_blocks);
    own_begin[0] = 0;
    uint32_t b = 0;
    for (uint32_t k = 1; k < (uint32_t)n_shards; ++k) {
        if (k >= shards) { own_begin[k] = n_blocks; continue; }  // more shards than blocks: the extra ones are empty
        const double target = work[n_blocks] * (double)k / (double)shards;
        while (b < n_blocks && work[b] < target) ++b;
        b = std::max(b, own_begin[k - 1] + 1);                    // every shard owns at least one block ...
        b = std::min(b, n_blocks - (shards - k));                 // ... and leaves one for each later shard
        own_begin[k] = b;
    }
    own_begin[n_shards] = n_blocks;
    for (uint32_t k = 0; k < (uint32_t)n_shards; ++k) {
        uint32_t e = own_begin[k + 1];
        for (uint32_t bi = own_begin[k]; bi < own_begin[k + 1]; ++bi) e = std::max(e, prune[bi]);
        halo_end[k] = e;
    }
    return TWKB_OK;
}

// ---- exception-safe entry points (see guarded_buf / guarded_ctx)
int twkb_calc_file_intervals(const twkb_settings* s, const char* in_path, const char* out_path, const char* const* intervals,
                             int32_t n_intervals, twkb_stats* stats_out, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_calc_file_intervals_impl(s, in_path, out_path, intervals, n_intervals, stats_out, errbuf, errbuf_len); });
}

int twkb_twk_open_intervals(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                            int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_twk_open_intervals_impl(path, n_threads, intervals, n_intervals, emulate_quirks, handle, errbuf, errbuf_len); });
}

int twkb_twk_open_runs(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                       int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_twk_open_runs_impl(path, n_threads, intervals, n_intervals, emulate_quirks, handle, errbuf, errbuf_len); });
}

int twkb_two_open(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t b_size,
                  void** writer, char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_two_open_impl(path, twk_handle, command_line, c_level, b_size, writer, errbuf, errbuf_len); });
}

int twkb_two_sort(const char* in_path, const char* out_path, int32_t c_level, int32_t n_threads, uint64_t* n_records,
                  char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_two_sort_impl(in_path, out_path, c_level, n_threads, n_records, errbuf, errbuf_len); });
}

int twkb_two_add(void* writer, const uint8_t* records, uint64_t n) {
    return guarded_buf(nullptr, 0, [&] { return twkb_two_add_impl(writer, records, n); });
}

int twkb_two_close(void* writer) {
    return guarded_buf(nullptr, 0, [&] { return twkb_two_close_impl(writer); });
}

int twkb_two_open_sorted(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t n_threads, void** writer,
                         char* errbuf, size_t errbuf_len) {
    return guarded_buf(errbuf, errbuf_len, [&] { return twkb_two_open_sorted_impl(path, twk_handle, command_line, c_level, n_threads, writer, errbuf, errbuf_len); });
}

int twkb_two_add_sorted(void* writer, const uint8_t* records, uint64_t n) {
    if (!writer || (!records && n)) return TWKB_EINVAL;
    return guarded_buf(nullptr, 0, [&] { return static_cast<SortedTwoWriter*>(writer)->add(records, n); });
}

int twkb_two_close_sorted(void* writer) {
    if (!writer) return TWKB_EINVAL;
    return guarded_buf(nullptr, 0, [&] {
        SortedTwoWriter* w = static_cast<SortedTwoWriter*>(writer);
        const int rc = w->finish();
        delete w;
        return rc;
    });
}

int twkb_plan_tiles(const twkb_settings* s, uint32_t n_variants, const twkb_variant* meta, uint32_t tile_i, uint32_t tile_j,
                    uint32_t* out_ij, uint64_t capacity, uint64_t* n_tiles, uint64_t* n_pairs) {
    return guarded_buf(nullptr, 0, [&] { return twkb_plan_tiles_impl(s, n_variants, meta, tile_i, tile_j, out_ij, capacity, n_tiles, n_pairs); });
}

int twkb_plan_shards(const uint32_t* block_first, uint32_t n_blocks, const twkb_variant* meta, uint32_t n_variants, int32_t l_window,
                     int32_t n_shards, uint32_t* own_begin, uint32_t* halo_end) {
    return guarded_buf(nullptr, 0, [&] { return twkb_plan_shards_impl(block_first, n_blocks, meta, n_variants, l_window, n_shards, own_begin, halo_end); });
}

int twkb_create(const twkb_settings* s, void** out) {
    return guarded_buf(nullptr, 0, [&] { return twkb_create_impl(s, out); });
}

int twkb_debug_candidates(void* c, int screen_off, uint32_t* out, uint64_t capacity, uint64_t* n_out) {
    return guarded_ctx(c, [&] { return twkb_debug_candidates_impl(c, screen_off, out, capacity, n_out); });
}

}  // extern "C"
