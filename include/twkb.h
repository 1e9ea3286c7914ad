/* twkb.h -- C-ABI of the B200-native `tomahawk calc` engine (libtwkb.so).
 *
 * This is the drop-in boundary for ONE path of mklarqvist/tomahawk: pairwise LD
 * (`tomahawk calc`, lib/ld in the reference). The reference has no FFI/plugin
 * registry; the seam it does have is the library API
 *     bool twk_ld::Compute(const twk_ld_settings&)            include/ld.h:53
 * called by the CLI (lib/calc.h:237-238) and by the external R/Python bindings.
 * Everything below is what a binding for that seam needs, in plain C types:
 *
 *   reference                                         | here
 *   --------------------------------------------------+--------------------------
 *   twk_ld_settings           include/core.h:909-924  | twkb_settings
 *   twk1_t (pos/ac/an/rid/hwe) include/core.h:291-295 | twkb_variant
 *   twk_igt_vec data/mask     include/core.h:724-753  | data_bits / mask_bits rows
 *   twk_ld::Compute           lib/ld/ld.cpp:477-671   | twkb_create + twkb_load_matrix
 *                                                     |   + twkb_compute
 *   twk1_two_t serializer     lib/core.cpp:470-490    | 106-byte records handed to
 *                                                     |   twkb_sink_fn
 *   twk_ld_progress           lib/ld/ld_progress.h    | twkb_get_stats
 *
 * All functions return 0 on success or a negative TWKB_E* code; the message for
 * the last failure on a context is available from twkb_last_error(). Nothing
 * here ever falls back to a CPU implementation: without a CUDA device (or with
 * a device other than sm_100) twkb_create fails with TWKB_ENODEVICE.
 */
#ifndef TWKB_H_
#define TWKB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWKB_OK 0
#define TWKB_EINVAL (-1)    /* bad argument / inconsistent settings          */
#define TWKB_ENODEVICE (-2) /* no usable sm_100 CUDA device                  */
#define TWKB_ECUDA (-3)     /* CUDA runtime error (see twkb_last_error)      */
#define TWKB_ENOMEM (-4)    /* host or device allocation failed              */
#define TWKB_ESTATE (-5)    /* call order violated (e.g. compute before load)*/
#define TWKB_ESINK (-6)     /* the record sink returned non-zero             */
#define TWKB_EIO (-7)       /* file I/O (.twk reader / .two writer)          */

#define TWKB_RECORD_BYTES 106 /* packed twk1_two_t, include/core.h:758-759 */

/* Count kernel selection. AUTO picks the tensor-core (tcgen05 int8) kernel for
 * phased data without missing genotypes and the LOP3+POPC kernel otherwise. */
#define TWKB_KERNEL_AUTO 0
#define TWKB_KERNEL_POPC 1
#define TWKB_KERNEL_UMMA 2     /* tcgen05 kind::i8, int8 0/1 operands, int32 accumulate        */
#define TWKB_KERNEL_UMMA_FP4 3 /* tcgen05 kind::mxf4, e2m1 0/1 operands, exact fp32 accumulate */

/* 1:1 with the fields of twk_ld_settings that `calc` reads (include/core.h:909-924;
 * defaults lib/core.cpp:297-306 -- see twkb_settings_init), plus device placement.
 * Flags accepted for CLI compatibility but with no effect on results are kept so a
 * caller can pass its twk_ld_settings through unchanged. */
typedef struct twkb_settings {
    uint8_t square;          /* parsed by the reference, never read (lib/calc.h) */
    uint8_t window;          /* -w given                                          */
    uint8_t low_memory;      /* -m: CPU RAM trick; accepted, no-op                */
    uint8_t bitmaps;         /* -M: EWAH bitmaps; accepted, same kernels          */
    uint8_t single;          /* scalc / twk_ld::ComputeSingle: target site(s) against their
                                neighbourhood (see single_targets)                 */
    uint8_t force_phased;    /* -p                                                */
    uint8_t forced_unphased; /* -u                                                */
    uint8_t emulate_quirks;  /* 1 (default): reproduce count-slot quirk Q3 of the
                                reference for low-AC pairs with missing data      */
    int32_t c_level;         /* zstd level of the .two writer (-c)                */
    int32_t bl_size;         /* parsed by the reference, never read               */
    int32_t b_size;          /* records per output block (10000)                  */
    int32_t l_window;        /* -w window in bp                                   */
    int32_t n_threads;       /* host writer threads (-t)                          */
    int32_t l_surrounding;   /* scalc only                                        */
    int32_t n_chunks;        /* -c: number of sub-problems, k(k+1)/2              */
    int32_t c_chunk;         /* -C: chosen sub-problem, 0-based                   */
    double minP, minR2, maxR2, minDprime, maxDprime; /* -P -r -R -d -D           */
    /* --- B200 additions --- */
    int32_t device;          /* CUDA device ordinal of this context               */
    int32_t part_index;      /* multi-GPU: this context's share of the tile grid  */
    int32_t part_count;      /*   (rank, world size); 0/1 = everything            */
    int32_t kernel;          /* TWKB_KERNEL_*                                     */
    int32_t twk_block_size;  /* .twk block length that defines window-mode tiles
                                (500, lib/importer.h:36)                          */
    int32_t sparse_max_words; /* rare-variant (list) path: a variant whose haplotype row has at
                                most this many non-zero 32-bit words is kept as a word list
                                and served by the sparse kernel (reference: twk_igt_list +
                                PhasedListVector). 0 = automatic (ceil(2N/32)/64 when
                                2N >= 32768, phased data without missing genotypes, no -c
                                chunking), > 0 explicit, < 0 never                       */
    int32_t host_unpack;     /* twkb_calc_file: 0 (default) = the .twk run-length records are decoded
                                on the device (twkb_load_runs); 1 = unpack the rows on the host
                                (twkb_load_matrix), the reference's twk_igt_vec::Build arrangement */
    int32_t single_targets;  /* single mode: the FIRST single_targets resident variants are the target site(s); every
                                pair (target, other variant) and (target, later target) is computed, the target named
                                first (twk_ld_slave::CalculateSingle, lib/ld/ld_engine.cpp:2226-2332: auto phasing per
                                pair, no ac_i + ac_j <= 2 skip). twkb_calc_file* fill it from the file; callers of
                                twkb_load_matrix order their rows [targets | neighbours] and set it themselves */
    int32_t shard_blocks;    /* position-sharded window runs (matrix larger than one GPU): this context holds ONE
                                shard of the variants -- its own .twk blocks followed by the halo blocks a -w window can reach
                                (twkb_plan_shards) -- and computes only the pairs whose earlier member lies in its first
                                shard_blocks blocks; the union over the shards is the whole-matrix -w result. 0 = off;
                                needs window = 1. The block structure is the one twkb_set_blocks gave (or twk_block_size) */
    int32_t sorted_output;   /* twkb_calc_file*: 1 = order the records on the device (twkb_compute_sorted) and write a SORTED .two
                                (the reference needs calc, then sort); one device, output to a file */
} twkb_settings;

/* Subset of twk1_t (include/core.h:291-295) the LD path reads. */
typedef struct twkb_variant {
    uint32_t rid;       /* contig id                                   */
    uint32_t pos;       /* 0-based position                            */
    uint32_t ac;        /* alt allele count                            */
    uint32_t an;        /* number of MISSING alleles (reference naming) */
    double hwe;         /* Hardy-Weinberg P                            */
    uint8_t gt_missing; /* variant has a missing mask                  */
    uint8_t gt_phase;   /* all genotypes phased                        */
    uint8_t pad[6];
} twkb_variant;

/* Where the run-length genotype words of one variant sit inside a byte buffer (a .twk block body,
 * lib/core.cpp:75-101). A run word is twk1_igt_t<uint8_t|uint16_t|uint32_t> (include/core.h:188-256):
 * miss = 0: len << 2 | alleleA << 1 | alleleB; miss = 1: len << 4 | alleleA << 2 | alleleB with
 * allele codes 0 ref, 1 alt, 2 missing (lib/genotype_encoder.h:11-17); len counts samples. */
typedef struct twkb_run_desc {
    uint64_t offset; /* byte offset of the first run word (any alignment) */
    uint32_t n_runs;
    uint8_t width;   /* bytes per run word: 1, 2 or 4 */
    uint8_t miss;    /* 1 = 2-bit allele codes */
    uint8_t pad[2];
} twkb_run_desc;

/* Counters of one twkb_compute call (reference: twk_ld_progress n_var/n_out). */
typedef struct twkb_stats {
    uint64_t pairs_visited;   /* reference n_var: every pair of every processed tile   */
    uint64_t pairs_screened;  /* pairs that survived the in-kernel R2 pre-screen       */
    uint64_t records_out;     /* forward records handed to the sink                    */
    uint64_t count_launches;  /* count-kernel launches                                 */
    uint64_t stats_launches;  /* statistics-kernel launches                            */
    uint64_t other_launches;  /* pack / expand / memset kernels                        */
    double seconds_total;     /* wall clock of the call                                */
    double ms_count_kernel;   /* CUDA-event time summed over count-kernel launches     */
    double ms_stats_kernel;   /* CUDA-event time summed over stats-kernel launches     */
    double ms_h2d;            /* CUDA-event time of the matrix upload + device packing */
    uint64_t bytes_h2d, bytes_d2h;
    int32_t kernel_used;      /* TWKB_KERNEL_POPC or TWKB_KERNEL_UMMA                  */
    int32_t n_planes;         /* bit planes per variant on the device                  */
    uint64_t word_ops;        /* 32-bit AND+POPC word operations (POPC kernel)         */
    uint64_t mma_macs;        /* int8 multiply-accumulates issued (UMMA kernel)        */
    double ms_device_total;   /* CUDA-event time from the first launch of the call to
                                 the completion of its last kernel / copy             */
    uint64_t sparse_variants; /* variants served by the list (sparse) kernel           */
    uint64_t sparse_launches; /* sparse-kernel launches                                */
    uint64_t sparse_word_ops; /* AND+POPC word operations issued by the sparse kernel  */
    double ms_sparse_kernel;  /* CUDA-event time summed over sparse-kernel launches    */
    double ms_decode_kernel;  /* CUDA-event time of decode_runs_kernel (twkb_load_runs)  */
    /* twkb_calc_file only: wall clock of its phases */
    double seconds_file_read;  /* .twk read + zstd inflate (+ host unpack when host_unpack) */
    double seconds_file_load;  /* upload + device decode / transpose                        */
    double seconds_file_total; /* open to the closed .two file                               */
} twkb_stats;

/* Receives `n` packed 106-byte records (forward orientation: A is the variant
 * with the lower index). Called from the context's record-drain thread (never
 * concurrently, and never after twkb_compute has returned) while the device
 * already computes the next batch. Return non-zero to abort the run
 * (-> TWKB_ESINK). The reverse copies the reference also writes
 * (lib/ld/ld_engine.cpp:1290-1298) are synthesised by the .two writer
 * (twkb_two_writer_*), not by the device. */
typedef int (*twkb_sink_fn)(void* user, const uint8_t* records, uint64_t n);

void twkb_settings_init(twkb_settings* s); /* reference defaults, lib/core.cpp:297-306 */

int twkb_create(const twkb_settings* s, void** ctx);
void twkb_destroy(void* ctx);
const char* twkb_last_error(void* ctx); /* ctx may be NULL: last create error */

/* Replace thresholds/mode flags between runs without re-uploading the matrix. */
int twkb_update_settings(void* ctx, const twkb_settings* s);

/* Upload the genotype matrix. Rows follow twk_igt_vec (lib/core.cpp:349-383):
 * haplotype p of a variant is bit (p % 64) of word (p / 64); sample s owns bits
 * 2s and 2s+1; mask rows (nullable) have both bits of a sample set when either
 * allele is missing. Host memory is copied; the caller keeps ownership.
 * The device keeps the matrix transposed (word-major) -- see DESIGN.md. */
int twkb_load_matrix(void* ctx, uint32_t n_samples, uint32_t n_variants, const uint64_t* data_bits,
                     const uint64_t* mask_bits, size_t row_stride_words, const twkb_variant* meta);

/* Same, for rows that already live in device memory of this context's device
 * (e.g. received by an NCCL broadcast). Pointers are CUDA device pointers. */
int twkb_load_matrix_device(void* ctx, uint32_t n_samples, uint32_t n_variants, const uint64_t* d_data_bits,
                            const uint64_t* d_mask_bits, size_t row_stride_words, const twkb_variant* meta);

/* Upload run-length encoded genotypes and decode them ON THE DEVICE into the same resident rows
 * (+ mask rows) twkb_load_matrix would have uploaded -- the device counterpart of
 * twk_igt_vec::Build (lib/core.cpp:349-383). run_bytes is host memory (copied); desc[v] locates
 * variant v's run words in it. Fails with TWKB_EINVAL if the runs of a variant do not cover
 * exactly n_samples samples. */
int twkb_load_runs(void* ctx, uint32_t n_samples, uint32_t n_variants, const uint8_t* run_bytes, size_t n_run_bytes,
                   const twkb_run_desc* desc, const twkb_variant* meta);

/* ---- multi-GPU data plane: NCCL over NVLink, used once per load ----------------------------------------
 * Reference analogue: twk_ld::LoadAllBlocks (lib/ld/ld.cpp:370-465) unpacks the file once into host memory
 * every slave thread shares; twk_ld_balancer (lib/ld/ld_balancing.h:23-80) then deals block pairs. Here one
 * context per GPU (threads of one process -- include/twkb_ld.hpp, `twkb_calc -g 0,1,..` -- or one process per
 * GPU -- bench.py under torchrun) forms a communicator; a sliced load sends only rows
 * [row_begin, row_end) of this rank (twkb_comm_slice) over this GPU's own PCIe link and completes the matrix
 * on every GPU with one in-place ncclAllGather over NVLink. No collective runs during twkb_compute: tiles are
 * independent (settings.part_index / part_count select this context's tiles).
 * libnccl.so.2 is loaded on first use; without it these calls fail with TWKB_ENODEVICE. */
#define TWKB_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
int twkb_comm_unique_id(uint8_t* id /* [TWKB_COMM_ID_BYTES] */);  /* one rank creates it, all ranks pass it to init */
int twkb_comm_init(void* ctx, const uint8_t* id, int32_t rank, int32_t n_ranks); /* collective (ncclCommInitRank) */
int twkb_comm_slice(uint32_t n_variants, int32_t rank, int32_t n_ranks, uint32_t* row_begin, uint32_t* row_end);
/* Collective. slice_*_bits point at row `row_begin` of this rank (the rows of the slice, contiguous, same
 * layout as twkb_load_matrix); meta holds ALL n_variants entries on every rank. Without a communicator
 * (or n_ranks == 1) identical to twkb_load_matrix. */
int twkb_load_matrix_sliced(void* ctx, uint32_t n_samples, uint32_t n_variants, const uint64_t* slice_data_bits,
                            const uint64_t* slice_mask_bits, size_t row_stride_words, const twkb_variant* meta);
/* Collective. Arguments as twkb_load_runs (every rank sees all descriptors, e.g. the threads of one process
 * sharing the inflated file); a rank uploads and decodes the run words of its own variants only. */
int twkb_load_runs_sliced(void* ctx, uint32_t n_samples, uint32_t n_variants, const uint8_t* run_bytes, size_t n_run_bytes,
                          const twkb_run_desc* desc, const twkb_variant* meta);

/* The .twk block structure of the loaded variants: block_first[b] = index (file order) of the first variant of block b,
 * n_blocks entries, strictly increasing from 0. The reference's -w rules (balancer row prune, block-pair abort) and its -c / -C
 * chunks are defined on .twk blocks, whose length `import -b` makes configurable; call this after a load (twkb_calc_file* and
 * the C++ mirror pass the file's index). Without it blocks of settings.twk_block_size variants per contig are assumed. */
int twkb_set_blocks(void* ctx, const uint32_t* block_first, uint32_t n_blocks);
/* Borrow the block structure of an open .twk handle (n_blocks + 1 entries, the last one = n_variants). */
int twkb_twk_blocks(void* handle, const uint32_t** block_first, uint32_t* n_blocks);

/* Test hook: copy the resident reference-layout rows (file order unless the rare-variant class
 * re-ordered them) back to the host. mask_bits may be NULL. */
int twkb_debug_rows(void* ctx, uint64_t* data_bits, uint64_t* mask_bits, size_t row_stride_words);

/* Run the LD computation over this context's share of the pair grid. */
int twkb_compute(void* ctx, twkb_sink_fn sink, void* user);

/* Like twkb_compute, but keeps the records in a device buffer and only counts
 * them (used to time the device-resident path; no D2H of records). */
int twkb_compute_resident(void* ctx);

/* Like twkb_compute, with the output already in the order of `tomahawk sort` (two_reader::Sort, lib/two_reader.cpp:162-420;
 * twk1_two_t::operator<, lib/core.cpp:458-468: ridA, ridB, posA, posB): the forward records stay in device memory until the
 * computation ends, both orientations (the reverse copies the reference's calc also writes) are ordered by a radix sort ON
 * THE DEVICE, and the sink receives forward AND reverse records in file order -- write them with twkb_two_open_sorted /
 * twkb_two_add_sorted and the reference's `view -I` can seek in the file without a `sort` run. One device (part_count <= 1);
 * device memory: 106 B per forward record + 40 B per sorted item. */
int twkb_compute_sorted(void* ctx, twkb_sink_fn sink, void* user);

/* Downstream consumer fed from the device-resident records: LD decay over distance, the reference's
 * two_reader::Decay (lib/two_reader.cpp:424-475), without writing and re-reading a .two file. Runs the LD computation like
 * twkb_compute_resident and reduces every record with ridA == ridB and posA < posB into
 *     bin = min((posB - posA) / (window_bp / n_bins), n_bins - 1):  sum_r2[bin] += R2, count[bin] += 1
 * on the device; the reference prints From = bin * width, To = (bin + 1) * width, Mean = sum / max(count, 1), Frequency = count.
 * Only 16 bytes per bin leave the GPU. sum_r2 / count: host arrays of n_bins entries. */
int twkb_compute_decay(void* ctx, int64_t window_bp, int32_t n_bins, double* sum_r2, uint64_t* count);

/* Downstream consumer fed from the device-resident records: `tomahawk aggregate`, the reference's
 * two_reader::Aggregate (lib/two_reader.cpp:543-853, lib/aggregation.h:127-175) -- a xbins x ybins raster of the LD landscape --
 * without writing and re-reading a .two file. Every record lands in bin [coord(ridA, posA) / bpx][coord(ridB, posB) / bpy] and, as
 * the reverse copy the reference writes, in [coord(ridB, posB) / bpx][coord(ridA, posA) / bpy]; a bin keeps the reference's
 * twk_sstats (include/core.h:929-990: n, total, total_squared, min and max -- both of which start at 0 there) of the chosen field.
 * Two passes like the reference: the first finds the position range of every contig that occurs in a record, the second
 * rasterises; the LD computation runs once per pass and only the raster leaves the GPU (40 bytes per bin).
 * Coordinates (two_reader.cpp:743-797, aggregation.h:157): one contig with records (any contig but the first -- the
 * reference counts contig 0 twice, :737-740; kept with settings.emulate_quirks) spans [min, max] of the positions seen,
 * otherwise every contig with records spans its whole length contig_n_bases[rid]; bpx = ceil((float)range / xbins).
 * Not reproduced: the reference skips .two blocks with fewer than 5 records (aggregation.h:131,152; depends on how its
 * writer threads happened to cut the file) and indexes one bin past the raster for the very last coordinate (clamped here). */
#define TWKB_AGG_R2 0
#define TWKB_AGG_R 1
#define TWKB_AGG_D 2
#define TWKB_AGG_DPRIME 3
#define TWKB_AGG_P 4
#define TWKB_AGG_HETS 5 /* (cnt[1] + cnt[2]) / sum(cnt), twk_sstats::AddHets */
#define TWKB_AGG_ALTS 6 /* cnt[3] / sum(cnt),            twk_sstats::AddAlts */
typedef struct twkb_agg_bin { /* twk_sstats */
    uint64_t n;
    double total, total_squared, min, max;
} twkb_agg_bin;
typedef struct twkb_agg_layout { /* twk1_aggregate_t header fields, include/core.h:1014-1017 */
    uint64_t range;      /* bases covered by the raster                */
    uint32_t bpx, bpy;   /* bases per bin                              */
    uint32_t n_contigs_set;
    uint64_t n_records;  /* records rasterised, reverse copies included */
} twkb_agg_layout;
/* bins: xbins * ybins entries, row-major [x][y]. contig_offset / contig_min / contig_max (nullable, n_contigs entries each):
 * the reference's rid_offsets (cumulative range, min, max). */
int twkb_compute_aggregate(void* ctx, int32_t field, int32_t xbins, int32_t ybins, const int64_t* contig_n_bases, uint32_t n_contigs,
                           twkb_agg_bin* bins, twkb_agg_layout* layout, uint64_t* contig_offset, uint32_t* contig_min,
                           uint32_t* contig_max);
/* Contig lengths of an open .twk handle in header order (what twkb_compute_aggregate takes). n_bases may be NULL to count. */
int twkb_twk_contigs(void* handle, int64_t* n_bases, uint32_t capacity, uint32_t* n_contigs);

int twkb_get_stats(void* ctx, twkb_stats* out);

/* Test hook: run the count kernel the settings select (tensor-core or LOP3+POPC, exactly as
 * twkb_compute would) over this context's share of the grid and return
 * the raw candidate entries (12 uint32 each: i, j, c[9], mode) instead of records.
 * screen_off != 0 disables the R2 pre-screen so every enumerated pair is returned
 * with its exact contingency counts: mode 0 -> c = {REFREF, slot1 (A alt,B ref),
 * slot4 (A ref,B alt), ALTALT}; mode 1 -> c = 3x3 genotype table t[gA][gB].
 * out may be NULL to query the count. */
int twkb_debug_candidates(void* ctx, int screen_off, uint32_t* out, uint64_t capacity, uint64_t* n_out);

/* ---- host I/O around the engine (reference: lib/twk_reader.cpp, include/writer.h) ----
 * Host-only (no CUDA call); zstd runs on the host. */

/* `tomahawk calc` end to end = twk_ld::Compute(settings) (lib/ld/ld.cpp:477-671):
 * read in_path (.twk), compute on settings->device, write out_path (.two; the
 * reference's rule of forcing a ".two" suffix applies, ld.cpp:589-598). */
int twkb_calc_file(const twkb_settings* s, const char* in_path, const char* out_path, twkb_stats* stats_out,
                   char* errbuf, size_t errbuf_len);

/* Same with the `-I` interval strings of `calc` (twk_ld_settings::ival_strings, include/core.h:923;
 * grammar "chr", "chr:pos", "chr:from-to", lib/intervals.cpp:91-136). The reference works at .twk
 * BLOCK granularity (lib/ld/ld.cpp:257-365): every variant of every block overlapping an interval
 * takes part. With settings->emulate_quirks the reference's loading rule is reproduced exactly
 * (n consecutive blocks from the first overlapping one); without it, the distinct union. */
int twkb_calc_file_intervals(const twkb_settings* s, const char* in_path, const char* out_path,
                             const char* const* intervals, int32_t n_intervals, twkb_stats* stats_out, char* errbuf,
                             size_t errbuf_len);

/* `tomahawk scalc` = twk_ld::ComputeSingle (lib/ld/ld.cpp:673-876): settings->single = 1 and ONE interval string naming
 * the target site ("chr:pos", 1-based) go through twkb_calc_file_intervals; the variants within settings->l_surrounding
 * bases on either side are loaded (twk_ld_impl::LoadTargetSingle, lib/ld/ld.cpp:123-255). The reader below does the same
 * selection for callers that drive the engine themselves. */
int twkb_twk_open_single(const char* path, int n_threads, const char* interval, int32_t l_surrounding, int32_t emulate_quirks,
                         int32_t runs_mode, void** handle, uint32_t* n_targets, char* errbuf, size_t errbuf_len);
/* emulate_quirks: the reference gathers the neighbours in blocks of 100 and drops the last, partial block (lib/ld/ld.cpp:193-195,
 * :241); with 0 every neighbour takes part. */

/* .twk reader (twk_reader::Open + twk1_blk_iterator::NextBlock + twk_igt_vec::Build):
 * unpacks every block into the row layout twkb_load_matrix takes. */
int twkb_twk_open(const char* path, int n_threads, void** handle, char* errbuf, size_t errbuf_len);
int twkb_twk_open_intervals(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                            int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len);
int twkb_twk_dims(void* handle, uint32_t* n_samples, uint32_t* n_variants, size_t* row_stride_words,
                  int32_t* any_missing, uint32_t* n_blocks);
int twkb_twk_copy(void* handle, uint64_t* data_bits, uint64_t* mask_bits /* nullable */, twkb_variant* meta);
/* Borrow the unpacked rows instead of copying them (valid until twkb_twk_close; mask_bits is set
 * to NULL when no variant has missing genotypes). */
int twkb_twk_view(void* handle, const uint64_t** data_bits, const uint64_t** mask_bits, const twkb_variant** meta);
/* Runs mode: inflate the blocks but leave the genotypes run-length encoded (no host unpack);
 * twkb_twk_runs_view then yields what twkb_load_runs takes. twkb_twk_copy / twkb_twk_view fail with
 * TWKB_ESTATE on such a handle; twkb_twk_dims and twkb_two_open work on both kinds. */
int twkb_twk_open_runs(const char* path, int n_threads, const char* const* intervals, int32_t n_intervals,
                       int32_t emulate_quirks, void** handle, char* errbuf, size_t errbuf_len);
int twkb_twk_runs_view(void* handle, const uint8_t** run_bytes, size_t* n_run_bytes, const twkb_run_desc** desc,
                       const twkb_variant** meta);
void twkb_twk_close(void* handle);

/* .two writer (twk_two_writer_t + twk_ld_engine::CompressFwd/Rev + IndexOutput): takes
 * forward records, writes forward and reverse blocks of <= b_size records, index, EOF.
 * twk_handle supplies the VcfHeader that is copied into the .two header. */
int twkb_two_open(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t b_size,
                  void** writer, char* errbuf, size_t errbuf_len);
int twkb_two_set_threads(void* writer, int32_t n_threads); /* zstd block compression threads (default 1) */
int twkb_two_add(void* writer, const uint8_t* records, uint64_t n);
int twkb_two_close(void* writer); /* finishes the file and frees the writer */

/* Writer of a SORTED .two (what `tomahawk sort` produces): takes records that are already in order (twkb_compute_sorted:
 * forward and reverse copies), cuts blocks of <= 10,000 records at every change of ridA and writes the sorted-state index
 * with the per-contig entries (lib/two_reader.cpp:350-420, include/writer.h:363-374). A failed close removes the file. */
int twkb_two_open_sorted(const char* path, void* twk_handle, const char* command_line, int32_t c_level, int32_t n_threads,
                         void** writer, char* errbuf, size_t errbuf_len);
int twkb_two_add_sorted(void* writer, const uint8_t* records, uint64_t n);
int twkb_two_close_sorted(void* writer);

/* `tomahawk sort` (two_reader::Sort, lib/two_reader.cpp:162-420):
 * orders the records by (ridA, ridB, posA, posB) -- twk1_two_t::operator<, lib/core.cpp:458-468 -- and
 * writes blocks cut at every change of ridA with the sorted-state index (per-block rid / ridB / minpos /
 * maxpos and the per-contig entries) that lets the reference's `view -I` seek. Host only; zstd blocks are
 * inflated and compressed by up to n_threads threads. n_records may be NULL. */
int twkb_two_sort(const char* in_path, const char* out_path, int32_t c_level, int32_t n_threads, uint64_t* n_records,
                  char* errbuf, size_t errbuf_len);
/* The same with a memory budget in bytes (the reference's `sort -m`; 0 = unbounded): a file whose records + sort keys exceed it
 * is sorted in runs spilled to temporary files "<out>_<pid>_<k>.tmp" and merged k-way (two_reader::Sort's external merge,
 * lib/two_reader.cpp:262-420). The output is identical whatever the budget. A failed run removes its partial output. */
int twkb_two_sort_mem(const char* in_path, const char* out_path, int32_t c_level, int32_t n_threads, uint64_t memory_limit_bytes,
                      uint64_t* n_records, char* errbuf, size_t errbuf_len);

/* Tile scheduler, host only (the B200 counterpart of twk_ld_balancer /
 * twk_ld_dynamic_balancer, lib/ld/ld_balancing.h): the (i0, j0) variant offsets of the
 * tile_i x tile_j tiles that settings->part_index of settings->part_count computes
 * (-c/-C chunk, -w window band and diagonal rules applied). out may be NULL to count. */
int twkb_plan_tiles(const twkb_settings* s, uint32_t n_variants, const twkb_variant* meta, uint32_t tile_i,
                    uint32_t tile_j, uint32_t* out_ij, uint64_t capacity, uint64_t* n_tiles, uint64_t* n_pairs);

/* Position shards of a -w run, host only ("position-sharding with +-window halo"). The variants are cut
 * at .twk block boundaries into n_shards consecutive ranges of about equal pair work; shard k owns blocks
 * [own_begin[k], own_begin[k+1]) and must also hold the halo blocks up to halo_end[k] (exclusive): every block the
 * reference's row prune (ld_balancing.h:189-196, positions only) leaves reachable from one of its own blocks. A context
 * that loads variants block_first[own_begin[k]] .. block_first[halo_end[k]] - 1 with settings.shard_blocks =
 * own_begin[k+1] - own_begin[k] (and the same blocks, re-based, through twkb_set_blocks) computes exactly the pairs of
 * the whole-matrix run whose earlier member it owns. block_first: n_blocks + 1 entries (last = n_variants);
 * own_begin: n_shards + 1 entries out; halo_end: n_shards entries out. */
int twkb_plan_shards(const uint32_t* block_first, uint32_t n_blocks, const twkb_variant* meta, uint32_t n_variants,
                     int32_t l_window, int32_t n_shards, uint32_t* own_begin, uint32_t* halo_end);

int twkb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TWKB_H_ */
