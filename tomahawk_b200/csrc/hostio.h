// hostio.h -- host-side file formats around the engine: the `.twk` reader
// (reference lib/twk_reader.cpp:49-125, lib/core.cpp:75-101,253-261,349-383) and
// the `.two` writer (reference include/writer.h:70-87,225-242,293-313,
// lib/ld/ld_engine.cpp:1268-1298,1742-1810, lib/index.cpp:41-52,242-251).
// zstd stays on the host (BASELINE.json north_star).
#pragma once
#include <cstdint>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/twkb.h"

namespace twkb {

// Declarations of the stable zstd C ABI we link against (libzstd.so.1); the
// image ships the runtime library but no headers.
extern "C" {
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char* ZSTD_getErrorName(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
}

// Uninitialised byte buffer (malloc): the pages are first touched by the threads that fill them.
struct ByteBuf {
    uint8_t* p = nullptr;
    size_t n = 0;
    ByteBuf() = default;
    ByteBuf(const ByteBuf&) = delete;
    ByteBuf& operator=(const ByteBuf&) = delete;
    ~ByteBuf() { release(); }
    bool alloc(size_t bytes) {
        release();
        p = static_cast<uint8_t*>(std::malloc(bytes ? bytes : 1));
        n = p ? bytes : 0;
        return p != nullptr;
    }
    void release() {
        std::free(p);
        p = nullptr;
        n = 0;
    }
    uint8_t* data() { return p; }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
};

// A whole .twk file unpacked into the matrix layout twkb_load_matrix takes.
struct TwkFile {
    uint32_t n_samples = 0;
    uint32_t n_variants = 0;
    uint32_t n_targets = 0;  // single mode (scalc): the first n_targets variants are the target site(s)
    size_t stride = 0;  // u64 words per row (128-bit aligned)
    bool any_missing = false;
    std::vector<uint64_t> data, mask;
    std::vector<twkb_variant> meta;
    // VcfHeader pieces (lib/header.cpp:330-345) needed to write the .two header
    std::string fileformat, literals;
    std::string header_tail;  // serialized samples + contigs, copied through verbatim
    uint32_t n_contigs = 0;
    std::vector<int64_t> contig_n_bases;  // VcfContig::n_bases in header order (the aggregate consumer's coordinate system)
    uint32_t n_blocks = 0;
    std::vector<uint32_t> block_first;  // first variant of every loaded .twk block (+ n_variants at the end): the block structure the
                                        // reference's window rule and -c chunks are defined on (`import -b` makes the length configurable)
    // Runs mode (keep_runs): the rows are NOT unpacked on the host. `raw` holds the inflated .twk
    // blocks back to back and run_desc[v] locates the run-length words of variant v inside it
    // (twk1_igt_t, include/core.h:188-256); the device decodes them (decode.cuh, twkb_load_runs).
    bool runs_mode = false;
    ByteBuf raw;
    std::vector<twkb_run_desc> run_desc;
};

// Reads and unpacks `path` with up to n_threads host threads. Returns 0 or TWKB_EIO.
// `intervals` (nullable): the -I strings of `calc` ("chr", "chr:pos", "chr:from-to"); only the
// .twk blocks that overlap them are loaded (reference lib/ld/ld.cpp:257-365, lib/intervals.cpp).
// keep_runs: leave the genotypes run-length encoded for the device decoder (see TwkFile::raw).
// single_surrounding >= 0: scalc selection (lib/ld/ld.cpp:123-255) -- `intervals` holds ONE string naming the target site(s);
// the result holds [targets | variants within single_surrounding bases], see TwkFile::n_targets.
int read_twk(const std::string& path, int n_threads, TwkFile& out, std::string& err,
             const std::vector<std::string>* intervals = nullptr, bool emulate_quirks = true, bool keep_runs = false,
             int32_t single_surrounding = -1);

// Streaming .two writer: takes forward records, writes forward and reverse
// blocks of <= b_size records, the index and the EOF marker. Finished blocks queue up and are
// zstd-compressed by up to `threads` host threads at a time, then written in order: the file is
// byte-identical to the single-threaded one. With more than one thread add() only queues a copy
// of the records and a writer thread does the splitting, compression and file I/O, so the
// caller (the device batch loop) never waits for zstd; errors surface at the next add()/finish().
class TwoWriter {
public:
    TwoWriter() = default;
    ~TwoWriter();
    int open(const std::string& path, const TwkFile& src, const std::string& command_line, int c_level, int b_size,
             std::string& err);
    int add(const uint8_t* records, uint64_t n);  // forward records, TWKB_RECORD_BYTES each
    int finish();
    void set_threads(int n);
    uint64_t records_written() const { return n_written_; }
    const std::string& error() const { return err_; }

private:
    struct IndexEntry {
        int32_t rid, ridB;
        uint32_t n, minpos, maxpos, b_unc, b_cmp;
        uint64_t foff, fend;
    };
    struct Side {
        std::vector<uint8_t> buf;  // records only
        uint32_t n = 0;
        IndexEntry ent{};
    };
    struct Pending {  // a finished block waiting for compression
        std::vector<uint8_t> raw, z;
        size_t zn = 0;
        IndexEntry ent{};
    };
    int flush_side(Side& s);
    int drain();  // compress (in parallel) and write every pending block
    int add_sync(const uint8_t* records, uint64_t n);
    void writer_loop();
    int stop_writer();

    bool emit(const void* p, size_t n);

    FILE* fp_ = nullptr;
    bool to_stdout_ = false;  // path "-": stream the file to stdout (the reference's default out)
    uint64_t pos_ = 0;        // bytes written so far (index offsets; a pipe has no ftell)
    int c_level_ = 1;
    uint32_t b_size_ = 10000;
    uint32_t n_contigs_ = 0;
    Side fwd_, rev_;
    std::vector<IndexEntry> index_;
    std::vector<Pending> pending_;
    int threads_ = 1;
    // asynchronous mode (threads_ > 1)
    std::thread writer_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::vector<uint8_t>> queue_;
    size_t queued_bytes_ = 0;
    bool stop_ = false;
    int async_rc_ = 0;
    uint64_t n_written_ = 0;
    std::string err_;
};

// `tomahawk sort` for files that fit in host memory (reference two_reader::Sort,
// lib/two_reader.cpp:162-420, which merges per-thread sorted runs through temporary files): reads
// every block of `in`, orders the records by twk1_two_t::operator< (ridA, ridB, posA, posB;
// lib/core.cpp:458-468) and writes `out` the way the reference's merge phase does -- blocks of
// <= 10,000 records cut at every change of ridA, index state TWK_IDX_SORTED with per-block
// (rid, ridB, minpos, maxpos) and the per-contig summary entries that make `view -I` seek
// (include/writer.h:344-390, lib/index.cpp:70-88).
// memory_limit (bytes, 0 = unbounded): inputs whose records + keys exceed it are sorted in runs that are spilled to
// temporary files next to `out` and merged k-way, like the reference's external merge (lib/two_reader.cpp:262-420).
// Writer of a SORTED .two from a record stream that is already in twk1_two_t::operator< order (both orientations): blocks cut
// at every change of ridA, sorted-state index with per-contig entries -- the file `tomahawk sort` would have produced
// (two_reader::Sort, lib/two_reader.cpp:350-420). add() rejects a record that sorts before its predecessor.
class SortedTwoWriter {
public:
    SortedTwoWriter() = default;
    ~SortedTwoWriter();
    SortedTwoWriter(const SortedTwoWriter&) = delete;
    SortedTwoWriter& operator=(const SortedTwoWriter&) = delete;
    int open(const std::string& path, const TwkFile& src, const std::string& command_line, int c_level, int n_threads, std::string& err);
    int add(const uint8_t* records, uint64_t n);
    int finish();
    const std::string& error() const { return err_; }

private:
    struct Impl;
    Impl* p_ = nullptr;
    std::string err_;
};

int sort_two(const std::string& in, const std::string& out, int c_level, int n_threads, std::string& err,
             uint64_t* n_records = nullptr, uint64_t memory_limit = 0);

}  // namespace twkb
