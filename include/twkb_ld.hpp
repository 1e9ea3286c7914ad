// twkb_ld.hpp -- C++ host mirror of the reference's LD driver, header-only, on top of the
// C-ABI of twkb.h (libtwkb.so). It keeps the reference's interface for this path:
//
//   reference                                            | here (namespace twkb_host)
//   -----------------------------------------------------+------------------------------
//   struct twk_ld_settings      include/core.h:909-924   | struct twk_ld_settings (same field
//     defaults                  lib/core.cpp:297-306     |   names, meaning and defaults)
//     GetString()               lib/core.cpp:308-332     |   GetString()
//   class twk_ld                include/ld.h:40-69       | class twk_ld
//     bool Compute(const twk_ld_settings&)  ld.h:53      |   bool Compute(const twk_ld_settings&)
//     bool Compute()            lib/ld/ld.cpp:477-671    |   bool Compute()
//
// Error behaviour follows the reference: a message "[date][ERROR] ..." on stderr and a `false`
// result; the stages log "[date][LOG]..." lines like lib/ld/ld.cpp does. There is no CPU compute
// path: without an sm_100 device Compute() fails with the library's TWKB_ENODEVICE message.
//
// B200 additions to the settings: `devices` (CUDA ordinals; one host thread and one device
// context per entry, the tile grid dealt between them -- the replacement of the reference's
// n_threads worker slaves, lib/ld/ld.cpp:623-644) and `kernel` (TWKB_KERNEL_*).
#ifndef TWKB_LD_HPP_
#define TWKB_LD_HPP_

#include <sys/time.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <ctime>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "twkb.h"

namespace twkb_host {

inline std::string datetime() {  // "YYYY-MM-DD HH:MM:SS,mmm", lib/utility.cpp:60-86
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    struct tm tmv;
    localtime_r(&tv.tv_sec, &tmv);
    char buf[64], out[80];
    std::strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", &tmv);
    std::snprintf(out, sizeof(out), "%s,%03d", buf, (int)(tv.tv_usec / 1000));
    return out;
}
inline std::string timestamp(const std::string& type) { return "[" + datetime() + "][" + type + "] "; }
inline std::string timestamp(const std::string& type, const std::string& type2) {
    return "[" + datetime() + "][" + type + "][" + type2 + "] ";
}
inline std::string pretty(uint64_t v) {  // utility::ToPrettyString: thousands separators
    std::string s = std::to_string(v), o;
    for (size_t i = 0; i < s.size(); ++i) {
        o += s[i];
        const size_t left = s.size() - 1 - i;
        if (left && left % 3 == 0) o += ',';
    }
    return o;
}

struct twk_ld_settings {
    bool square = true, window = false, low_memory = false, bitmaps = false, single = false;
    bool force_phased = false, forced_unphased = false, force_cross_intervals = false;
    int32_t c_level = 1, bl_size = 500, b_size = 10000, l_window = 1000000;
    int32_t n_threads = (int32_t)std::thread::hardware_concurrency(), cycle_threshold = 0, ldd_load_type = 7;
    int32_t l_surrounding = 500000;
    std::string in, out = "-";
    double minP = 1, minR2 = 0.1, maxR2 = 100, minDprime = 0, maxDprime = 100;
    int32_t n_chunks = 1, c_chunk = 0;
    std::vector<std::string> ival_strings;
    // --- B200 additions ---
    std::vector<int32_t> devices{0};
    int32_t kernel = TWKB_KERNEL_AUTO;
    bool emulate_quirks = true;
    bool host_unpack = false;  // true: unpack the .twk rows on the host instead of decoding the runs on the device
    bool silent = false;  // suppress the LOG lines (errors are always printed)
    bool sorted_output = false;  // order the records on the device and write a SORTED .two (calc + sort of the reference in one run);
                                 // one device, output to a file
    bool position_shards = true;  // -w on several devices: every device loads only its own .twk blocks + the halo blocks the
                                  // window reaches (twkb_plan_shards) instead of the whole matrix; false = deal tiles of
                                  // the whole matrix like the all-pairs modes

    std::string GetString() const {
        auto tf = [](bool b) { return std::string(b ? "TRUE" : "FALSE"); };
        std::string s = "square=" + tf(square) + ",window=" + tf(window) + ",low_memory=" + tf(low_memory) + ",bitmaps=" + tf(bitmaps) +
                        ",single=" + tf(single) + ",force_phased=" + tf(force_phased) + ",force_unphased=" + tf(forced_unphased) +
                        ",compression_level=" + std::to_string(c_level) + ",block_size=" + std::to_string(bl_size) +
                        ",output_block_size=" + std::to_string(b_size) +
                        (window ? std::string(",window_size=") + std::to_string(l_window) : "") +
                        ",l_surrounding=" + std::to_string(l_surrounding) + ",minP=" + std::to_string(minP) +
                        ",minR2=" + std::to_string(minR2) + ",maxR2=" + std::to_string(maxR2) + ",minDprime=" + std::to_string(minDprime) +
                        ",maxDprime=" + std::to_string(maxDprime) + ",n_chunks=" + std::to_string(n_chunks) +
                        ",c_chunk=" + std::to_string(c_chunk) + ",n_threads=" + std::to_string(n_threads) +
                        ",ldd_type=" + std::to_string((int)ldd_load_type) + ",cycle_threshold=" + std::to_string(cycle_threshold);
        s += ",devices=";
        for (size_t i = 0; i < devices.size(); ++i) s += (i ? "+" : "") + std::to_string(devices[i]);
        return s;
    }

    void ToC(twkb_settings* c) const {
        twkb_settings_init(c);
        c->square = square; c->window = window; c->low_memory = low_memory; c->bitmaps = bitmaps; c->single = single;
        c->force_phased = force_phased; c->forced_unphased = forced_unphased; c->emulate_quirks = emulate_quirks;
        c->c_level = c_level; c->bl_size = bl_size; c->b_size = b_size; c->l_window = l_window;
        c->n_threads = n_threads > 0 ? n_threads : 1; c->l_surrounding = l_surrounding;
        c->n_chunks = n_chunks; c->c_chunk = c_chunk;
        c->minP = minP; c->minR2 = minR2; c->maxR2 = maxR2; c->minDprime = minDprime; c->maxDprime = maxDprime;
        c->kernel = kernel;
        c->host_unpack = host_unpack ? 1 : 0;
        c->single_targets = 0;
    }
};

class twk_ld {
public:
    twk_ld() = default;
    void operator=(const twk_ld_settings& s) { settings = s; }

    bool Compute(const twk_ld_settings& s) {
        settings = s;
        return Compute();
    }

    // twk_ld::ComputeSingle (include/ld.h:54, lib/ld/ld.cpp:673-876): the target site(s) named by the one interval
    // string against every variant within l_surrounding bases (`scalc`).
    bool ComputeSingle(const twk_ld_settings& s) {
        settings = s;
        settings.single = true;
        return Compute();
    }

    // twk_ld::Compute, lib/ld/ld.cpp:477-671: open, select blocks, load, compute, write.
    bool Compute() {
        stats = twkb_stats{};
        if (settings.in.empty()) return error("No file-name provided...");
        if (settings.window && settings.n_chunks != 1) return error("Cannot use chunking in window mode!");
        if (settings.devices.empty()) return error("No device selected...");
        if (settings.single) {  // ld.cpp:679-697
            if (settings.n_chunks != 1) return error("Cannot use chunking in single mode!");
            if (settings.window) return error("Cannot use window in single mode!");
            if (settings.ival_strings.empty()) return error("An interval has to be provided in single mode!");
            if (settings.ival_strings.size() != 1) return error("Only a single interval can be provided in single mode!");
            settings.devices.resize(1);  // one target row: one device
        }
        const bool to_stdout = settings.out.empty() || settings.out == "-";  // the reference's default: stream the .two to stdout (ld.cpp:585-588)
        if (settings.sorted_output) {
            if (to_stdout) return error("Sorted output needs an output file (-o)...");
            if (settings.devices.size() != 1) return error("Sorted output runs on one device...");
        }
        log("READER") << "Opening " << settings.in << "..." << std::endl;
        char errbuf[1024] = {0};
        std::vector<const char*> iv;
        for (const std::string& x : settings.ival_strings) iv.push_back(x.c_str());
        void* twk = nullptr;
        // default: blocks are inflated on the host, the run-length genotypes are decoded on the device
        const bool runs = !settings.host_unpack;
        uint32_t n_targets = 0;
        int rc = settings.single
                     ? twkb_twk_open_single(settings.in.c_str(), settings.n_threads > 0 ? settings.n_threads : 1, iv[0], settings.l_surrounding,
                                            settings.emulate_quirks ? 1 : 0, runs ? 1 : 0, &twk, &n_targets, errbuf, sizeof(errbuf))
                     : (runs ? twkb_twk_open_runs : twkb_twk_open_intervals)(
                           settings.in.c_str(), settings.n_threads > 0 ? settings.n_threads : 1, iv.empty() ? nullptr : iv.data(),
                           (int32_t)iv.size(), settings.emulate_quirks ? 1 : 0, &twk, errbuf, sizeof(errbuf));
        if (rc) return error(errbuf[0] ? errbuf : "Failed to open file: " + settings.in + "...");
        uint32_t n_samples = 0, n_variants = 0, n_blocks = 0;
        size_t stride = 0;
        int32_t any_missing = 0;
        twkb_twk_dims(twk, &n_samples, &n_variants, &stride, &any_missing, &n_blocks);
        const uint64_t* data = nullptr;
        const uint64_t* mask = nullptr;
        const twkb_variant* meta = nullptr;
        const uint8_t* run_bytes = nullptr;
        size_t n_run_bytes = 0;
        const twkb_run_desc* run_desc = nullptr;
        if (runs) twkb_twk_runs_view(twk, &run_bytes, &n_run_bytes, &run_desc, &meta);
        else twkb_twk_view(twk, &data, &mask, &meta);
        const uint32_t* file_blocks = nullptr;  // the file's block structure: window rules and -c chunks are defined on it
        uint32_t n_file_blocks = 0;
        twkb_twk_blocks(twk, &file_blocks, &n_file_blocks);
        log() << "Samples: " << pretty(n_samples) << "..." << std::endl;
        log() << pretty(n_variants) << " variants from " << pretty(n_blocks) << " blocks..." << std::endl;
        log("PARAMS") << settings.GetString() << std::endl;

        // output name: a ".two" suffix is forced (ld.cpp:589-598)
        std::string out = settings.out;
        {
            const size_t slash = out.find_last_of('/'), dot = out.find_last_of('.');
            const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
            std::string ext = has_ext ? out.substr(dot + 1) : "";
            for (char& ch : ext) ch = (char)std::tolower((unsigned char)ch);
            if (ext != "two") out = (has_ext ? out.substr(0, dot) : out) + ".two";
        }
        if (to_stdout) out = "-";
        log("WRITER") << (to_stdout ? std::string("Writing to stdout...") : "Opening " + out + "...") << std::endl;
        void* writer = nullptr;
        rc = settings.sorted_output
                 ? twkb_two_open_sorted(out.c_str(), twk, command_line.c_str(), settings.c_level, settings.n_threads > 0 ? settings.n_threads : 1,
                                        &writer, errbuf, sizeof(errbuf))
                 : twkb_two_open(out.c_str(), twk, command_line.c_str(), settings.c_level, settings.b_size, &writer, errbuf, sizeof(errbuf));
        if (rc) {
            twkb_twk_close(twk);
            return error(errbuf[0] ? errbuf : "Failed to open file: " + out + "...");
        }
        if (!settings.sorted_output) twkb_two_set_threads(writer, settings.n_threads > 0 ? settings.n_threads : 1);
        auto close_writer = [&]() { return settings.sorted_output ? twkb_two_close_sorted(writer) : twkb_two_close(writer); };

        const int n_dev = (int)settings.devices.size();
        log("THREAD") << "Spawning " << n_dev << " device context(s)..." << std::endl;
        Shared shared;
        shared.writer = writer;
        std::vector<std::string> errors(n_dev);
        std::vector<twkb_stats> st(n_dev);
        std::vector<int> rcs(n_dev, 0);
        const auto t0 = std::chrono::steady_clock::now();
        // Several DISTINCT devices: the contexts form an NCCL communicator and every device uploads (and
        // decodes) only its slice of the variant rows over its own PCIe link; the slices are exchanged over
        // NVLink (twkb_load_*_sliced). Contexts are created up front so that a bad device ordinal fails
        // before anybody waits in a collective.
        // -w on several devices: position shards. The dependency range of a window run is bounded, so
        // device k holds only its own blocks and their halo -- the matrix as a whole never has to fit one GPU and
        // nothing is exchanged between the devices.
        std::vector<uint32_t> own_begin(n_dev + 1, 0), halo_end(n_dev, 0);
        const bool shard = settings.position_shards && settings.window && !settings.single && n_dev > 1 && file_blocks &&
                           n_file_blocks >= (uint32_t)n_dev &&
                           twkb_plan_shards(file_blocks, n_file_blocks, meta, n_variants, settings.l_window, n_dev, own_begin.data(),
                                            halo_end.data()) == TWKB_OK;
        std::vector<void*> ctxs(n_dev, nullptr);
        for (int k = 0; k < n_dev; ++k) {
            twkb_settings cs;
            settings.ToC(&cs);
            cs.device = settings.devices[k];
            cs.part_index = shard ? 0 : k;
            cs.part_count = shard ? 1 : n_dev;
            cs.shard_blocks = shard ? (int32_t)(own_begin[k + 1] - own_begin[k]) : 0;
            cs.single_targets = (int32_t)n_targets;
            const int r = twkb_create(&cs, &ctxs[k]);
            if (r) {
                const std::string why = twkb_last_error(nullptr);
                for (void* c : ctxs) twkb_destroy(c);
                close_writer();
                twkb_twk_close(twk);
                return error("device " + std::to_string(settings.devices[k]) + ": " + why);
            }
        }
        if (shard)
            log("THREAD") << "Window mode on " << n_dev << " devices: position shards (own blocks + halo within " << settings.l_window
                          << " bases), no matrix exchange" << std::endl;
        bool distinct = true;
        for (int a = 0; a < n_dev; ++a)
            for (int b = a + 1; b < n_dev; ++b) distinct = distinct && settings.devices[a] != settings.devices[b];
        uint8_t uid[TWKB_COMM_ID_BYTES];
        const bool use_comm = !shard && n_dev > 1 && distinct && !std::getenv("TWKB_NO_NCCL") && twkb_comm_unique_id(uid) == TWKB_OK;
        if (n_dev > 1 && !shard)
            log("THREAD") << (use_comm ? "NCCL communicator over " + std::to_string(n_dev) + " devices: sliced upload + exchange over NVLink"
                                       : std::string("no communicator (repeated device or NCCL unavailable): every context uploads the whole matrix"))
                          << std::endl;
        auto shard_worker = [&](int k) {
            void* ctx = ctxs[k];
            const uint32_t v0 = file_blocks[own_begin[k]], v1 = file_blocks[halo_end[k]], nv = v1 - v0;
            int r = TWKB_OK;
            if (nv && own_begin[k + 1] > own_begin[k]) {
                if (runs) {  // the shard's run words: one contiguous byte range of the inflated blocks, descriptors re-based
                    std::vector<twkb_run_desc> d(run_desc + v0, run_desc + v1);
                    uint64_t lo = UINT64_MAX, hi = 0;
                    for (const twkb_run_desc& x : d) {
                        lo = std::min<uint64_t>(lo, x.offset);
                        hi = std::max<uint64_t>(hi, x.offset + (uint64_t)x.n_runs * x.width);
                    }
                    for (twkb_run_desc& x : d) x.offset -= lo;
                    r = twkb_load_runs(ctx, n_samples, nv, run_bytes + lo, (size_t)(hi - lo), d.data(), meta + v0);
                } else {
                    r = twkb_load_matrix(ctx, n_samples, nv, data + (size_t)v0 * stride, mask ? mask + (size_t)v0 * stride : nullptr, stride,
                                         meta + v0);
                }
                std::vector<uint32_t> lb;
                for (uint32_t b = own_begin[k]; b < halo_end[k]; ++b) lb.push_back(file_blocks[b] - v0);
                if (r == TWKB_OK) r = twkb_set_blocks(ctx, lb.data(), (uint32_t)lb.size());
                if (r == TWKB_OK) r = twkb_compute(ctx, &twk_ld::sink, &shared);
            }
            if (r) { errors[k] = twkb_last_error(ctx); rcs[k] = r; }
            else twkb_get_stats(ctx, &st[k]);
            twkb_destroy(ctx);
        };
        auto worker = [&](int k) {
            if (shard) return shard_worker(k);
            void* ctx = ctxs[k];
            int r = TWKB_OK;
            if (use_comm) r = twkb_comm_init(ctx, uid, k, n_dev);
            if (r == TWKB_OK) {
                if (use_comm) {
                    uint32_t b = 0, e = 0;
                    twkb_comm_slice(n_variants, k, n_dev, &b, &e);
                    r = runs ? twkb_load_runs_sliced(ctx, n_samples, n_variants, run_bytes, n_run_bytes, run_desc, meta)
                             : twkb_load_matrix_sliced(ctx, n_samples, n_variants, data + (size_t)b * stride,
                                                       mask ? mask + (size_t)b * stride : nullptr, stride, meta);
                } else {
                    r = runs ? twkb_load_runs(ctx, n_samples, n_variants, run_bytes, n_run_bytes, run_desc, meta)
                             : twkb_load_matrix(ctx, n_samples, n_variants, data, mask, stride, meta);
                }
            }
            if (r == TWKB_OK && file_blocks && n_file_blocks) r = twkb_set_blocks(ctx, file_blocks, n_file_blocks);
            if (r == TWKB_OK) r = settings.sorted_output ? twkb_compute_sorted(ctx, &twk_ld::sorted_sink, &shared) : twkb_compute(ctx, &twk_ld::sink, &shared);
            if (r) { errors[k] = twkb_last_error(ctx); rcs[k] = r; }
            else twkb_get_stats(ctx, &st[k]);
            twkb_destroy(ctx);
        };
        std::vector<std::thread> pool;
        for (int k = 1; k < n_dev; ++k) pool.emplace_back(worker, k);
        worker(0);
        for (std::thread& t : pool) t.join();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        bool ok = true;
        for (int k = 0; k < n_dev; ++k)
            if (rcs[k]) { error("device " + std::to_string(settings.devices[k]) + ": " + errors[k]); ok = false; }
        rc = close_writer();
        twkb_twk_close(twk);
        if (!ok) return false;
        if (rc) return error("Failed to write final block!");
        for (int k = 0; k < n_dev; ++k) {
            stats.pairs_visited += st[k].pairs_visited; stats.pairs_screened += st[k].pairs_screened;
            stats.records_out += st[k].records_out; stats.count_launches += st[k].count_launches;
            stats.stats_launches += st[k].stats_launches; stats.other_launches += st[k].other_launches;
            stats.sparse_launches += st[k].sparse_launches;
            stats.ms_count_kernel += st[k].ms_count_kernel; stats.ms_stats_kernel += st[k].ms_stats_kernel;
            stats.bytes_h2d += st[k].bytes_h2d; stats.bytes_d2h += st[k].bytes_d2h;
            stats.kernel_used = st[k].kernel_used;
        }
        stats.seconds_total = secs;
        // twk_ld_progress::PrintFinal, lib/ld/ld_progress.h:89-96 (genotypes = pairs x samples)
        log("PROGRESS") << "Finished in " << secs << "s. Variants: " << pretty(stats.pairs_visited)
                        << ", genotypes: " << pretty(stats.pairs_visited * n_samples) << ", output: " << pretty(stats.records_out) << std::endl;
        log("PROGRESS") << pretty((uint64_t)(stats.pairs_visited / (secs > 0 ? secs : 1e-9))) << " variants/s and "
                        << pretty((uint64_t)(stats.pairs_visited * (double)n_samples / (secs > 0 ? secs : 1e-9))) << " genotypes/s" << std::endl;
        log("PROGRESS") << "All done..." << std::endl;
        return true;
    }

    twk_ld_settings settings;
    twkb_stats stats{};            // aggregated over the device contexts of the last Compute()
    std::string command_line = "twkb_calc";  // recorded in the .two header (##tomahawk_calcCommand)

private:
    struct Shared { std::mutex mu; void* writer = nullptr; };
    static int sink(void* user, const uint8_t* recs, uint64_t n) {
        Shared* s = static_cast<Shared*>(user);
        std::lock_guard<std::mutex> g(s->mu);
        return twkb_two_add(s->writer, recs, n);
    }
    static int sorted_sink(void* user, const uint8_t* recs, uint64_t n) {
        return twkb_two_add_sorted(static_cast<Shared*>(user)->writer, recs, n);
    }
    bool error(const std::string& m) const {
        std::cerr << timestamp("ERROR") << m << std::endl;
        return false;
    }
    struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
    std::ostream& log(const char* sub = nullptr) {
        static NullBuf nb;
        static std::ostream null_stream(&nb);
        if (settings.silent) return null_stream;
        std::cerr << (sub ? timestamp("LOG", sub) : timestamp("LOG"));
        return std::cerr;
    }
};

}  // namespace twkb_host

#endif  // TWKB_LD_HPP_
