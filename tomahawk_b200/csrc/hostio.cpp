// hostio.cpp -- `.twk` reader and `.two` writer (host side; zstd on the host).
// Layouts follow the reference's serializers; citations inline.
#include "hostio.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace twkb {

namespace {

const char kTwkMagic[9] = {'T', 'O', 'M', 'A', 'H', 'A', 'W', 'K', 1};  // include/tomahawk.h:47-48
const char kTwoMagic[4] = {'T', 'W', 'O', 1};                           // include/tomahawk.h:50-51
const char kEof[] = "a4f54f39f5e251a6993796f48164ccf5";                 // first 32 chars, tomahawk.h:66-67
const uint64_t kIndexMarker = 1954702206512158641ull;                   // tomahawk.h:68

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    template <typename T>
    T get() {
        T v{};
        if ((size_t)(end - p) < sizeof(T)) { ok = false; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {  // u32 length + bytes, lib/buffer.cpp:410-421
        uint32_t n = get<uint32_t>();
        if (!ok || (size_t)(end - p) < n) { ok = false; return {}; }
        std::string s(reinterpret_cast<const char*>(p), n);
        p += n;
        return s;
    }
};

// Inflates one zstd frame whose size the file declares. The declared size is attacker-controlled
// (a corrupt header must not become a 2^62-byte allocation): it has to agree with the size recorded in
// the frame itself, or -- for frames written without one -- stay within zstd's maximum expansion.
bool zstd_inflate(const uint8_t* src, size_t n_cmp, size_t n_unc, std::vector<uint8_t>& dst, std::string& err) {
    const unsigned long long fcs = ZSTD_getFrameContentSize(src, n_cmp);
    const unsigned long long kUnknown = 0ULL - 1, kError = 0ULL - 2;  // ZSTD_CONTENTSIZE_UNKNOWN / _ERROR
    if (fcs == kError || (fcs != kUnknown && fcs != (unsigned long long)n_unc) ||
        (fcs == kUnknown && (unsigned long long)n_unc > (1ull << 20) + 32768ull * (unsigned long long)n_cmp)) {
        err = "zstd decompress failed: declared size does not match the frame";
        return false;
    }
    try {
        dst.resize(n_unc ? n_unc : 1);
    } catch (const std::bad_alloc&) {
        err = "out of memory inflating a block";
        return false;
    }
    const size_t r = ZSTD_decompress(dst.data(), n_unc, src, n_cmp);
    if (ZSTD_isError(r) || r != n_unc) {
        err = std::string("zstd decompress failed: ") + (ZSTD_isError(r) ? ZSTD_getErrorName(r) : "size mismatch");
        return false;
    }
    dst.resize(n_unc);
    return true;
}

// OR `pattern` (period 2 bits) into bits [start, end) of a row; start is even.
inline void or_range(uint64_t* row, uint64_t start, uint64_t end, uint64_t pattern) {
    if (pattern == 0 || start >= end) return;
    uint64_t w0 = start >> 6, w1 = (end - 1) >> 6;
    const uint64_t m0 = ~0ull << (start & 63);
    const uint64_t m1 = (end & 63) ? ((1ull << (end & 63)) - 1) : ~0ull;
    if (w0 == w1) { row[w0] |= pattern & m0 & m1; return; }
    row[w0] |= pattern & m0;
    for (uint64_t w = w0 + 1; w < w1; ++w) row[w] |= pattern;
    row[w1] |= pattern & m1;
}

struct BlockRef { uint64_t foff; uint32_t n, first_variant; int32_t rid; uint32_t minpos, maxpos; };

struct Contig { std::string name; int64_t n_bases; uint32_t idx; };  // idx: VcfContig::idx, the key blocks' rid refers to (include/header.h:115-148)
struct Ival { uint32_t start, stop; };

// Grammar of the reference's interval strings (include/tomahawk.h:57-59): a number is
// digits, an optional decimal part and an optional one-digit exponent; values go through
// atof and are truncated to u32 (lib/intervals.cpp:103-104).
bool is_name(const std::string& s) {
    if (s.empty()) return false;
    for (char ch : s)
        if (!(std::isalnum((unsigned char)ch) || ch == '-' || ch == '_')) return false;
    return true;
}
bool is_number(const std::string& s) {
    size_t i = 0;
    while (i < s.size() && std::isdigit((unsigned char)s[i])) ++i;
    if (i == 0) return false;
    if (i < s.size() && s[i] == '.') {
        const size_t j = ++i;
        while (i < s.size() && std::isdigit((unsigned char)s[i])) ++i;
        if (i == j) return false;
    }
    if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
        if (i + 2 != s.size() || !std::isdigit((unsigned char)s[i + 1])) return false;
        i += 2;
    }
    return i == s.size();
}

// twk_intervals::ParseIntervalString + Dedupe + Build (lib/intervals.cpp:35-136) and the block
// loading rule of twk_ld_impl::LoadTargetBlocks (lib/ld/ld.cpp:279-365): `calc -I` works at
// .twk BLOCK granularity -- every variant of every block that overlaps an interval takes part.
// Returns the index-entry numbers to load, in load order.
// One interval string -> (position of the contig in `contigs`, start, stop); grammar and value rules of
// twk_intervals::ParseIntervalString (lib/intervals.cpp:91-136).
int parse_interval_string(const std::string& s, const std::vector<Contig>& contigs, int& c, uint32_t& start, uint32_t& stop, std::string& err) {
    auto contig_of = [&](const std::string& name) -> int {
        for (size_t k = 0; k < contigs.size(); ++k)
            if (contigs[k].name == name) return (int)k;
        return -1;
    };
    const size_t colon = s.find(':');
    const std::string name = s.substr(0, colon);
    if (colon != std::string::npos && s.find(':', colon + 1) != std::string::npos) { err = "Illegal format: " + s; return TWKB_EINVAL; }
    if (!is_name(name)) { err = "Illegal interval: " + s; return TWKB_EINVAL; }
    c = contig_of(name);
    if (colon == std::string::npos) {  // contig only: [0, n_bases]
        if (c < 0) { err = "Contig does not exist in string " + s; return TWKB_EINVAL; }
        start = 0u;
        stop = (uint32_t)contigs[c].n_bases;
        return TWKB_OK;
    }
    const std::string rest = s.substr(colon + 1);
    // a '-' can only separate the two numbers (names were cut at the colon)
    const size_t dash = rest.find('-');
    if (dash == std::string::npos) {  // contig:pos -> [pos, pos + 1]
        if (!is_number(rest)) { err = "Illegal interval: " + s; return TWKB_EINVAL; }
        if (c < 0) { err = "Contig does not exist in string " + s; return TWKB_EINVAL; }
        start = (uint32_t)std::atof(rest.c_str());
        stop = start + 1;
    } else {
        const std::string a = rest.substr(0, dash), b = rest.substr(dash + 1);
        if (!is_number(a) || !is_number(b)) { err = "Illegal interval: " + s; return TWKB_EINVAL; }
        if (c < 0) { err = "Contig does not exist in string " + s; return TWKB_EINVAL; }
        start = (uint32_t)std::atof(a.c_str());
        stop = (uint32_t)std::atof(b.c_str());
    }
    return TWKB_OK;
}

int select_interval_blocks(const std::vector<std::string>& strings, const std::vector<Contig>& contigs,
                           const std::vector<BlockRef>& blocks, bool emulate_quirks, std::vector<uint32_t>& sel,
                           std::string& err) {
    // the reference keys its interval vectors by VcfContig::idx (lib/intervals.cpp:109) and walks them in idx order
    uint32_t max_idx = 0;
    for (const Contig& ct : contigs) max_idx = std::max(max_idx, ct.idx);
    if (max_idx > (1u << 24)) { err = "corrupt contig index"; return TWKB_EIO; }
    std::vector<std::vector<Ival>> ivecs((size_t)max_idx + 1);
    for (const std::string& s : strings) {
        int c = -1;
        uint32_t a = 0, b = 0;
        const int rc = parse_interval_string(s, contigs, c, a, b, err);
        if (rc) return rc;
        ivecs[contigs[c].idx].push_back({a, b});
    }
    std::vector<uint32_t> overlap;
    for (size_t c = 0; c < ivecs.size(); ++c) {
        std::vector<Ival>& v = ivecs[c];
        if (v.empty()) continue;
        std::stable_sort(v.begin(), v.end(), [](const Ival& x, const Ival& y) { return x.start != y.start ? x.start < y.start : x.stop < y.stop; });
        std::vector<Ival> merged{v[0]};
        for (size_t j = 1; j < v.size(); ++j) {  // Dedupe, lib/intervals.cpp:35-52
            if (v[j].start < merged.back().stop && v[j].stop >= merged.back().start) merged.back().stop = v[j].stop;
            else merged.push_back(v[j]);
        }
        for (const Ival& iv : merged)  // Index::FindOverlap, lib/index.cpp:125-134 (minpos/maxpos are 1-based)
            for (size_t b = 0; b < blocks.size(); ++b)
                if (blocks[b].rid == (int32_t)c && blocks[b].minpos <= iv.stop && blocks[b].maxpos >= iv.start) overlap.push_back((uint32_t)b);
    }
    if (overlap.empty()) { err = "Found no blocks overlapping the provided range(s)..."; return TWKB_EINVAL; }
    sel.clear();
    if (emulate_quirks) {
        // The reference seeks to the FIRST overlapping block and reads overlap.size() consecutive
        // blocks from there (ld.cpp:323-333), whatever the other entries were: identical to the
        // union for one interval, a run of neighbours for several disjoint ones.
        for (uint32_t k = 0; k < overlap.size(); ++k) {
            const uint64_t b = (uint64_t)overlap[0] + k;
            if (b >= blocks.size()) { err = "Failed to load block " + std::to_string(k) + "..."; return TWKB_EIO; }
            sel.push_back((uint32_t)b);
        }
    } else {
        sel = overlap;
        std::sort(sel.begin(), sel.end());
        sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
    }
    return TWKB_OK;
}

}  // namespace

// lib/twk_reader.cpp:49-125 (Open), :8-44 (NextBlock), lib/core.cpp:75-101 (twk1_t),
// include/core.h:195-215 (run words), lib/core.cpp:365-383 (bitvector + mask).
int read_twk(const std::string& path, int n_threads, TwkFile& out, std::string& err, const std::vector<std::string>* intervals,
             bool emulate_quirks, bool keep_runs, int32_t single_surrounding) {
    // the file is mapped, not copied: the worker threads fault its pages in as they inflate blocks
    struct Mapped {
        const uint8_t* p = nullptr;
        size_t n = 0;
        ~Mapped() { if (p) munmap(const_cast<uint8_t*>(p), n); }
        const uint8_t* data() const { return p; }
        size_t size() const { return n; }
    } file;
    {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) { err = "Failed to open \"" + path + "\"!"; return TWKB_EIO; }
        struct stat sb;
        if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { ::close(fd); err = "Failed to read \"" + path + "\""; return TWKB_EIO; }
        void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) { err = "Failed to read \"" + path + "\""; return TWKB_EIO; }
        file.p = static_cast<const uint8_t*>(m);
        file.n = (size_t)sb.st_size;
    }
    if (file.size() < 9 + 16 + 8 + 32 || std::memcmp(file.data(), kTwkMagic, 9) != 0) { err = "Failed to read MAGIC!"; return TWKB_EIO; }
    Cursor c{file.data() + 9, file.data() + file.size()};
    const uint64_t h_unc = c.get<uint64_t>(), h_cmp = c.get<uint64_t>();
    if (!c.ok || (uint64_t)(c.end - c.p) < h_cmp) { err = "truncated header"; return TWKB_EIO; }
    std::vector<uint8_t> hdr;
    std::vector<Contig> contigs;
    if (!zstd_inflate(c.p, h_cmp, h_unc, hdr, err)) return TWKB_EIO;
    {
        Cursor h{hdr.data(), hdr.data() + hdr.size()};
        out.fileformat = h.str();
        out.literals = h.str();
        const uint8_t* tail = h.p;
        out.n_samples = h.get<uint32_t>();
        for (uint32_t i = 0; i < out.n_samples && h.ok; ++i) h.str();
        out.n_contigs = h.get<uint32_t>();
        if (!h.ok) { err = "corrupt VcfHeader"; return TWKB_EIO; }
        out.header_tail.assign(reinterpret_cast<const char*>(tail), hdr.data() + hdr.size() - tail);
        for (uint32_t i = 0; i < out.n_contigs && h.ok; ++i) {  // VcfContig, include/header.h:115-128
            Contig ct;
            ct.idx = h.get<uint32_t>();
            ct.name = h.str();
            h.str();                                              // description
            ct.n_bases = h.get<int64_t>();
            const uint32_t n_extra = h.get<uint32_t>();
            for (uint32_t x = 0; x < n_extra && h.ok; ++x) { h.str(); h.str(); }
            contigs.push_back(ct);
        }
        if (!h.ok) contigs.clear();  // names are only needed for -I
        out.contig_n_bases.clear();
        for (const Contig& ct : contigs) out.contig_n_bases.push_back(ct.n_bases);
    }
    // footer: ... u64 offset_of_index, 32-byte EOF
    uint64_t idx_off = 0;
    std::memcpy(&idx_off, file.data() + file.size() - 32 - 8, 8);
    if (idx_off + 17 > file.size()) { err = "Failed to seek in file!"; return TWKB_EIO; }
    Cursor f{file.data() + idx_off, file.data() + file.size()};
    const uint8_t marker = f.get<uint8_t>();
    const uint64_t i_unc = f.get<uint64_t>(), i_cmp = f.get<uint64_t>();
    if (marker != 0 || !f.ok || (uint64_t)(f.end - f.p) < i_cmp) { err = "corrupt index footer"; return TWKB_EIO; }
    std::vector<uint8_t> idx;
    if (!zstd_inflate(f.p, i_cmp, i_unc, idx, err)) { err = "Failed to decompress index!"; return TWKB_EIO; }
    Cursor ix{idx.data(), idx.data() + idx.size()};
    if (ix.get<uint64_t>() != kIndexMarker) { err = "bad index marker"; return TWKB_EIO; }
    uint64_t n_ent = ix.get<uint64_t>();
    ix.get<uint64_t>();  // m
    ix.get<uint64_t>();  // m_ent
    if (!ix.ok || n_ent > (uint64_t)(ix.end - ix.p) / 40) { err = "corrupt index (entry count)"; return TWKB_EIO; }
    std::vector<BlockRef> blocks(n_ent);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_ent; ++i) {  // IndexEntry, lib/index.cpp:8-18
        const int32_t rid = ix.get<int32_t>();
        const uint32_t n = ix.get<uint32_t>();
        const uint32_t minpos = ix.get<uint32_t>(), maxpos = ix.get<uint32_t>();
        ix.get<uint32_t>(); ix.get<uint32_t>();
        const uint64_t foff = ix.get<uint64_t>();
        ix.get<uint64_t>();
        blocks[i] = {foff, n, (uint32_t)total, rid, minpos, maxpos};
        total += n;
    }
    if (!ix.ok || total == 0 || total > 0xffffffffull) { err = "No valid data available..."; return TWKB_EIO; }
    // scalc (twk_ld_impl::LoadTargetSingle, lib/ld/ld.cpp:123-255): ONE interval names the target site(s); the variants
    // within l_surrounding bases on either side are its neighbourhood. Matching is per VARIANT, 1-based, inclusive.
    const bool single = single_surrounding >= 0;
    int single_c = -1;
    uint32_t s_start = 0, s_stop = 0, s_left0 = 0, s_left1 = 0, s_right1 = 0;
    if (single) {
        if (!intervals || intervals->empty()) { err = "An interval has to be provided in single mode!"; return TWKB_EINVAL; }  // ld.cpp:689
        if (intervals->size() != 1) { err = "Only a single interval can be provided in single mode!"; return TWKB_EINVAL; }   // ld.cpp:694
        const int rc = parse_interval_string((*intervals)[0], contigs, single_c, s_start, s_stop, err);
        if (rc) return rc;
        // flanks (ld.cpp:147-156): [max(start - L, 0), max(start - 1, 0)] and [stop, stop + L]
        s_left0 = (uint32_t)std::max<int32_t>((int32_t)s_start - single_surrounding, 0);
        s_left1 = (uint32_t)std::max<int32_t>((int32_t)s_start - 1, 0);
        s_right1 = (uint32_t)((int32_t)s_stop + single_surrounding);
        const Ival iv[3] = {{s_left0, s_left1}, {s_start, s_stop}, {s_stop, s_right1}};
        std::vector<BlockRef> kept;
        total = 0;
        for (const BlockRef& b : blocks) {  // Index::FindOverlap, lib/index.cpp:125-134, distinct blocks in file order
            bool hit = false;
            for (const Ival& v : iv) hit = hit || (b.rid == (int32_t)contigs[single_c].idx && b.minpos <= v.stop && b.maxpos >= v.start);
            if (!hit) continue;
            kept.push_back(b);
            kept.back().first_variant = (uint32_t)total;
            total += b.n;
        }
        if (kept.empty()) { err = "Found no blocks overlapping the provided range(s)..."; return TWKB_EINVAL; }
        blocks.swap(kept);
    } else if (intervals && !intervals->empty()) {  // calc -I: keep the overlapping blocks only
        std::vector<uint32_t> sel;
        const int rc = select_interval_blocks(*intervals, contigs, blocks, emulate_quirks, sel, err);
        if (rc) return rc;
        std::vector<BlockRef> kept;
        total = 0;
        for (uint32_t b : sel) {
            kept.push_back(blocks[b]);
            kept.back().first_variant = (uint32_t)total;
            total += blocks[b].n;
        }
        blocks.swap(kept);
    }
    n_ent = blocks.size();
    // Plausibility of the (attacker-controlled) index BEFORE anything is sized by it: a block's variant count has to fit the
    // block's own uncompressed size -- 38 bytes of fixed twk1_t fields and at least one run word per variant behind the 12-byte
    // block header -- and that size has to be one a zstd frame of the block's compressed size can expand to. Without it a
    // corrupt count (up to 2^32 - 1 variants) became a 137 GB metadata allocation that the kernel's OOM killer answered.
    {
        uint64_t checked = 0;
        for (uint64_t b = 0; b < n_ent; ++b) {
            const BlockRef& br = blocks[b];
            if (br.foff > file.size() || file.size() - br.foff < 9) { err = "block offset beyond file"; return TWKB_EIO; }
            Cursor bc{file.data() + br.foff, file.data() + file.size()};
            if (bc.get<uint8_t>() != 1) { err = "bad block marker"; return TWKB_EIO; }
            const uint64_t unc = bc.get<uint32_t>(), cmp = bc.get<uint32_t>();
            if (cmp > (uint64_t)(bc.end - bc.p) || unc > (1ull << 20) + 32768ull * cmp) { err = "Failed to load block " + std::to_string(b); return TWKB_EIO; }
            if ((uint64_t)br.n * 39ull + 12ull > unc) { err = "index/block variant count mismatch"; return TWKB_EIO; }
            checked += br.n;
        }
        if (checked != total) { err = "corrupt index (variant count)"; return TWKB_EIO; }
    }
    out.n_blocks = (uint32_t)blocks.size();
    out.block_first.clear();
    for (const BlockRef& b : blocks) out.block_first.push_back(b.first_variant);
    out.block_first.push_back((uint32_t)total);
    out.n_variants = (uint32_t)total;
    const uint64_t H = 2ull * out.n_samples;
    out.stride = ((H + 63) / 64 + 1) / 2 * 2;
    out.runs_mode = keep_runs;
    out.meta.assign(total, twkb_variant{});
    // Runs mode: every block is inflated straight into its slot of out.raw (sizes from the block
    // headers), the variant headers are parsed in place and the run words stay where they are.
    std::vector<uint64_t> raw_off(n_ent + 1, 0);
    if (keep_runs) {
        for (uint64_t b = 0; b < n_ent; ++b) {
            if (blocks[b].foff + 9 > file.size()) { err = "block offset beyond file"; return TWKB_EIO; }
            uint32_t unc = 0;
            std::memcpy(&unc, file.data() + blocks[b].foff + 1, 4);
            raw_off[b + 1] = raw_off[b] + ((uint64_t)unc + 15) / 16 * 16;
        }
        if (!out.raw.alloc(raw_off[n_ent] + 16)) { err = "out of memory"; return TWKB_ENOMEM; }
        out.run_desc.assign(total, twkb_run_desc{});
    } else {
        out.data.assign((size_t)total * out.stride, 0);
    }
    std::vector<std::vector<uint64_t>> block_masks(n_ent);  // only for blocks that have missing data
    std::atomic<bool> any_miss_flag{false};
    std::atomic<uint64_t> next{0};
    std::atomic<bool> failed{false};
    std::string first_err;
    std::mutex err_mu;
    auto worker_body = [&]() {
        std::vector<uint8_t> raw;
        for (;;) {
            const uint64_t b = next.fetch_add(1);
            if (b >= n_ent || failed.load()) return;
            auto fail = [&](const std::string& m) {
                std::lock_guard<std::mutex> g(err_mu);
                if (!failed.exchange(true)) first_err = m;
            };
            const BlockRef& br = blocks[b];
            if (br.foff + 9 > file.size()) { fail("block offset beyond file"); return; }
            Cursor bc{file.data() + br.foff, file.data() + file.size()};
            if (bc.get<uint8_t>() != 1) { fail("bad block marker"); return; }
            const uint32_t unc = bc.get<uint32_t>(), cmp = bc.get<uint32_t>();
            std::string e;
            const uint8_t* raw_base = nullptr;
            if (keep_runs) {
                uint8_t* dst = out.raw.data() + raw_off[b];
                const size_t got = (uint64_t)(bc.end - bc.p) < cmp ? (size_t)-1 : ZSTD_decompress(dst, unc, bc.p, cmp);
                if (got == (size_t)-1 || ZSTD_isError(got) || got != unc) { fail("Failed to load block " + std::to_string(b)); return; }
                raw_base = dst;
            } else {
                if ((uint64_t)(bc.end - bc.p) < cmp || !zstd_inflate(bc.p, cmp, unc, raw, e)) { fail("Failed to load block " + std::to_string(b)); return; }
                raw_base = raw.data();
            }
            Cursor r{raw_base, raw_base + unc};
            const uint32_t n = r.get<uint32_t>();
            r.get<uint32_t>();  // m
            r.get<uint32_t>();  // rid
            if (n != br.n) { fail("index/block variant count mismatch"); return; }
            for (uint32_t v = 0; v < n; ++v) {
                const uint8_t pack = r.get<uint8_t>();
                r.get<uint8_t>();  // alleles
                twkb_variant& mv = out.meta[br.first_variant + v];
                mv.pos = r.get<uint32_t>();
                mv.ac = r.get<uint32_t>();
                mv.an = r.get<uint32_t>();
                mv.rid = r.get<uint32_t>();
                r.get<uint32_t>();  // n_het
                r.get<uint32_t>();  // n_hom
                mv.hwe = r.get<double>();
                const int ptype = pack >> 3;
                mv.gt_phase = (pack >> 1) & 1;
                mv.gt_missing = pack & 1;
                const uint32_t nw = r.get<uint32_t>();
                const uint32_t n_runs = nw >> 1, miss = nw & 1;
                if (!r.ok || (ptype != 1 && ptype != 2 && ptype != 4) || (size_t)(r.end - r.p) < (size_t)n_runs * ptype) {
                    fail("illegal gt primitive type / truncated runs");
                    return;
                }
                if (keep_runs) {  // the device walks the runs (decode.cuh); only locate them
                    twkb_run_desc& rd = out.run_desc[br.first_variant + v];
                    rd.offset = (uint64_t)(r.p - out.raw.data());
                    rd.n_runs = n_runs;
                    rd.width = (uint8_t)ptype;
                    rd.miss = (uint8_t)miss;
                    if (mv.gt_missing || miss || mv.an) any_miss_flag.store(true, std::memory_order_relaxed);
                    r.p += (size_t)n_runs * ptype;
                    continue;
                }
                uint64_t* row = out.data.data() + (size_t)(br.first_variant + v) * out.stride;
                uint64_t* mrow = nullptr;
                if (mv.gt_missing || miss) {
                    if (block_masks[b].empty()) block_masks[b].assign((size_t)n * out.stride, 0);
                    mrow = block_masks[b].data() + (size_t)v * out.stride;
                }
                const int lshift = 2 + 2 * miss, ashift = 1 + miss;
                const uint32_t amask = (1u << (1 + miss)) - 1;
                uint64_t cum = 0;
                for (uint32_t k = 0; k < n_runs; ++k) {
                    uint32_t word = 0;
                    std::memcpy(&word, r.p, ptype);
                    r.p += ptype;
                    const uint64_t len = word >> lshift;
                    const uint32_t a = (word >> ashift) & amask, bb = word & amask;
                    const uint64_t s = cum, e2 = cum + 2 * len;
                    if (e2 > H) { fail("run lengths exceed sample count"); return; }
                    or_range(row, s, e2, (a == 1 ? 0x5555555555555555ull : 0) | (bb == 1 ? 0xAAAAAAAAAAAAAAAAull : 0));
                    if (mrow && (a == 2 || bb == 2)) or_range(mrow, s, e2, ~0ull);
                    cum = e2;
                }
                if (cum != H) { fail("run lengths do not cover all samples"); return; }
            }
        }
    };
    auto worker = [&]() {  // an exception must never leave a std::thread (std::terminate would kill the caller's process)
        try {
            worker_body();
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> g(err_mu);
            if (!failed.exchange(true)) first_err = std::string("reader thread: ") + e.what();
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads, (int)n_ent));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    if (failed.load()) { err = first_err; return TWKB_EIO; }
    out.any_missing = any_miss_flag.load();
    for (auto& m : block_masks) if (!m.empty()) out.any_missing = true;
    if (single) {
        // per-variant classification (ld.cpp:185-222): pos + 1 against the three inclusive intervals; a variant that
        // matches the target interval is a target whatever else it matches. Resident order = [targets | neighbours],
        // both in file order: the reference's records name the target first (block 0 of CalculateSingle).
        std::vector<uint32_t> targets, others;
        for (uint32_t v = 0; v < (uint32_t)total; ++v) {
            const twkb_variant& mv = out.meta[v];
            if (mv.rid != contigs[single_c].idx) continue;
            const uint32_t p1 = mv.pos + 1;
            const bool t = p1 >= s_start && p1 <= s_stop;
            const bool l = p1 >= s_left0 && p1 <= s_left1, r = p1 >= s_stop && p1 <= s_right1;
            if (t) targets.push_back(v);
            else if (l && r) { err = "Corrupted intervals! (Too many matches)"; return TWKB_EINVAL; }
            else if (l || r) others.push_back(v);
        }
        if (targets.empty()) { err = "no data found for reference"; return TWKB_EINVAL; }   // ld.cpp:232
        // The reference collects the neighbours in blocks of 100 and only counts a block once it is FULL (ld.cpp:193-195,
        // n_blks = ldd2_n at :241): the last (n mod 100) neighbours in file order never take part, and fewer than 100
        // neighbours are "no surrounding variants". Reproduced with emulate_quirks, like the other quirks of the path.
        if (emulate_quirks) others.resize(others.size() / 100 * 100);
        if (others.empty()) { err = "no surrounding variants"; return TWKB_EINVAL; }        // ld.cpp:237
        std::vector<uint32_t> order(targets);
        order.insert(order.end(), others.begin(), others.end());
        const size_t n2 = order.size();
        std::vector<twkb_variant> meta2(n2);
        for (size_t k = 0; k < n2; ++k) meta2[k] = out.meta[order[k]];
        if (keep_runs) {
            std::vector<twkb_run_desc> rd2(n2);
            for (size_t k = 0; k < n2; ++k) rd2[k] = out.run_desc[order[k]];
            out.run_desc.swap(rd2);
        } else {
            std::vector<uint64_t> d2(n2 * out.stride);
            for (size_t k = 0; k < n2; ++k) std::memcpy(d2.data() + k * out.stride, out.data.data() + (size_t)order[k] * out.stride, out.stride * 8);
            out.data.swap(d2);
        }
        // masks of the host-unpack path are gathered below from the per-block buffers
        out.meta.swap(meta2);
        out.n_targets = (uint32_t)targets.size();
        out.n_variants = (uint32_t)n2;
        out.block_first.clear();  // the rows were re-ordered [targets | neighbours]: no block structure
        out.any_missing = false;
        if (!keep_runs) {
            bool any = false;
            for (const auto& m : block_masks) any = any || !m.empty();
            if (any) {
                std::vector<uint64_t> m2(n2 * out.stride, 0);
                for (size_t k = 0; k < n2; ++k) {
                    // block of the source variant
                    size_t b = 0;
                    while (b + 1 < blocks.size() && blocks[b + 1].first_variant <= order[k]) ++b;
                    if (!block_masks[b].empty())
                        std::memcpy(m2.data() + k * out.stride, block_masks[b].data() + (size_t)(order[k] - blocks[b].first_variant) * out.stride, out.stride * 8);
                }
                out.mask.swap(m2);
                out.any_missing = true;
            }
        }
        for (auto& mv : out.meta) if (mv.an || mv.gt_missing) out.any_missing = true;
        if (keep_runs)
            for (const auto& rd : out.run_desc) if (rd.miss) out.any_missing = true;
        if (out.any_missing && !keep_runs && out.mask.empty()) out.mask.assign(n2 * out.stride, 0);
        return TWKB_OK;
    }
    for (auto& mv : out.meta) if (mv.an || mv.gt_missing) out.any_missing = true;
    if (out.any_missing && !keep_runs) {
        out.mask.assign((size_t)total * out.stride, 0);
        for (uint64_t b = 0; b < n_ent; ++b)
            if (!block_masks[b].empty())
                std::memcpy(out.mask.data() + (size_t)blocks[b].first_variant * out.stride, block_masks[b].data(),
                            block_masks[b].size() * 8);
    }
    return TWKB_OK;
}

// ------------------------------------------------------------------ .two writer
TwoWriter::~TwoWriter() {
    stop_writer();
    if (fp_ && !to_stdout_) std::fclose(fp_);
}

// Every byte goes through here: the file offsets of the index come from this counter, not from ftell(), so the
// writer also works on a pipe ("-o -": the reference's default streams the .two to stdout, lib/ld/ld.cpp:585-588).
bool TwoWriter::emit(const void* p, size_t n) {
    if (n && std::fwrite(p, 1, n, fp_) != n) return false;
    pos_ += n;
    return true;
}

void TwoWriter::set_threads(int n) {
    threads_ = n < 1 ? 1 : n;
    if (threads_ > 1 && !writer_.joinable()) writer_ = std::thread(&TwoWriter::writer_loop, this);
}

// Writer thread: takes record batches off the queue in arrival order.
void TwoWriter::writer_loop() {
    for (;;) {
        std::vector<uint8_t> batch;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return;  // stop requested and everything written
            batch.swap(queue_.front());
            queue_.pop_front();
        }
        const int rc = add_sync(batch.data(), batch.size() / TWKB_RECORD_BYTES);
        {
            std::lock_guard<std::mutex> lk(mu_);
            queued_bytes_ -= batch.size();
            if (rc && !async_rc_) async_rc_ = rc;
        }
        cv_.notify_all();
    }
}

int TwoWriter::stop_writer() {
    if (!writer_.joinable()) return async_rc_;
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    writer_.join();
    return async_rc_;
}

int TwoWriter::add(const uint8_t* records, uint64_t n) {
    if (!writer_.joinable()) return add_sync(records, n);
    if (n == 0) return async_rc_;
    const size_t bytes = (size_t)n * TWKB_RECORD_BYTES;
    std::vector<uint8_t> copy(records, records + bytes);
    std::unique_lock<std::mutex> lk(mu_);
    // back-pressure: at most ~512 MB of records waiting for the writer
    cv_.wait(lk, [&] { return async_rc_ != 0 || queued_bytes_ < ((size_t)512 << 20); });
    if (async_rc_) return async_rc_;
    queued_bytes_ += bytes;
    queue_.emplace_back(std::move(copy));
    lk.unlock();
    cv_.notify_all();
    return TWKB_OK;
}

static void put_str(std::vector<uint8_t>& b, const std::string& s) {
    const uint32_t n = (uint32_t)s.size();
    b.insert(b.end(), reinterpret_cast<const uint8_t*>(&n), reinterpret_cast<const uint8_t*>(&n) + 4);
    b.insert(b.end(), s.begin(), s.end());
}

// include/writer.h:225-242 + lib/ld/ld.cpp:609-612
int TwoWriter::open(const std::string& path, const TwkFile& src, const std::string& command_line, int c_level, int b_size,
                    std::string& err) {
    to_stdout_ = path.empty() || path == "-";
    fp_ = to_stdout_ ? stdout : std::fopen(path.c_str(), "wb");
    if (!fp_) { err = "Failed to open file: " + path + "..."; return TWKB_EIO; }
    pos_ = 0;
    c_level_ = c_level;
    b_size_ = (uint32_t)std::max(2, b_size);
    n_contigs_ = src.n_contigs;
    char date[64];
    std::time_t now = std::time(nullptr);
    std::strftime(date, sizeof(date), "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    std::string literals = src.literals + "\n##tomahawk_calcVersion=b200-0.1.0\n##tomahawk_calcCommand=" + command_line +
                           "; Date=" + date + "\n";
    std::vector<uint8_t> hdr;
    put_str(hdr, src.fileformat);
    put_str(hdr, literals);
    hdr.insert(hdr.end(), src.header_tail.begin(), src.header_tail.end());
    std::vector<uint8_t> z(ZSTD_compressBound(hdr.size()));
    const size_t zn = ZSTD_compress(z.data(), z.size(), hdr.data(), hdr.size(), c_level_);
    if (ZSTD_isError(zn)) { err = "failed to compress"; return TWKB_EIO; }
    const uint64_t unc = hdr.size(), cmp = zn;
    if (!emit(kTwoMagic, 4) || !emit(&unc, 8) || !emit(&cmp, 8) || !emit(z.data(), zn)) { err = "Failed to write header!"; return TWKB_EIO; }
    fwd_.ent = IndexEntry{-1, -1, 0, 0, 0, 0, 0, 0, 0};
    rev_.ent = fwd_.ent;
    return TWKB_OK;
}

// Compresses every pending block (up to threads_ at a time; blocks are independent zstd frames)
// and writes them in queue order. One block on disk: u8 1, u32 unc, u32 cmp, payload
// (include/writer.h:70-87); the index entry gets its offsets here.
int TwoWriter::drain() {
    if (pending_.empty()) return TWKB_OK;
    const bool trace = std::getenv("TWKB_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    std::atomic<size_t> next{0};
    std::atomic<bool> bad{false};
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= pending_.size()) return;
            Pending& pb = pending_[k];
            pb.z.resize(ZSTD_compressBound(pb.raw.size()));
            pb.zn = ZSTD_compress(pb.z.data(), pb.z.size(), pb.raw.data(), pb.raw.size(), c_level_);
            if (ZSTD_isError(pb.zn)) bad.store(true);
        }
    };
    const int nt = (int)std::min<size_t>((size_t)threads_, pending_.size());
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (bad.load()) { err_ = "failed compression"; return TWKB_EIO; }
    const auto t1 = std::chrono::steady_clock::now();
    for (Pending& pb : pending_) {
        const uint8_t marker = 1;
        const uint32_t unc = (uint32_t)pb.raw.size(), cmp = (uint32_t)pb.zn;
        pb.ent.foff = pos_;
        if (!emit(&marker, 1) || !emit(&unc, 4) || !emit(&cmp, 4) || !emit(pb.z.data(), pb.zn)) {
            err_ = "write failed";
            return TWKB_EIO;
        }
        pb.ent.fend = pos_;
        pb.ent.b_cmp = cmp;
        index_.push_back(pb.ent);
    }
    if (trace)
        std::fprintf(stderr, "[twkb trace] writer drain: %zu blocks, compress %.1f ms (%d threads), write %.1f ms\n", pending_.size(),
                     std::chrono::duration<double, std::milli>(t1 - t0).count(), nt,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    pending_.clear();
    return TWKB_OK;
}

// lib/ld/ld_engine.cpp:1742-1802 (CompressFwd/CompressRev): the block is queued; drain() writes it.
int TwoWriter::flush_side(Side& s) {
    if (s.n == 0) return TWKB_OK;
    pending_.emplace_back();
    Pending& pb = pending_.back();
    pb.raw.resize(8 + s.buf.size());
    const uint32_t n = s.n, m = b_size_ + 100;  // twk1_two_block_t {n, m}, lib/core.cpp:626-631
    std::memcpy(pb.raw.data(), &n, 4);
    std::memcpy(pb.raw.data() + 4, &m, 4);
    std::memcpy(pb.raw.data() + 8, s.buf.data(), s.buf.size());
    pb.ent = s.ent;
    pb.ent.n = n;
    pb.ent.b_unc = TWKB_RECORD_BYTES * n + 8;
    n_written_ += n;
    s.buf.clear();
    s.n = 0;
    s.ent.rid = -1; s.ent.ridB = -1; s.ent.minpos = 0; s.ent.n = 0; s.ent.foff = 0; s.ent.fend = 0;
    if (pending_.size() >= (size_t)std::max(2, 4 * threads_)) return drain();
    return TWKB_OK;
}

// lib/ld/ld_engine.cpp:1268-1298: flush rule, index bookkeeping, forward + swapped copy. The
// reference tests every record; a run of records with the same (ridA, ridB) that fits the open
// block changes nothing but maxpos, so such runs are appended in bulk.
int TwoWriter::add_sync(const uint8_t* records, uint64_t n) {
    auto rids = [&](uint64_t r, int32_t& a, int32_t& b) {
        std::memcpy(&a, records + r * TWKB_RECORD_BYTES + 2, 4);
        std::memcpy(&b, records + r * TWKB_RECORD_BYTES + 6, 4);
    };
    uint64_t r = 0;
    while (r < n) {
        const uint8_t* rec = records + r * TWKB_RECORD_BYTES;
        int32_t ridA, ridB;
        uint32_t packA, packB;
        rids(r, ridA, ridB);
        std::memcpy(&packA, rec + 10, 4);
        std::memcpy(&packB, rec + 14, 4);
        if (fwd_.n == b_size_ || fwd_.ent.rid != ridA || rev_.ent.rid != ridB) {
            int rc = flush_side(fwd_);
            if (rc) return rc;
            rc = flush_side(rev_);
            if (rc) return rc;
            fwd_.ent.rid = ridA; fwd_.ent.ridB = ridB; fwd_.ent.minpos = packA >> 2; fwd_.ent.maxpos = packA >> 2;
            rev_.ent.rid = ridB; rev_.ent.ridB = ridA; rev_.ent.minpos = packB >> 2; rev_.ent.maxpos = packB >> 2;
        }
        if (fwd_.ent.ridB != ridB) fwd_.ent.ridB = -1;
        if (rev_.ent.ridB != ridA) rev_.ent.ridB = -1;
        // the run [r, e): same contig pair, fits the open block
        uint64_t e = r + 1;
        const uint64_t room = r + (b_size_ - fwd_.n);
        while (e < n && e < room) {
            int32_t a2, b2;
            rids(e, a2, b2);
            if (a2 != ridA || b2 != ridB) break;
            ++e;
        }
        const uint64_t cnt = e - r;
        const uint8_t* last = records + (e - 1) * TWKB_RECORD_BYTES;
        uint32_t lastA, lastB;
        std::memcpy(&lastA, last + 10, 4);
        std::memcpy(&lastB, last + 14, 4);
        fwd_.ent.maxpos = lastA >> 2;
        rev_.ent.maxpos = lastB >> 2;
        fwd_.buf.insert(fwd_.buf.end(), rec, rec + cnt * TWKB_RECORD_BYTES);
        fwd_.n += (uint32_t)cnt;
        // reverse copies: only (rid, pos) are swapped, counts and flags stay A-major (:1292-1298)
        const size_t o = rev_.buf.size();
        rev_.buf.insert(rev_.buf.end(), rec, rec + cnt * TWKB_RECORD_BYTES);
        for (uint64_t k = 0; k < cnt; ++k) {
            uint8_t* d = rev_.buf.data() + o + k * TWKB_RECORD_BYTES;
            const uint8_t* src = rec + k * TWKB_RECORD_BYTES;
            std::memcpy(d + 2, src + 6, 4);
            std::memcpy(d + 6, src + 2, 4);
            std::memcpy(d + 10, src + 14, 4);
            std::memcpy(d + 14, src + 10, 4);
        }
        rev_.n += (uint32_t)cnt;
        r = e;
    }
    return TWKB_OK;
}

// include/writer.h:293-313 + lib/index.cpp:242-251
int TwoWriter::finish() {
    if (!fp_) return TWKB_EIO;
    int rc = stop_writer();
    if (rc) return rc;
    rc = flush_side(fwd_);
    if (rc) return rc;
    rc = flush_side(rev_);
    if (rc) return rc;
    rc = drain();
    if (rc) return rc;
    std::vector<uint8_t> idx;
    auto put = [&](const void* p, size_t n) { idx.insert(idx.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    const uint8_t state = 0;  // TWK_IDX_UNSORTED
    uint64_t n = index_.size(), m = 500;
    while (m < n) m *= 2;
    const uint64_t m_ent = n_contigs_;
    put(&kIndexMarker, 8); put(&state, 1); put(&n, 8); put(&m, 8); put(&m_ent, 8);
    for (const IndexEntry& e : index_) {  // IndexEntryOutput, lib/index.cpp:41-52
        put(&e.rid, 4); put(&e.n, 4); put(&e.minpos, 4); put(&e.maxpos, 4); put(&e.b_unc, 4); put(&e.b_cmp, 4);
        put(&e.foff, 8); put(&e.fend, 8); put(&e.ridB, 4);
    }
    for (uint64_t c = 0; c < m_ent; ++c) {  // untouched IndexEntryEntry, lib/index.cpp:90-99
        const int32_t rid = 0; const uint32_t z32 = 0; const uint64_t z64 = 0;
        put(&rid, 4); put(&z32, 4); put(&z32, 4); put(&z32, 4); put(&z64, 8); put(&z64, 8); put(&z64, 8);
    }
    std::vector<uint8_t> z(ZSTD_compressBound(idx.size()));
    const size_t zn = ZSTD_compress(z.data(), z.size(), idx.data(), idx.size(), c_level_);
    if (ZSTD_isError(zn)) { err_ = "failed compression"; return TWKB_EIO; }
    const uint64_t off = pos_, unc = idx.size(), cmp = zn;
    const uint8_t marker = 0;
    bool ok = emit(&marker, 1) && emit(&unc, 8) && emit(&cmp, 8) && emit(z.data(), zn) && emit(&off, 8) && emit(kEof, 32);
    ok = ok && std::fflush(fp_) == 0 && !std::ferror(fp_);
    if (!to_stdout_) ok = (std::fclose(fp_) == 0) && ok;
    fp_ = nullptr;
    if (!ok) { err_ = "Failed to write final block!"; return TWKB_EIO; }
    return TWKB_OK;
}

// ------------------------------------------------------------------ .two sorter
namespace {

struct TwoIndexEntry {  // IndexEntryOutput, lib/index.cpp:41-52
    int32_t rid;
    uint32_t n, minpos, maxpos, b_unc, b_cmp;
    uint64_t foff, fend;
    int32_t ridB;
};
struct TwoMetaEntry {  // IndexEntryEntry, lib/index.cpp:90-99
    int32_t rid = 0;
    uint32_t n = 0, minpos = 0, maxpos = 0;
    uint64_t foff = 0, fend = 0, nn = 0;
};

inline void get_rec_key(const uint8_t* rec, int32_t& ridA, int32_t& ridB, uint32_t& posA, uint32_t& posB) {
    std::memcpy(&ridA, rec + 2, 4);
    std::memcpy(&ridB, rec + 6, 4);
    std::memcpy(&posA, rec + 10, 4);
    std::memcpy(&posB, rec + 14, 4);
}

}  // namespace

namespace {

struct SortKey { uint64_t hi, lo; uint64_t idx; };  // (ridA, ridB | posA, posB | original record number)
inline bool key_less(const SortKey& a, const SortKey& b) { return a.hi != b.hi ? a.hi < b.hi : (a.lo != b.lo ? a.lo < b.lo : a.idx < b.idx); }
inline SortKey key_of(const uint8_t* rec, uint64_t idx) {
    int32_t ridA, ridB; uint32_t pA, pB;
    get_rec_key(rec, ridA, ridB, pA, pB);
    // the reference compares rid as int32 (twk1_two_t::operator<, lib/core.cpp:458-468): bias to keep the order under unsigned comparison
    return SortKey{((uint64_t)((uint32_t)ridA ^ 0x80000000u) << 32) | ((uint32_t)ridB ^ 0x80000000u), ((uint64_t)pA << 32) | pB, idx};
}

// A sorted run: either resident (records + sorted keys in memory) or spilled to a temporary file as zstd chunks of
// <= kRunChunk records ([u32 n][u32 n_cmp][frame]) that are read back one chunk at a time by the merge.
constexpr uint32_t kRunChunk = 16384;
struct SortedRun {
    // resident
    const uint8_t* recs = nullptr;
    const SortKey* keys = nullptr;
    uint64_t n = 0, next = 0, base_idx = 0;
    // spilled
    FILE* fp = nullptr;
    std::string path;
    std::vector<uint8_t> buf, zbuf;
    uint32_t buf_n = 0, buf_next = 0;
    uint64_t emitted = 0;

    bool fill(std::string& err) {  // spilled: load the next chunk
        uint32_t hdr[2];
        if (std::fread(hdr, 4, 2, fp) != 2) { err = "failed to read a temporary run"; return false; }
        zbuf.resize(hdr[1]);
        buf.resize((size_t)hdr[0] * TWKB_RECORD_BYTES);
        if (hdr[0] == 0 || hdr[0] > kRunChunk || std::fread(zbuf.data(), 1, hdr[1], fp) != hdr[1]) { err = "failed to read a temporary run"; return false; }
        const size_t r = ZSTD_decompress(buf.data(), buf.size(), zbuf.data(), zbuf.size());
        if (ZSTD_isError(r) || r != buf.size()) { err = "corrupt temporary run"; return false; }
        buf_n = hdr[0];
        buf_next = 0;
        return true;
    }
    bool empty() const { return fp ? emitted == n : next == n; }
    const uint8_t* head() const { return fp ? buf.data() + (size_t)buf_next * TWKB_RECORD_BYTES : recs + (size_t)(keys[next].idx - base_idx) * TWKB_RECORD_BYTES; }
    bool pop(std::string& err) {
        if (fp) {
            ++emitted;
            if (++buf_next == buf_n && emitted < n) return fill(err);
            return true;
        }
        ++next;
        return true;
    }
};

// Sorted output: blocks of <= 10,000 records (twk_two_writer_t::n_blk_lim, include/writer.h:164), cut at every change of
// ridA, compressed in parallel batches, sorted-state index with per-contig entries.
struct SortedWriter {
    FILE* fp = nullptr;
    uint64_t pos = 0;
    int c_level = 1, n_threads = 1;
    uint64_t n_contigs = 0;
    struct Blk { std::vector<uint8_t> raw, z; size_t zn = 0; TwoIndexEntry ent{}; bool uniform = true; int32_t ridB0 = 0; uint32_t pA_first = 0, pA_last = 0; };
    std::vector<Blk> batch;
    std::vector<TwoIndexEntry> index;
    std::vector<TwoMetaEntry> meta;
    Blk cur;
    uint32_t cur_n = 0;
    int32_t cur_rid = 0;
    std::string err;
    static constexpr uint32_t blk_lim = 10000;

    bool emit(const void* p, size_t n) {
        if (n && std::fwrite(p, 1, n, fp) != n) { err = "write failed (disk full?)"; return false; }
        pos += n;
        return true;
    }
    void close_block() {
        if (!cur_n) return;
        const uint32_t m = blk_lim;
        std::memcpy(cur.raw.data(), &cur_n, 4);
        std::memcpy(cur.raw.data() + 4, &m, 4);
        cur.ent.rid = cur_rid;
        cur.ent.ridB = cur.uniform ? cur.ridB0 : -1;
        cur.ent.n = cur_n;
        // include/writer.h:363-374: minpos / maxpos are Apos (the position, flag bits dropped) of the first / last record
        cur.ent.minpos = cur.pA_first >> 2;
        cur.ent.maxpos = cur.pA_last >> 2;
        batch.push_back(std::move(cur));
        cur = Blk{};
        cur_n = 0;
    }
    bool add(const uint8_t* rec) {
        int32_t ra, rb; uint32_t pa, pb;
        get_rec_key(rec, ra, rb, pa, pb);
        if (cur_n && (ra != cur_rid || cur_n == blk_lim)) {
            close_block();
            if (batch.size() >= (size_t)std::max(8, 4 * n_threads) && !flush()) return false;
        }
        if (!cur_n) {
            cur.raw.resize(8);
            cur.raw.reserve(8 + (size_t)blk_lim * TWKB_RECORD_BYTES);
            cur_rid = ra; cur.ridB0 = rb; cur.pA_first = pa; cur.uniform = true;
        } else if (rb != cur.ridB0) cur.uniform = false;
        cur.pA_last = pa;
        cur.raw.insert(cur.raw.end(), rec, rec + TWKB_RECORD_BYTES);
        ++cur_n;
        return true;
    }
    bool flush() {  // compress the closed blocks in parallel, then write them in order
        if (batch.empty()) return true;
        std::atomic<size_t> next{0};
        std::atomic<bool> bad{false};
        auto work = [&]() {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= batch.size()) return;
                Blk& b = batch[k];
                b.z.resize(ZSTD_compressBound(b.raw.size()));
                b.zn = ZSTD_compress(b.z.data(), b.z.size(), b.raw.data(), b.raw.size(), c_level);
                if (ZSTD_isError(b.zn)) { bad.store(true); return; }
            }
        };
        std::vector<std::thread> pool;
        const int nt = (int)std::min<size_t>((size_t)n_threads, batch.size());
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        if (bad.load()) { err = "failed compression"; return false; }
        for (Blk& b : batch) {
            const uint8_t mk = 1;
            b.ent.b_unc = (uint32_t)b.raw.size();
            b.ent.b_cmp = (uint32_t)b.zn;
            b.ent.foff = pos;
            if (!emit(&mk, 1) || !emit(&b.ent.b_unc, 4) || !emit(&b.ent.b_cmp, 4) || !emit(b.z.data(), b.zn)) return false;
            b.ent.fend = pos;
            if (b.ent.rid < 0 || (uint64_t)b.ent.rid >= n_contigs) { err = "record with a contig id outside the header"; return false; }
            TwoMetaEntry& me = meta[b.ent.rid];  // IndexEntryEntry::operator+=, lib/index.cpp:70-88
            if (me.n == 0) { me.minpos = b.ent.minpos; me.foff = b.ent.foff; me.rid = b.ent.rid; }
            me.n += b.ent.n;
            me.maxpos = b.ent.maxpos;
            me.fend = b.ent.fend;
            ++me.nn;
            index.push_back(b.ent);
        }
        batch.clear();
        return true;
    }
};

// Sorted-state index (TWK_IDX_SORTED, include/index.h:105; block entries + per-contig entries) and the EOF marker; closes the file.
static bool write_sorted_tail(SortedWriter& w, uint64_t n_contigs, int c_level) {
    std::vector<uint8_t> ob_idx;
    auto put = [&](const void* p, size_t n) { ob_idx.insert(ob_idx.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    const uint8_t state = 2;
    uint64_t n_out = w.index.size(), m_out = 500;
    while (m_out < n_out) m_out *= 2;
    put(&kIndexMarker, 8); put(&state, 1); put(&n_out, 8); put(&m_out, 8); put(&n_contigs, 8);
    for (const TwoIndexEntry& e : w.index) {
        put(&e.rid, 4); put(&e.n, 4); put(&e.minpos, 4); put(&e.maxpos, 4); put(&e.b_unc, 4); put(&e.b_cmp, 4);
        put(&e.foff, 8); put(&e.fend, 8); put(&e.ridB, 4);
    }
    for (const TwoMetaEntry& me : w.meta) {
        put(&me.rid, 4); put(&me.n, 4); put(&me.minpos, 4); put(&me.maxpos, 4); put(&me.foff, 8); put(&me.fend, 8); put(&me.nn, 8);
    }
    std::vector<uint8_t> z(ZSTD_compressBound(ob_idx.size()));
    const size_t zn = ZSTD_compress(z.data(), z.size(), ob_idx.data(), ob_idx.size(), c_level);
    bool ok = !ZSTD_isError(zn);
    const uint64_t off = w.pos, unc = ob_idx.size(), cmp = zn;
    const uint8_t mk0 = 0;
    ok = ok && w.emit(&mk0, 1) && w.emit(&unc, 8) && w.emit(&cmp, 8) && w.emit(z.data(), zn) && w.emit(&off, 8) && w.emit(kEof, 32);
    ok = ok && std::fflush(w.fp) == 0 && !std::ferror(w.fp);
    ok = (std::fclose(w.fp) == 0) && ok;
    w.fp = nullptr;
    return ok;
}

}  // namespace

// `tomahawk sort` (two_reader::Sort, lib/two_reader.cpp:162-420). Bounded memory like the reference's: the input blocks are
// taken in runs that fit `memory_limit` bytes (records + keys); a run is inflated and sorted in parallel; if everything is
// one run it is merged straight from memory, otherwise every run is spilled to "<out>_<pid>_<k>.tmp" (zstd chunks) and the
// runs are merged k-way with a heap. Ties (identical ridA, ridB, posA, posB) keep the input order in both arrangements,
// so the output does not depend on the limit.
int sort_two(const std::string& in, const std::string& out_path, int c_level, int n_threads, std::string& err, uint64_t* n_records,
             uint64_t memory_limit) {
    n_threads = std::max(1, n_threads);
    const int fd = ::open(in.c_str(), O_RDONLY);
    if (fd < 0) { err = "Failed to open \"" + in + "\"..."; return TWKB_EIO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < (off_t)(4 + 16 + 17 + 8 + 32)) { ::close(fd); err = "Failed to open \"" + in + "\"..."; return TWKB_EIO; }
    const size_t fsz = (size_t)sb.st_size;
    void* mp = mmap(nullptr, fsz, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (mp == MAP_FAILED) { err = "Failed to open \"" + in + "\"..."; return TWKB_EIO; }
    struct Unmap { void* p; size_t n; ~Unmap() { munmap(p, n); } } unmap{mp, fsz};
    const uint8_t* file = static_cast<const uint8_t*>(mp);
    if (std::memcmp(file, kTwoMagic, 4) != 0) { err = "Failed to read MAGIC!"; return TWKB_EIO; }
    Cursor c{file + 4, file + fsz};
    const uint64_t h_unc = c.get<uint64_t>(), h_cmp = c.get<uint64_t>();
    if (!c.ok || (uint64_t)(c.end - c.p) < h_cmp) { err = "truncated header"; return TWKB_EIO; }
    std::vector<uint8_t> hdr;
    if (!zstd_inflate(c.p, h_cmp, h_unc, hdr, err)) return TWKB_EIO;
    uint64_t idx_off = 0;
    std::memcpy(&idx_off, file + fsz - 32 - 8, 8);
    if (idx_off + 17 > fsz) { err = "Failed to seek in file!"; return TWKB_EIO; }
    Cursor f{file + idx_off, file + fsz};
    const uint8_t marker = f.get<uint8_t>();
    const uint64_t i_unc = f.get<uint64_t>(), i_cmp = f.get<uint64_t>();
    if (marker != 0 || !f.ok || (uint64_t)(f.end - f.p) < i_cmp) { err = "corrupt index footer"; return TWKB_EIO; }
    std::vector<uint8_t> idx;
    if (!zstd_inflate(f.p, i_cmp, i_unc, idx, err)) { err = "Failed to decompress index!"; return TWKB_EIO; }
    Cursor ix{idx.data(), idx.data() + idx.size()};
    if (ix.get<uint64_t>() != kIndexMarker) { err = "bad index marker"; return TWKB_EIO; }
    ix.get<uint8_t>();  // state of the input: any
    const uint64_t n_ent = ix.get<uint64_t>();
    ix.get<uint64_t>();  // m
    const uint64_t n_contigs = ix.get<uint64_t>();  // m_ent
    if (!ix.ok || n_ent > (uint64_t)(ix.end - ix.p) / 44 || n_contigs > (1ull << 24)) { err = "corrupt index (entry count)"; return TWKB_EIO; }
    std::vector<TwoIndexEntry> ents(n_ent);
    std::vector<uint64_t> first(n_ent + 1, 0);
    for (uint64_t i = 0; i < n_ent; ++i) {
        TwoIndexEntry& e = ents[i];
        e.rid = ix.get<int32_t>(); e.n = ix.get<uint32_t>(); e.minpos = ix.get<uint32_t>(); e.maxpos = ix.get<uint32_t>();
        e.b_unc = ix.get<uint32_t>(); e.b_cmp = ix.get<uint32_t>(); e.foff = ix.get<uint64_t>(); e.fend = ix.get<uint64_t>();
        e.ridB = ix.get<int32_t>();
        first[i + 1] = first[i] + e.n;
    }
    if (!ix.ok) { err = "corrupt index"; return TWKB_EIO; }
    const uint64_t total = first[n_ent];
    if (total == 0) { err = "Cannot sort empty file..."; return TWKB_EINVAL; }  // two_reader.cpp:191-194
    if (n_records) *n_records = total;

    // ---- runs of input blocks that fit the memory budget (records + keys)
    const uint64_t per_rec = TWKB_RECORD_BYTES + sizeof(SortKey);
    if (memory_limit == 0) memory_limit = ~0ull;
    std::vector<std::pair<uint64_t, uint64_t>> run_blocks;  // [first block, last block)
    {
        uint64_t b0 = 0, acc = 0;
        for (uint64_t b = 0; b < n_ent; ++b) {
            const uint64_t need = (uint64_t)ents[b].n * per_rec;
            if (acc && acc + need > memory_limit) { run_blocks.push_back({b0, b}); b0 = b; acc = 0; }
            acc += need;
        }
        run_blocks.push_back({b0, n_ent});
    }
    const bool external = run_blocks.size() > 1;

    // ---- output file: header (+ sort provenance lines, two_reader.cpp:346-349)
    std::string out = out_path;
    {
        const size_t slash = out.find_last_of('/'), dot = out.find_last_of('.');
        const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
        std::string ext = has_ext ? out.substr(dot + 1) : "";
        for (auto& ch : ext) ch = (char)std::tolower((unsigned char)ch);
        if (ext != "two") out += ".two";  // two_reader.cpp:330-333
    }
    SortedWriter w;
    w.c_level = c_level;
    w.n_threads = n_threads;
    w.n_contigs = n_contigs;
    w.meta.resize((size_t)n_contigs);
    std::vector<std::string> temp_files;
    auto cleanup = [&](bool remove_out) {
        if (w.fp) { std::fclose(w.fp); w.fp = nullptr; }
        for (const std::string& t : temp_files) ::unlink(t.c_str());
        if (remove_out) ::unlink(out.c_str());  // never leave a truncated .two behind
    };
    w.fp = std::fopen(out.c_str(), "wb");
    if (!w.fp) { err = "Failed top open \"" + out + "\"..."; return TWKB_EIO; }
    {
        Cursor h{hdr.data(), hdr.data() + hdr.size()};
        const std::string fileformat = h.str();
        std::string literals = h.str();
        if (!h.ok) { cleanup(true); err = "corrupt VcfHeader"; return TWKB_EIO; }
        char date[64];
        std::time_t now = std::time(nullptr);
        std::strftime(date, sizeof(date), "%Y-%m-%d %H:%M:%S", std::localtime(&now));
        literals += "\n##tomahawk_sortVersion=b200-0.1.0\n##tomahawk_sortCommand=tomahawk_b200 sort -i " + in + " -o " + out_path + "; Date=" + date + "\n";
        std::vector<uint8_t> nh;
        put_str(nh, fileformat);
        put_str(nh, literals);
        nh.insert(nh.end(), h.p, h.end);
        std::vector<uint8_t> z(ZSTD_compressBound(nh.size()));
        const size_t zn = ZSTD_compress(z.data(), z.size(), nh.data(), nh.size(), c_level);
        if (ZSTD_isError(zn)) { cleanup(true); err = "failed to compress"; return TWKB_EIO; }
        const uint64_t unc = nh.size(), cmp = zn;
        if (!w.emit(kTwoMagic, 4) || !w.emit(&unc, 8) || !w.emit(&cmp, 8) || !w.emit(z.data(), zn)) { cleanup(true); err = w.err; return TWKB_EIO; }
    }

    // ---- build the runs
    ByteBuf recs;
    std::vector<SortKey> keys;
    std::vector<SortedRun> runs(run_blocks.size());
    for (size_t k = 0; k < run_blocks.size(); ++k) {
        const uint64_t b_lo = run_blocks[k].first, b_hi = run_blocks[k].second;
        const uint64_t r_lo = first[b_lo], n_run = first[b_hi] - r_lo;
        if (!recs.alloc((size_t)n_run * TWKB_RECORD_BYTES)) { cleanup(true); err = "out of memory"; return TWKB_ENOMEM; }
        try {
            keys.resize((size_t)n_run);
        } catch (const std::bad_alloc&) { cleanup(true); err = "out of memory"; return TWKB_ENOMEM; }
        {   // inflate the run's blocks (parallel over blocks)
            std::atomic<uint64_t> next{b_lo};
            std::atomic<bool> failed{false};
            auto worker = [&]() {
                std::vector<uint8_t> raw;
                std::string e;
                try {
                    for (;;) {
                        const uint64_t b = next.fetch_add(1);
                        if (b >= b_hi || failed.load()) return;
                        const TwoIndexEntry& en = ents[b];
                        if (en.foff + 9 > fsz) { failed.store(true); return; }
                        Cursor bc{file + en.foff, file + fsz};
                        if (bc.get<uint8_t>() != 1) { failed.store(true); return; }
                        const uint32_t unc = bc.get<uint32_t>(), cmp = bc.get<uint32_t>();
                        if ((uint64_t)(bc.end - bc.p) < cmp || !zstd_inflate(bc.p, cmp, unc, raw, e)) { failed.store(true); return; }
                        uint32_t n = 0;
                        if (raw.size() >= 8) std::memcpy(&n, raw.data(), 4);
                        if (n != en.n || raw.size() < 8 + (size_t)n * TWKB_RECORD_BYTES) { failed.store(true); return; }
                        uint8_t* dst = recs.data() + (first[b] - r_lo) * TWKB_RECORD_BYTES;
                        std::memcpy(dst, raw.data() + 8, (size_t)n * TWKB_RECORD_BYTES);
                        for (uint32_t r = 0; r < n; ++r) keys[first[b] - r_lo + r] = key_of(dst + (size_t)r * TWKB_RECORD_BYTES, first[b] + r);
                    }
                } catch (...) { failed.store(true); }
            };
            std::vector<std::thread> pool;
            const int nt = (int)std::min<uint64_t>((uint64_t)n_threads, b_hi - b_lo);
            for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
            worker();
            for (auto& th : pool) th.join();
            if (failed.load()) { cleanup(true); err = "Failed to load a block of \"" + in + "\""; return TWKB_EIO; }
        }
        {   // sorted sub-ranges per thread, then pairwise merges
            const int nt = (int)std::min<uint64_t>((uint64_t)n_threads, std::max<uint64_t>(1, n_run / 65536));
            std::vector<size_t> cut(nt + 1);
            for (int t = 0; t <= nt; ++t) cut[t] = (size_t)(n_run * (uint64_t)t / nt);
            std::vector<std::thread> pool;
            for (int t = 1; t < nt; ++t) pool.emplace_back([&, t]() { std::sort(keys.begin() + cut[t], keys.begin() + cut[t + 1], key_less); });
            std::sort(keys.begin() + cut[0], keys.begin() + cut[1], key_less);
            for (auto& th : pool) th.join();
            for (int width = 1; width < nt; width *= 2)
                for (int t = 0; t + width < nt; t += 2 * width)
                    std::inplace_merge(keys.begin() + cut[t], keys.begin() + cut[t + width], keys.begin() + cut[std::min(t + 2 * width, nt)], key_less);
        }
        SortedRun& run = runs[k];
        run.n = n_run;
        run.base_idx = r_lo;
        if (!external) {
            run.recs = recs.data();
            run.keys = keys.data();
            break;
        }
        // spill: zstd chunks in sorted order
        run.path = out + "_" + std::to_string((long)getpid()) + "_" + std::to_string(k) + ".tmp";
        FILE* tf = std::fopen(run.path.c_str(), "wb");
        if (!tf) { cleanup(true); err = "Failed to open temporary file \"" + run.path + "\""; return TWKB_EIO; }
        temp_files.push_back(run.path);
        std::vector<uint8_t> chunk, z;
        bool ok = true;
        for (uint64_t r = 0; r < n_run && ok; r += kRunChunk) {
            const uint32_t n = (uint32_t)std::min<uint64_t>(kRunChunk, n_run - r);
            chunk.resize((size_t)n * TWKB_RECORD_BYTES);
            for (uint32_t q = 0; q < n; ++q)
                std::memcpy(chunk.data() + (size_t)q * TWKB_RECORD_BYTES, recs.data() + (size_t)(keys[r + q].idx - r_lo) * TWKB_RECORD_BYTES, TWKB_RECORD_BYTES);
            z.resize(ZSTD_compressBound(chunk.size()));
            const size_t zn = ZSTD_compress(z.data(), z.size(), chunk.data(), chunk.size(), 1);
            const uint32_t hdr2[2] = {n, (uint32_t)zn};
            ok = !ZSTD_isError(zn) && std::fwrite(hdr2, 4, 2, tf) == 2 && std::fwrite(z.data(), 1, zn, tf) == zn;
        }
        ok = (std::fclose(tf) == 0) && ok;
        if (!ok) { cleanup(true); err = "Failed to write temporary file \"" + run.path + "\" (disk full?)"; return TWKB_EIO; }
    }
    if (external) {
        recs.release();
        std::vector<SortKey>().swap(keys);
        for (SortedRun& run : runs) {
            run.fp = std::fopen(run.path.c_str(), "rb");
            if (!run.fp || !run.fill(err)) { for (SortedRun& r2 : runs) if (r2.fp) std::fclose(r2.fp); cleanup(true); if (err.empty()) err = "Failed to reopen a temporary run"; return TWKB_EIO; }
        }
    }

    // ---- k-way merge (one run: a plain walk). Heap of run ids ordered by the head record's key; equal keys by run id = input order.
    int rc = TWKB_OK;
    {
        struct Head { SortKey k; uint32_t run; };
        auto worse = [](const Head& a, const Head& b) { return key_less(b.k, a.k) || (!key_less(a.k, b.k) && a.run > b.run); };  // min-heap
        std::vector<Head> heap;
        auto head_of = [&](uint32_t r) {
            SortKey k = key_of(runs[r].head(), 0);
            k.idx = 0;  // ties: run order (runs are consecutive input ranges and stable inside)
            return Head{k, r};
        };
        for (uint32_t r = 0; r < runs.size(); ++r)
            if (!runs[r].empty()) heap.push_back(head_of(r));
        std::make_heap(heap.begin(), heap.end(), worse);
        while (!heap.empty() && rc == TWKB_OK) {
            std::pop_heap(heap.begin(), heap.end(), worse);
            const uint32_t r = heap.back().run;
            heap.pop_back();
            if (!w.add(runs[r].head())) { rc = TWKB_EIO; err = w.err; break; }
            if (!runs[r].pop(err)) { rc = TWKB_EIO; break; }
            if (!runs[r].empty()) {
                heap.push_back(head_of(r));
                std::push_heap(heap.begin(), heap.end(), worse);
            }
        }
    }
    for (SortedRun& run : runs) if (run.fp) { std::fclose(run.fp); run.fp = nullptr; }
    if (rc == TWKB_OK) {
        w.close_block();
        if (!w.flush()) { rc = TWKB_EIO; err = w.err; }
    }
    if (rc != TWKB_OK) { cleanup(true); return rc; }

    // ---- sorted index + EOF
    const bool ok = write_sorted_tail(w, n_contigs, c_level);
    cleanup(!ok);
    if (!ok) { err = "Failed to write final block!"; return TWKB_EIO; }
    return TWKB_OK;
}

// ------------------------------------------------------------------ writer of an already sorted record stream
struct SortedTwoWriter::Impl {
    SortedWriter w;
    std::string path;
    uint64_t n_contigs = 0;
    int c_level = 1;
    bool have_prev = false;
    SortKey prev{};
};

SortedTwoWriter::~SortedTwoWriter() {
    if (p_) {
        if (p_->w.fp) { std::fclose(p_->w.fp); ::unlink(p_->path.c_str()); }  // never finished: no truncated file left behind
        delete p_;
    }
}

int SortedTwoWriter::open(const std::string& path, const TwkFile& src, const std::string& command_line, int c_level, int n_threads,
                          std::string& err) {
    p_ = new Impl;
    p_->path = path;
    p_->c_level = c_level;
    p_->n_contigs = src.n_contigs;
    SortedWriter& w = p_->w;
    w.c_level = c_level;
    w.n_threads = std::max(1, n_threads);
    w.n_contigs = src.n_contigs;
    w.meta.resize(src.n_contigs);
    w.fp = std::fopen(path.c_str(), "wb");
    if (!w.fp) { err = "Failed to open file: " + path + "..."; return TWKB_EIO; }
    char date[64];
    std::time_t now = std::time(nullptr);
    std::strftime(date, sizeof(date), "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    // the provenance of both steps the reference would have run: calc (ld.cpp:609-612), then sort (two_reader.cpp:346-349)
    const std::string literals = src.literals + "\n##tomahawk_calcVersion=b200-0.1.0\n##tomahawk_calcCommand=" + command_line + "; Date=" + date +
                                 "\n##tomahawk_sortVersion=b200-0.1.0\n##tomahawk_sortCommand=device radix sort of the resident records; Date=" + date + "\n";
    std::vector<uint8_t> hdr;
    put_str(hdr, src.fileformat);
    put_str(hdr, literals);
    hdr.insert(hdr.end(), src.header_tail.begin(), src.header_tail.end());
    std::vector<uint8_t> z(ZSTD_compressBound(hdr.size()));
    const size_t zn = ZSTD_compress(z.data(), z.size(), hdr.data(), hdr.size(), c_level);
    if (ZSTD_isError(zn)) { err = "failed to compress"; return TWKB_EIO; }
    const uint64_t unc = hdr.size(), cmp = zn;
    if (!w.emit(kTwoMagic, 4) || !w.emit(&unc, 8) || !w.emit(&cmp, 8) || !w.emit(z.data(), zn)) { err = "Failed to write header!"; return TWKB_EIO; }
    return TWKB_OK;
}

int SortedTwoWriter::add(const uint8_t* records, uint64_t n) {
    if (!p_ || !p_->w.fp) return TWKB_ESTATE;
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* rec = records + r * TWKB_RECORD_BYTES;
        SortKey k = key_of(rec, 0);
        if (p_->have_prev && key_less(k, p_->prev)) { err_ = "records are not in sorted order"; return TWKB_EINVAL; }
        p_->prev = k;
        p_->have_prev = true;
        if (!p_->w.add(rec)) { err_ = p_->w.err; return TWKB_EIO; }
    }
    return TWKB_OK;
}

int SortedTwoWriter::finish() {
    if (!p_ || !p_->w.fp) return TWKB_ESTATE;
    SortedWriter& w = p_->w;
    w.close_block();
    bool ok = w.flush();
    if (!ok) err_ = w.err;
    ok = ok && write_sorted_tail(w, p_->n_contigs, p_->c_level);
    if (w.fp) { std::fclose(w.fp); w.fp = nullptr; }
    if (!ok) {
        ::unlink(p_->path.c_str());
        if (err_.empty()) err_ = "Failed to write final block!";
        return TWKB_EIO;
    }
    return TWKB_OK;
}

}  // namespace twkb
