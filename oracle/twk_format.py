"""Test-side tooling for the Tomahawk on-disk formats and synthetic inputs.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. The product (tomahawk_b200/)
never imports this module.

What is here, and the reference layout each piece follows (paths relative to the
reference tree):

* ``synth_genotypes``    seeded synthetic diploid genotypes with a skewed
                         allele-frequency spectrum and LD blocks.
* ``pack_bits``          genotype codes -> the 1-bit/haplotype bitvector + missing
                         mask of ``twk_igt_vec::Build`` (lib/core.cpp:349-383):
                         haplotype p is bit p%64 of word p/64, sample s owns bits
                         2s, 2s+1; both mask bits are set if either allele is
                         missing.
* ``write_twk``          a valid ``.twk`` file (magic, zstd(VcfHeader), blocks of
                         <=500 ``twk1_t`` with run-length genotypes, zstd(Index)
                         footer, offset, 32-char EOF): lib/importer.cpp:82-326,
                         lib/core.cpp:59-73,245-251, include/core.h:195-205,
                         lib/index.cpp:8-18,90-99,158-166, lib/header.cpp:330-345,
                         include/header.h:115-128.
* ``read_two``           ``.two`` -> numpy structured array of the 106-byte records
                         (lib/core.cpp:470-490,626-631; include/writer.h:70-87).
"""
from __future__ import annotations

import ctypes
import struct
import numpy as np

# --------------------------------------------------------------------------- zstd
_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        lib = ctypes.CDLL("libzstd.so.1")
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_decompress.restype = ctypes.c_size_t
        lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        _zstd = lib
    return _zstd


def zstd_compress(data: bytes, level: int = 1) -> bytes:
    lib = _libzstd()
    cap = lib.ZSTD_compressBound(len(data))
    dst = ctypes.create_string_buffer(cap)
    n = lib.ZSTD_compress(dst, cap, data, len(data), level)
    if lib.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress failed")
    return dst.raw[:n]


def zstd_decompress(data: bytes, n_unc: int) -> bytes:
    lib = _libzstd()
    dst = ctypes.create_string_buffer(max(n_unc, 1))
    n = lib.ZSTD_decompress(dst, n_unc, data, len(data))
    if lib.ZSTD_isError(n) or n != n_unc:
        raise RuntimeError("ZSTD_decompress failed")
    return dst.raw[:n]


# ---------------------------------------------------------------- synthetic data
# The generator and the bit packer are shared with bench.py and live in the
# package (tomahawk_b200/synth.py, numpy only); re-exported here for the tests.
from tomahawk_b200.synth import Synth, pack_bits, synth_genotypes, words_per_variant  # noqa: E402,F401


# ------------------------------------------------------------------- .twk writer
def _wstr(b: bytearray, s: str | bytes):
    if isinstance(s, str):
        s = s.encode()
    b += struct.pack("<I", len(s)) + s


def _header_bytes(n_samples: int, contigs, literals: str) -> bytes:
    b = bytearray()
    _wstr(b, "##fileformat=VCFv4.2")
    _wstr(b, literals)
    b += struct.pack("<I", n_samples)
    for i in range(n_samples):
        _wstr(b, f"S{i}")
    b += struct.pack("<I", len(contigs))
    for idx, (name, n_bases) in enumerate(contigs):
        b += struct.pack("<I", idx)
        _wstr(b, name)
        _wstr(b, "")
        b += struct.pack("<q", n_bases)
        b += struct.pack("<I", 0)
    return bytes(b)


def _rle_variant(al: np.ndarray, has_missing: bool, force_ptype: int | None = None):
    """Run-length encode one variant's 2N allele codes -> (ptype, runs ndarray).

    Run word = len << (2+2*miss) | refA << (1+miss) | refB (include/core.h:195-198).
    """
    n = al.shape[0] // 2
    a = al[0::2].astype(np.uint32)
    b = al[1::2].astype(np.uint32)
    shift = 2 if has_missing else 1
    code = (a << shift) | b
    brk = np.flatnonzero(np.diff(code)) + 1
    starts = np.concatenate(([0], brk))
    lens = np.diff(np.concatenate((starts, [n])))
    codes = code[starts]
    lbits = 2 + 2 * int(has_missing)
    for ptype, dt in ((1, np.uint8), (2, np.uint16), (4, np.uint32)):
        if force_ptype is not None and ptype != force_ptype:
            continue
        maxlen = (1 << (8 * ptype - lbits)) - 1
        # split long runs into pieces of <= maxlen
        pieces = (lens + maxlen - 1) // maxlen
        total = int(pieces.sum())
        if force_ptype is None and ptype < 4 and total > 2 * len(lens) + 8:
            continue  # too much splitting; use a wider primitive
        rl = np.repeat(lens, pieces)
        rc = np.repeat(codes, pieces)
        # every piece but the last of each run is maxlen long
        last = np.cumsum(pieces) - 1
        full = np.full(total, maxlen, dtype=np.int64)
        full[last] = lens - (pieces - 1) * maxlen
        del rl
        runs = ((full.astype(np.uint64) << lbits) | rc.astype(np.uint64)).astype(dt)
        return ptype, runs
    raise AssertionError


def encode_runs(s: Synth, widths=None, seed: int = 0):
    """Run-length records of every variant of ``s`` laid out back to back at arbitrary (odd)
    offsets: (raw uint8, desc) as twkb_load_runs takes them. ``widths``: run-word widths to draw
    from per variant (default: the writer's own choice)."""
    rng = np.random.default_rng(seed)
    an = s.an
    chunks, desc, off = [], [], 0
    for v in range(s.n_variants):
        pad = int(rng.integers(0, 4))
        chunks.append(np.zeros(pad, dtype=np.uint8))
        off += pad
        miss = bool(an[v] != 0)
        ptype, runs = _rle_variant(s.alleles[v], miss, None if widths is None else int(rng.choice(widths)))
        b = np.frombuffer(runs.astype(runs.dtype.newbyteorder("<")).tobytes(), dtype=np.uint8)
        desc.append((off, len(runs), ptype, int(miss)))
        chunks.append(b)
        off += len(b)
    raw = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    d = np.zeros(len(desc), dtype=[("offset", "<u8"), ("n_runs", "<u4"), ("width", "u1"), ("miss", "u1"), ("pad", "u1", (2,))])
    for k, (o, n, w, m) in enumerate(desc):
        d[k]["offset"], d[k]["n_runs"], d[k]["width"], d[k]["miss"] = o, n, w, m
    return raw, d


def decode_runs(raw: np.ndarray, desc: np.ndarray, n_samples: int, with_mask: bool):
    """CPU restatement of twk_igt_vec::Build (reference lib/core.cpp:349-383) over located run
    words: returns (data, mask) uint64 rows [n_variants, stride] in the layout pack_bits() makes.
    ``desc`` has fields offset / n_runs / width / miss (tomahawk_b200.RUN_DESC_DTYPE).
    Test oracle of the device decoder (decode.cuh); raises if a variant's runs do not cover
    exactly n_samples samples."""
    H = 2 * n_samples
    stride = ((H + 63) // 64 + 1) // 2 * 2
    M = len(desc)
    data = np.zeros((M, stride * 64), dtype=np.uint8)
    mask = np.zeros((M, stride * 64), dtype=np.uint8) if with_mask else None
    for v in range(M):
        d = desc[v]
        w = int(d["width"])
        words = np.frombuffer(raw[int(d["offset"]): int(d["offset"]) + int(d["n_runs"]) * w].tobytes(),
                              dtype={1: "<u1", 2: "<u2", 4: "<u4"}[w]).astype(np.uint64)
        miss = int(d["miss"])
        lens = (words >> np.uint64(2 + 2 * miss)).astype(np.int64)
        a = ((words >> np.uint64(1 + miss)) & np.uint64((1 << (1 + miss)) - 1)).astype(np.int64)
        b = (words & np.uint64((1 << (1 + miss)) - 1)).astype(np.int64)
        if int(lens.sum()) != n_samples:
            raise ValueError(f"variant {v}: runs cover {int(lens.sum())} of {n_samples} samples")
        ea = np.repeat(a, lens)
        eb = np.repeat(b, lens)
        data[v, 0:H:2] = ea == 1   # lib/core.cpp:371-376
        data[v, 1:H:2] = eb == 1
        if mask is not None:
            m = (ea == 2) | (eb == 2)  # :379-380: both bits of the sample
            mask[v, 0:H:2] = m
            mask[v, 1:H:2] = m
    pack = lambda bits: np.packbits(bits, axis=1, bitorder="little").view("<u8")
    return pack(data), (pack(mask) if mask is not None else None)


def write_twk(path: str, s: Synth, block_size: int = 500, c_level: int = 1, contigs=None):
    """Write ``s`` as a .twk file the reference's twk_reader::Open accepts."""
    if contigs is None:
        contigs = [("1", int(s.pos.max()) + 1000)]
    M = s.n_variants
    ac = s.ac
    an = s.an
    out = bytearray()
    out += b"TOMAHAWK\x01"
    hdr = _header_bytes(s.n_samples, contigs, "##fileformat=VCFv4.2\n##source=tomahawk_b200-synth\n")
    hc = zstd_compress(hdr, c_level)
    out += struct.pack("<QQ", len(hdr), len(hc)) + hc
    index_entries = []
    # blocks: <= block_size variants, one contig per block
    start = 0
    while start < M:
        end = min(start + block_size, M)
        same = np.flatnonzero(s.rid[start:end] != s.rid[start])
        if len(same):
            end = start + int(same[0])
        n = end - start
        blk = bytearray()
        blk += struct.pack("<III", n, max(n, 500), int(s.rid[start]))
        for v in range(start, end):
            al = s.alleles[v]
            miss = bool(an[v] != 0)
            ptype, runs = _rle_variant(al, miss)
            g = al.reshape(-1, 2)
            n_het = int(((g[:, 0] != g[:, 1]) & (g[:, 0] != 2) & (g[:, 1] != 2)).sum())
            n_hom = int(((g[:, 0] == 1) & (g[:, 1] == 1)).sum())
            pack = (ptype << 3) | (0 << 2) | (int(s.phased) << 1) | int(miss)
            blk += struct.pack("<BB", pack, (0 << 4) | 1)
            blk += struct.pack("<IIIIII", int(s.pos[v]), int(ac[v]), int(an[v]), int(s.rid[v]), n_het, n_hom)
            blk += struct.pack("<d", 1.0)
            blk += struct.pack("<I", (len(runs) << 1) | int(miss))
            blk += runs.astype(runs.dtype.newbyteorder("<")).tobytes()
        bc = zstd_compress(bytes(blk), c_level)
        foff = len(out)
        out += struct.pack("<BII", 1, len(blk), len(bc)) + bc
        fend = len(out)
        index_entries.append(
            (int(s.rid[start]), n, int(s.pos[start]) + 1, int(s.pos[end - 1]) + 1, len(blk), len(bc), foff, fend)
        )
        start = end
    # footer
    idx = bytearray()
    n_ent = len(index_entries)
    idx += struct.pack("<QQQQ", 1954702206512158641, n_ent, max(n_ent, 1), len(contigs))
    for e in index_entries:
        idx += struct.pack("<iIIIIIQQ", *e)
    for c in range(len(contigs)):
        ents = [e for e in index_entries if e[0] == c]
        if ents:
            idx += struct.pack(
                "<iIIIQQQ", c, sum(e[1] for e in ents), ents[0][2], ents[-1][3], ents[0][6], ents[-1][7], len(ents)
            )
        else:
            idx += struct.pack("<iIIIQQQ", 0, 0, 0, 0, 0, 0, 0)
    ic = zstd_compress(bytes(idx), c_level)
    off = len(out)
    out += struct.pack("<BQQ", 0, len(idx), len(ic)) + ic
    out += struct.pack("<Q", off)
    out += b"a4f54f39f5e251a6993796f48164ccf554f1b680c2ebbb13be301f3ff76f82cf"[:32]
    with open(path, "wb") as f:
        f.write(out)
    return n_ent


# ------------------------------------------------------------------ .two reader
TWO_DTYPE = np.dtype(
    [
        ("controller", "<u2"),
        ("ridA", "<u4"),
        ("ridB", "<u4"),
        ("packA", "<u4"),
        ("packB", "<u4"),
        ("cnt", "<f8", (4,)),
        ("D", "<f8"),
        ("Dprime", "<f8"),
        ("R", "<f8"),
        ("R2", "<f8"),
        ("P", "<f8"),
        ("ChiSqFisher", "<f8"),
        ("ChiSqModel", "<f8"),
    ]
)
assert TWO_DTYPE.itemsize == 106


def read_two(path: str) -> np.ndarray:
    """All records of a .two file (both orientations, file order)."""
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:4] == b"TWO\x01", "bad .two magic"
    p = 4
    unc, cmp_ = struct.unpack_from("<QQ", buf, p)
    p += 16 + cmp_
    chunks = []
    while True:
        marker = buf[p]
        p += 1
        if marker == 0:
            break
        assert marker == 1
        unc, cmp_ = struct.unpack_from("<II", buf, p)
        p += 8
        raw = zstd_decompress(buf[p : p + cmp_], unc)
        p += cmp_
        n, m = struct.unpack_from("<II", raw, 0)
        assert unc == 8 + 106 * n
        chunks.append(np.frombuffer(raw, dtype=TWO_DTYPE, count=n, offset=8))
    if not chunks:
        return np.zeros(0, dtype=TWO_DTYPE)
    return np.concatenate(chunks)


def read_two_index(path: str):
    """(state, block entries, per-contig entries) of a .two file's index (lib/index.cpp:41-52, 90-99, 242-251):
    block entry = (rid, n, minpos, maxpos, b_unc, b_cmp, foff, fend, ridB); contig entry = (rid, n, minpos,
    maxpos, foff, fend, nn)."""
    raw = open(path, "rb").read()
    off = struct.unpack("<Q", raw[-40:-32])[0]
    unc, cmp_ = struct.unpack("<QQ", raw[off + 1: off + 17])
    idx = zstd_decompress(raw[off + 17: off + 17 + cmp_], unc)
    marker, state, n, m, m_ent = struct.unpack("<QBQQQ", idx[:33])
    assert marker == 1954702206512158641
    p = 33
    ents, meta = [], []
    for _ in range(n):
        ents.append(struct.unpack("<iIIIIIQQi", idx[p:p + 44])); p += 44
    for _ in range(m_ent):
        meta.append(struct.unpack("<iIIIQQQ", idx[p:p + 40])); p += 40
    return state, ents, meta


def records_from_bytes(raw: bytes | np.ndarray) -> np.ndarray:
    return np.frombuffer(bytes(raw), dtype=TWO_DTYPE)


def canonical(recs: np.ndarray, forward_only: bool = True) -> np.ndarray:
    """Sort records by (ridA,posA,ridB,posB); optionally keep only the forward
    orientation (posA < posB on one contig) -- the reference emits each passing
    pair twice (lib/ld/ld_engine.cpp:1290-1298)."""
    posA = recs["packA"] >> 2
    posB = recs["packB"] >> 2
    if forward_only:
        keep = (recs["ridA"] < recs["ridB"]) | ((recs["ridA"] == recs["ridB"]) & (posA < posB))
        recs, posA, posB = recs[keep], posA[keep], posB[keep]
    order = np.lexsort((posB, recs["ridB"], posA, recs["ridA"]))
    return recs[order]
