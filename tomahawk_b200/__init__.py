"""tomahawk_b200 -- B200-native pairwise-LD engine (`tomahawk calc`).

The product is ``libtwkb.so`` (CUDA kernels for sm_100a behind the C-ABI of
``include/twkb.h``). This module is only the thin ctypes mirror of that ABI used
by the tests, ``bench.py`` and Python callers; it contains no compute and there
is no CPU fallback: if the shared library is missing or no B200 is present the
calls fail loudly.

Reference seam mirrored here: ``bool twk_ld::Compute(const twk_ld_settings&)``
(reference include/ld.h:53, lib/ld/ld.cpp:477-671) -> :class:`twk_ld`.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtwkb.so")
# the same sources built with -DTWKB_PROFILING: kernel ablation switches (TWKB_DEBUG_FLAGS) compiled in;
# measurement aid of bench.py / scripts only, results with a switch on are invalid
PROF_LIB_PATH = os.path.join(_HERE, "libtwkb_prof.so")

RECORD_BYTES = 106
KERNEL_AUTO, KERNEL_POPC, KERNEL_UMMA, KERNEL_UMMA_FP4 = 0, 1, 2, 3

ERRORS = {
    -1: "TWKB_EINVAL", -2: "TWKB_ENODEVICE", -3: "TWKB_ECUDA", -4: "TWKB_ENOMEM",
    -5: "TWKB_ESTATE", -6: "TWKB_ESINK", -7: "TWKB_EIO",
}


class TwkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Settings(ctypes.Structure):
    """twkb_settings: 1:1 with twk_ld_settings (reference include/core.h:909-924)."""

    _fields_ = [
        ("square", ctypes.c_uint8), ("window", ctypes.c_uint8), ("low_memory", ctypes.c_uint8),
        ("bitmaps", ctypes.c_uint8), ("single", ctypes.c_uint8), ("force_phased", ctypes.c_uint8),
        ("forced_unphased", ctypes.c_uint8), ("emulate_quirks", ctypes.c_uint8),
        ("c_level", ctypes.c_int32), ("bl_size", ctypes.c_int32), ("b_size", ctypes.c_int32),
        ("l_window", ctypes.c_int32), ("n_threads", ctypes.c_int32), ("l_surrounding", ctypes.c_int32),
        ("n_chunks", ctypes.c_int32), ("c_chunk", ctypes.c_int32),
        ("minP", ctypes.c_double), ("minR2", ctypes.c_double), ("maxR2", ctypes.c_double),
        ("minDprime", ctypes.c_double), ("maxDprime", ctypes.c_double),
        ("device", ctypes.c_int32), ("part_index", ctypes.c_int32), ("part_count", ctypes.c_int32),
        ("kernel", ctypes.c_int32), ("twk_block_size", ctypes.c_int32), ("sparse_max_words", ctypes.c_int32),
        ("host_unpack", ctypes.c_int32), ("single_targets", ctypes.c_int32), ("shard_blocks", ctypes.c_int32),
        ("sorted_output", ctypes.c_int32),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("pairs_visited", ctypes.c_uint64), ("pairs_screened", ctypes.c_uint64), ("records_out", ctypes.c_uint64),
        ("count_launches", ctypes.c_uint64), ("stats_launches", ctypes.c_uint64), ("other_launches", ctypes.c_uint64),
        ("seconds_total", ctypes.c_double), ("ms_count_kernel", ctypes.c_double), ("ms_stats_kernel", ctypes.c_double),
        ("ms_h2d", ctypes.c_double), ("bytes_h2d", ctypes.c_uint64), ("bytes_d2h", ctypes.c_uint64),
        ("kernel_used", ctypes.c_int32), ("n_planes", ctypes.c_int32), ("word_ops", ctypes.c_uint64),
        ("mma_macs", ctypes.c_uint64), ("ms_device_total", ctypes.c_double),
        ("sparse_variants", ctypes.c_uint64), ("sparse_launches", ctypes.c_uint64), ("sparse_word_ops", ctypes.c_uint64),
        ("ms_sparse_kernel", ctypes.c_double), ("ms_decode_kernel", ctypes.c_double),
        ("seconds_file_read", ctypes.c_double), ("seconds_file_load", ctypes.c_double), ("seconds_file_total", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


VARIANT_DTYPE = np.dtype(
    [("rid", "<u4"), ("pos", "<u4"), ("ac", "<u4"), ("an", "<u4"), ("hwe", "<f8"),
     ("gt_missing", "u1"), ("gt_phase", "u1"), ("pad", "u1", (6,))]
)
TWO_DTYPE = np.dtype(
    [("controller", "<u2"), ("ridA", "<u4"), ("ridB", "<u4"), ("packA", "<u4"), ("packB", "<u4"),
     ("cnt", "<f8", (4,)), ("D", "<f8"), ("Dprime", "<f8"), ("R", "<f8"), ("R2", "<f8"), ("P", "<f8"),
     ("ChiSqFisher", "<f8"), ("ChiSqModel", "<f8")]
)
RUN_DESC_DTYPE = np.dtype([("offset", "<u8"), ("n_runs", "<u4"), ("width", "u1"), ("miss", "u1"), ("pad", "u1", (2,))])
AGG_BIN_DTYPE = np.dtype([("n", "<u8"), ("total", "<f8"), ("total_squared", "<f8"), ("min", "<f8"), ("max", "<f8")])  # twk_sstats
AGG_LAYOUT_DTYPE = np.dtype([("range", "<u8"), ("bpx", "<u4"), ("bpy", "<u4"), ("n_contigs_set", "<u4"), ("pad", "<u4"), ("n_records", "<u8")])
AGG_FIELDS = {"r2": 0, "r": 1, "d": 2, "dprime": 3, "dp": 3, "p": 4, "hets": 5, "het": 5, "alts": 6, "alt": 6}  # two_reader.cpp:574-588
CAND_DTYPE = np.dtype([("i", "<u4"), ("j", "<u4"), ("c", "<u4", (9,)), ("mode", "<u4")])
SINK_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_uint64)

_lib = None
_prof_lib = None

EXPORTS = [
    "twkb_settings_init", "twkb_create", "twkb_destroy", "twkb_last_error", "twkb_update_settings",
    "twkb_load_matrix", "twkb_load_matrix_device", "twkb_compute", "twkb_compute_resident", "twkb_get_stats",
    "twkb_debug_candidates", "twkb_calc_file", "twkb_calc_file_intervals", "twkb_version",
    "twkb_twk_open", "twkb_twk_open_intervals", "twkb_twk_dims", "twkb_twk_copy", "twkb_twk_view", "twkb_twk_close",
    "twkb_two_open", "twkb_two_add", "twkb_two_close", "twkb_plan_tiles",
    "twkb_load_runs", "twkb_debug_rows", "twkb_twk_open_runs", "twkb_twk_runs_view", "twkb_two_set_threads",
    "twkb_two_sort",
    "twkb_compute_decay", "twkb_set_blocks", "twkb_twk_blocks", "twkb_two_sort_mem", "twkb_twk_open_single", "twkb_comm_unique_id", "twkb_comm_init", "twkb_comm_slice", "twkb_load_matrix_sliced", "twkb_load_runs_sliced",
    "twkb_plan_shards", "twkb_compute_aggregate", "twkb_twk_contigs",
    "twkb_compute_sorted", "twkb_two_open_sorted", "twkb_two_add_sorted", "twkb_two_close_sorted",
]


def lib(profiling: bool = False):
    """Load libtwkb.so (or, profiling=True, libtwkb_prof.so). Raises if it has not been built (no fallback of any kind)."""
    global _lib, _prof_lib
    if profiling:
        if _prof_lib is None:
            if not os.path.exists(PROF_LIB_PATH):
                raise ImportError(f"{PROF_LIB_PATH} is missing (make -C tomahawk_b200/csrc)")
            _prof_lib = _bind(ctypes.CDLL(PROF_LIB_PATH))
        return _prof_lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). tomahawk_b200 has no CPU fallback."
            )
        _lib = _bind(ctypes.CDLL(LIB_PATH))
    return _lib


def _bind(L):
    L.twkb_settings_init.argtypes = [ctypes.POINTER(Settings)]
    L.twkb_settings_init.restype = None
    L.twkb_create.argtypes = [ctypes.POINTER(Settings), ctypes.POINTER(ctypes.c_void_p)]
    L.twkb_destroy.argtypes = [ctypes.c_void_p]
    L.twkb_destroy.restype = None
    L.twkb_last_error.argtypes = [ctypes.c_void_p]
    L.twkb_last_error.restype = ctypes.c_char_p
    L.twkb_update_settings.argtypes = [ctypes.c_void_p, ctypes.POINTER(Settings)]
    for name in ("twkb_load_matrix", "twkb_load_matrix_device"):
        getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.twkb_load_runs.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_void_p, ctypes.c_void_p]
    L.twkb_load_matrix_sliced.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.twkb_load_runs_sliced.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_void_p]
    L.twkb_comm_unique_id.argtypes = [ctypes.c_void_p]
    L.twkb_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
    L.twkb_comm_slice.argtypes = [ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint32),
                                  ctypes.POINTER(ctypes.c_uint32)]
    L.twkb_debug_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.twkb_twk_open_runs.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
                                     ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_twk_open_single.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_twk_runs_view.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]
    L.twkb_two_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.twkb_two_sort.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
                                ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_compute_decay.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    L.twkb_compute_aggregate.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_uint32,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.twkb_twk_contigs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
    L.twkb_set_blocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    L.twkb_twk_blocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32)]
    L.twkb_two_sort_mem.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64,
                                    ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_compute.argtypes = [ctypes.c_void_p, SINK_FN, ctypes.c_void_p]
    L.twkb_compute_sorted.argtypes = [ctypes.c_void_p, SINK_FN, ctypes.c_void_p]
    L.twkb_two_open_sorted.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_two_add_sorted.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    L.twkb_two_close_sorted.argtypes = [ctypes.c_void_p]
    L.twkb_compute_resident.argtypes = [ctypes.c_void_p]
    L.twkb_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
    L.twkb_debug_candidates.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_uint64)]
    L.twkb_calc_file.argtypes = [ctypes.POINTER(Settings), ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(Stats),
                                 ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_twk_open.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_calc_file_intervals.argtypes = [ctypes.POINTER(Settings), ctypes.c_char_p, ctypes.c_char_p,
                                           ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32, ctypes.POINTER(Stats),
                                           ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_twk_open_intervals.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int32,
                                          ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_twk_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint32)]
    L.twkb_twk_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.twkb_twk_close.argtypes = [ctypes.c_void_p]
    L.twkb_twk_close.restype = None
    L.twkb_two_open.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32,
                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    L.twkb_two_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    L.twkb_two_close.argtypes = [ctypes.c_void_p]
    L.twkb_plan_tiles.argtypes = [ctypes.POINTER(Settings), ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                  ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    return L


def _c_strings(strings):
    if not strings:
        return None
    arr = (ctypes.c_char_p * len(strings))()
    arr[:] = [x.encode() for x in strings]
    return arr


class TwkFile:
    """A .twk file unpacked by the host reader (twkb_twk_*)."""

    def __init__(self, path: str, n_threads: int = 4, intervals=(), emulate_quirks: bool = True, runs: bool = False,
                 single_surrounding: int | None = None):
        """``intervals``: the ``-I`` strings of ``calc`` (block-granular selection, lib/ld/ld.cpp:257-365).
        ``runs``: keep the genotypes run-length encoded for the device decoder (:meth:`runs`).
        ``single_surrounding``: scalc selection (lib/ld/ld.cpp:123-255): ``intervals`` holds the ONE target string; the
        handle then holds [targets | variants within that many bases] and ``n_targets``."""
        L = lib()
        self._L = L
        self._h = ctypes.c_void_p()
        self.runs_mode = runs
        self.n_targets = 0
        err = ctypes.create_string_buffer(512)
        iv = _c_strings(intervals)
        if single_surrounding is not None:
            nt = ctypes.c_uint32(0)
            rc = L.twkb_twk_open_single(path.encode(), n_threads, intervals[0].encode() if intervals else None, single_surrounding,
                                        int(emulate_quirks), int(runs), ctypes.byref(self._h), ctypes.byref(nt), err, 512)
            self.n_targets = nt.value
        else:
            opener = L.twkb_twk_open_runs if runs else L.twkb_twk_open_intervals
            rc = opener(path.encode(), n_threads, iv, len(intervals), int(emulate_quirks), ctypes.byref(self._h), err, 512)
        if rc != 0:
            raise TwkbError(rc, err.value.decode())
        ns, nv, st, am, nb = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_size_t(), ctypes.c_int32(), ctypes.c_uint32()
        L.twkb_twk_dims(self._h, ctypes.byref(ns), ctypes.byref(nv), ctypes.byref(st), ctypes.byref(am), ctypes.byref(nb))
        self.n_samples, self.n_variants, self.stride = ns.value, nv.value, st.value
        self.any_missing, self.n_blocks = bool(am.value), nb.value

    def runs(self):
        """(run_bytes uint8[n], desc RUN_DESC_DTYPE[n_variants], meta) -- views into the handle's buffers,
        valid until close(); what Engine.load_runs takes."""
        b, n, d, m = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p(), ctypes.c_void_p()
        rc = self._L.twkb_twk_runs_view(self._h, ctypes.byref(b), ctypes.byref(n), ctypes.byref(d), ctypes.byref(m))
        if rc != 0:
            raise TwkbError(rc, "handle was not opened in runs mode")
        raw = np.ctypeslib.as_array(ctypes.cast(b, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value,))
        desc = np.frombuffer((ctypes.c_uint8 * (16 * self.n_variants)).from_address(d.value), dtype=RUN_DESC_DTYPE)
        meta = np.frombuffer((ctypes.c_uint8 * (32 * self.n_variants)).from_address(m.value), dtype=VARIANT_DTYPE)
        return raw, desc, meta

    def contigs(self) -> np.ndarray:
        """Contig lengths in header order (int64)."""
        n = ctypes.c_uint32(0)
        self._L.twkb_twk_contigs(self._h, None, 0, ctypes.byref(n))
        out = np.zeros(n.value, dtype=np.int64)
        rc = self._L.twkb_twk_contigs(self._h, out.ctypes.data, n.value, ctypes.byref(n))
        if rc != 0:
            raise TwkbError(rc, "twkb_twk_contigs")
        return out

    def blocks(self) -> np.ndarray:
        """First variant of every loaded .twk block (file order)."""
        p, n = ctypes.c_void_p(), ctypes.c_uint32()
        self._L.twkb_twk_blocks(self._h, ctypes.byref(p), ctypes.byref(n))
        if not n.value:
            return np.zeros(0, np.uint32)
        return np.frombuffer((ctypes.c_uint8 * (4 * n.value)).from_address(p.value), dtype=np.uint32).copy()

    def matrix(self):
        if self.runs_mode:
            raise TwkbError(-5, "handle holds run-length records; use runs()")
        data = np.zeros((self.n_variants, self.stride), dtype=np.uint64)
        mask = np.zeros_like(data) if self.any_missing else None
        meta = np.zeros(self.n_variants, dtype=VARIANT_DTYPE)
        self._L.twkb_twk_copy(self._h, data.ctypes.data, mask.ctypes.data if mask is not None else None, meta.ctypes.data)
        return data, mask, meta

    def close(self):
        if self._h:
            self._L.twkb_twk_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TwoWriter:
    """Streaming .two writer (twkb_two_*): forward records in, forward + reverse blocks out."""

    def __init__(self, path: str, twk: TwkFile, command_line: str = "", c_level: int = 1, b_size: int = 10000,
                 n_threads: int = 1):
        L = lib()
        self._L = L
        self._w = ctypes.c_void_p()
        err = ctypes.create_string_buffer(512)
        rc = L.twkb_two_open(path.encode(), twk._h, command_line.encode(), c_level, b_size, ctypes.byref(self._w), err, 512)
        if rc != 0:
            raise TwkbError(rc, err.value.decode())
        L.twkb_two_set_threads(self._w, n_threads)

    def add(self, records: np.ndarray):
        records = np.ascontiguousarray(records)
        assert records.dtype.itemsize == RECORD_BYTES
        rc = self._L.twkb_two_add(self._w, records.ctypes.data, len(records))
        if rc != 0:
            raise TwkbError(rc, "twkb_two_add failed")

    def close(self):
        if self._w:
            rc = self._L.twkb_two_close(self._w)
            self._w = ctypes.c_void_p()
            if rc != 0:
                raise TwkbError(rc, "twkb_two_close failed")


class SortedTwoWriter:
    """Writer of a SORTED .two (twkb_two_*_sorted): takes the record stream of Engine.compute_sorted()."""

    def __init__(self, path: str, twk: TwkFile, command_line: str = "", c_level: int = 1, n_threads: int = 1):
        L = lib()
        self._L = L
        self._w = ctypes.c_void_p()
        err = ctypes.create_string_buffer(512)
        rc = L.twkb_two_open_sorted(path.encode(), twk._h, command_line.encode(), c_level, n_threads, ctypes.byref(self._w), err, 512)
        if rc != 0:
            raise TwkbError(rc, err.value.decode())

    def add(self, records: np.ndarray):
        records = np.ascontiguousarray(records)
        assert records.dtype.itemsize == RECORD_BYTES
        rc = self._L.twkb_two_add_sorted(self._w, records.ctypes.data, len(records))
        if rc != 0:
            raise TwkbError(rc, "twkb_two_add_sorted failed (records out of order?)")

    def close(self):
        if self._w:
            rc = self._L.twkb_two_close_sorted(self._w)
            self._w = ctypes.c_void_p()
            if rc != 0:
                raise TwkbError(rc, "twkb_two_close_sorted failed")


def sort_two(in_path: str, out_path: str, c_level: int = 1, n_threads: int = 4, memory_limit: int = 0) -> int:
    """`tomahawk sort` (two_reader::Sort): writes the sorted, indexed .two file; returns the record count.
    ``memory_limit`` bytes (0 = unbounded): larger inputs go through spilled runs and a k-way merge."""
    n = ctypes.c_uint64(0)
    err = ctypes.create_string_buffer(512)
    rc = lib().twkb_two_sort_mem(in_path.encode(), out_path.encode(), c_level, n_threads, memory_limit, ctypes.byref(n), err, 512)
    if rc != 0:
        raise TwkbError(rc, err.value.decode())
    return int(n.value)


def plan_tiles(settings: Settings, meta: np.ndarray, tile_i: int, tile_j: int):
    """(tiles[n,2] uint32, n_pairs) of this settings.part_index's share of the grid."""
    L = lib()
    meta = np.ascontiguousarray(meta)
    n, pairs = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = L.twkb_plan_tiles(ctypes.byref(settings), len(meta), meta.ctypes.data, tile_i, tile_j, None, 0, ctypes.byref(n), ctypes.byref(pairs))
    if rc != 0:
        raise TwkbError(rc, "twkb_plan_tiles failed")
    out = np.zeros((int(n.value), 2), dtype=np.uint32)
    if n.value:
        rc = L.twkb_plan_tiles(ctypes.byref(settings), len(meta), meta.ctypes.data, tile_i, tile_j, out.ctypes.data, n.value, ctypes.byref(n), ctypes.byref(pairs))
        if rc != 0:
            raise TwkbError(rc, "twkb_plan_tiles failed")
    return out, int(pairs.value)


COMM_ID_BYTES = 128


def aggregate_reduce(bins: np.ndarray, reduce: str = "mean", min_cutoff: int = 5) -> np.ndarray:
    """The reduce functions of `tomahawk aggregate -r` over a raster of twk_sstats (include/core.h:957-976), including their
    cutoff rules: mean = 0 when n < cutoff or cutoff == 0; count = 0 when n < cutoff; sd = 0 when n < cutoff; total = 0 when
    total < cutoff; min / max ignore the cutoff."""
    n, tot, sq = bins["n"].astype(np.float64), bins["total"], bins["total_squared"]
    r = reduce.lower()
    safe = np.maximum(n, 1.0)
    if r == "mean":
        return np.where((n < min_cutoff) | (min_cutoff == 0), 0.0, tot / safe)
    if r in ("count", "n"):
        return np.where(n < min_cutoff, 0.0, n)
    if r == "total":
        return np.where(tot < min_cutoff, 0.0, tot)
    if r == "sd":
        with np.errstate(invalid="ignore"):   # like the reference: NaN where rounding makes the variance negative
            return np.where(n < min_cutoff, 0.0, np.sqrt(sq / safe - (tot / safe) * (tot / safe)))
    if r == "min":
        return bins["min"].copy()
    if r == "max":
        return bins["max"].copy()
    raise TwkbError(-1, f'Unknown reduce function "{reduce}"...')


def plan_shards(block_first, meta: np.ndarray, l_window: int, n_shards: int):
    """twkb_plan_shards: position shards of a -w run. block_first: first variant of every .twk block (+ n_variants at the
    end). Returns (own_begin [n_shards + 1], halo_end [n_shards]) in blocks."""
    bf = np.ascontiguousarray(block_first, dtype=np.uint32)
    meta = np.ascontiguousarray(meta)
    own = np.zeros(n_shards + 1, dtype=np.uint32)
    halo = np.zeros(n_shards, dtype=np.uint32)
    rc = lib().twkb_plan_shards(bf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(bf) - 1), meta.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_uint32(len(meta)), ctypes.c_int32(l_window), ctypes.c_int32(n_shards),
                                own.ctypes.data_as(ctypes.c_void_p), halo.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise TwkbError(rc, "twkb_plan_shards")
    return own, halo


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through libtwkb: create on one rank, ship the bytes to the others (any transport)."""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = lib().twkb_comm_unique_id(buf)
    if rc != 0:
        raise TwkbError(rc, lib().twkb_last_error(None).decode())
    return buf.raw


def comm_slice(n_variants: int, rank: int, n_ranks: int):
    """Rows [begin, end) a rank uploads in a sliced load."""
    b, e = ctypes.c_uint32(), ctypes.c_uint32()
    rc = lib().twkb_comm_slice(n_variants, rank, n_ranks, ctypes.byref(b), ctypes.byref(e))
    if rc != 0:
        raise TwkbError(rc, "twkb_comm_slice")
    return b.value, e.value


def default_settings(**kw) -> Settings:
    """Reference defaults (lib/core.cpp:297-306) with keyword overrides."""
    s = Settings()
    lib().twkb_settings_init(ctypes.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(f"twkb_settings has no field {k}")
        setattr(s, k, v)
    return s


class Engine:
    """One device context: resident genotype matrix + LD computation."""

    def __init__(self, settings: Settings | None = None, profiling: bool = False, **kw):
        self._L = lib(profiling)
        self.settings = settings if settings is not None else default_settings(**kw)
        self._ctx = ctypes.c_void_p()
        rc = self._L.twkb_create(ctypes.byref(self.settings), ctypes.byref(self._ctx))
        if rc != 0:
            raise TwkbError(rc, self._L.twkb_last_error(None).decode())
        self._keep = None

    def close(self):
        if self._ctx:
            self._L.twkb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise TwkbError(rc, self._L.twkb_last_error(self._ctx).decode())

    def update(self, **kw):
        for k, v in kw.items():
            setattr(self.settings, k, v)
        self._check(self._L.twkb_update_settings(self._ctx, ctypes.byref(self.settings)))

    def load(self, n_samples: int, data: np.ndarray, mask: np.ndarray | None, meta: np.ndarray):
        """data/mask: uint64 [n_variants, stride] rows in the twk_igt_vec layout."""
        data = np.ascontiguousarray(data, dtype=np.uint64)
        meta = np.ascontiguousarray(meta)
        assert meta.dtype.itemsize == 32
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint64)
            assert mask.shape == data.shape
        self._check(self._L.twkb_load_matrix(self._ctx, n_samples, data.shape[0], data.ctypes.data,
                                             mask.ctypes.data if mask is not None else None, data.shape[1],
                                             meta.ctypes.data))

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """Join the NCCL communicator of the sliced loads (collective over all ranks' contexts)."""
        assert len(unique_id) == COMM_ID_BYTES
        self._check(self._L.twkb_comm_init(self._ctx, unique_id, rank, n_ranks))

    def load_sliced(self, n_samples: int, n_variants: int, slice_data: np.ndarray, slice_mask: np.ndarray | None, meta: np.ndarray):
        """Collective: this rank's rows (comm_slice) go up over its own PCIe link, NCCL completes the matrix."""
        slice_data = np.ascontiguousarray(slice_data, dtype=np.uint64)
        meta = np.ascontiguousarray(meta)
        assert meta.dtype.itemsize == 32 and len(meta) == n_variants
        if slice_mask is not None:
            slice_mask = np.ascontiguousarray(slice_mask, dtype=np.uint64)
        self._check(self._L.twkb_load_matrix_sliced(self._ctx, n_samples, n_variants, slice_data.ctypes.data,
                                                    slice_mask.ctypes.data if slice_mask is not None else None,
                                                    slice_data.shape[1], meta.ctypes.data))

    def load_runs_sliced(self, n_samples: int, run_bytes: np.ndarray, desc: np.ndarray, meta: np.ndarray):
        run_bytes = np.ascontiguousarray(run_bytes, dtype=np.uint8)
        desc = np.ascontiguousarray(desc)
        meta = np.ascontiguousarray(meta)
        self._check(self._L.twkb_load_runs_sliced(self._ctx, n_samples, len(desc), run_bytes.ctypes.data, run_bytes.size,
                                                  desc.ctypes.data, meta.ctypes.data))

    def set_blocks(self, block_first):
        """The .twk block structure (first variant of every block) the window rules / -c chunks are defined on."""
        bf = np.ascontiguousarray(block_first, dtype=np.uint32)
        self._check(self._L.twkb_set_blocks(self._ctx, bf.ctypes.data, len(bf)))

    def load_runs(self, n_samples: int, run_bytes: np.ndarray, desc: np.ndarray, meta: np.ndarray):
        """Run-length records (twk1_igt_t words located by ``desc``) -> resident rows, decoded on the device."""
        run_bytes = np.ascontiguousarray(run_bytes, dtype=np.uint8)
        desc = np.ascontiguousarray(desc)
        meta = np.ascontiguousarray(meta)
        assert desc.dtype.itemsize == 16 and meta.dtype.itemsize == 32 and len(desc) == len(meta)
        self._check(self._L.twkb_load_runs(self._ctx, n_samples, len(desc), run_bytes.ctypes.data, run_bytes.size,
                                           desc.ctypes.data, meta.ctypes.data))

    def rows(self, n_variants: int, stride: int, with_mask: bool = False):
        """Test hook: the resident reference-layout rows copied back to the host."""
        data = np.zeros((n_variants, stride), dtype=np.uint64)
        mask = np.zeros_like(data) if with_mask else None
        self._check(self._L.twkb_debug_rows(self._ctx, data.ctypes.data, mask.ctypes.data if with_mask else None, stride))
        return data, mask

    def load_device(self, n_samples, n_variants, d_data_ptr, d_mask_ptr, stride, meta):
        meta = np.ascontiguousarray(meta)
        self._check(self._L.twkb_load_matrix_device(self._ctx, n_samples, n_variants, d_data_ptr, d_mask_ptr, stride,
                                                    meta.ctypes.data))

    def compute(self) -> np.ndarray:
        """Run and collect the forward records (host sink) as a TWO_DTYPE array."""
        chunks = []

        def _sink(user, ptr, n):
            buf = ctypes.string_at(ptr, int(n) * RECORD_BYTES)
            chunks.append(np.frombuffer(buf, dtype=TWO_DTYPE))
            return 0

        cb = SINK_FN(_sink)
        self._check(self._L.twkb_compute(self._ctx, cb, None))
        if not chunks:
            return np.zeros(0, dtype=TWO_DTYPE)
        return np.concatenate(chunks)

    def compute_sorted(self) -> np.ndarray:
        """Run and collect forward AND reverse records in `tomahawk sort` order (sorted on the device)."""
        chunks = []

        def _sink(user, ptr, n):
            chunks.append(np.frombuffer(ctypes.string_at(ptr, int(n) * RECORD_BYTES), dtype=TWO_DTYPE))
            return 0

        cb = SINK_FN(_sink)
        self._check(self._L.twkb_compute_sorted(self._ctx, cb, None))
        return np.concatenate(chunks) if chunks else np.zeros(0, dtype=TWO_DTYPE)

    def compute_discard(self) -> int:
        """Run with a sink that only counts (end-to-end timing: D2H included)."""
        n_total = [0]

        def _sink(user, ptr, n):
            n_total[0] += int(n)
            return 0

        cb = SINK_FN(_sink)
        self._check(self._L.twkb_compute(self._ctx, cb, None))
        return n_total[0]

    def compute_resident(self):
        self._check(self._L.twkb_compute_resident(self._ctx))

    def compute_decay(self, window_bp: int, n_bins: int):
        """LD decay over distance (two_reader::Decay) reduced on the device from the resident records:
        returns (sum_r2[n_bins], count[n_bins])."""
        n = max(int(n_bins), 1)
        sums = np.zeros(n, dtype=np.float64)
        cnt = np.zeros(n, dtype=np.uint64)
        self._check(self._L.twkb_compute_decay(self._ctx, window_bp, n_bins, sums.ctypes.data, cnt.ctypes.data))
        return sums, cnt

    def compute_aggregate(self, field: str, xbins: int, ybins: int, contig_n_bases):
        """`tomahawk aggregate` from the device-resident records (two_reader::Aggregate): returns (bins [xbins, ybins] of
        AGG_BIN_DTYPE, layout dict, rid_offsets dict). Reduce with aggregate_reduce()."""
        if field.lower() not in AGG_FIELDS:
            raise TwkbError(-1, f'Unknown aggregation function "{field}"...')
        nb = np.ascontiguousarray(contig_n_bases, dtype=np.int64)
        bins = np.zeros((xbins, ybins), dtype=AGG_BIN_DTYPE)
        layout = np.zeros(1, dtype=AGG_LAYOUT_DTYPE)
        off, cmin, cmax = np.zeros(len(nb), np.uint64), np.zeros(len(nb), np.uint32), np.zeros(len(nb), np.uint32)
        self._check(self._L.twkb_compute_aggregate(self._ctx, AGG_FIELDS[field.lower()], xbins, ybins, nb.ctypes.data, len(nb), bins.ctypes.data,
                                                   layout.ctypes.data, off.ctypes.data, cmin.ctypes.data, cmax.ctypes.data))
        lay = {k: int(layout[0][k]) for k in ("range", "bpx", "bpy", "n_contigs_set", "n_records")}
        return bins, lay, {"range": off, "min": cmin, "max": cmax}

    def stats(self) -> Stats:
        s = Stats()
        self._check(self._L.twkb_get_stats(self._ctx, ctypes.byref(s)))
        return s

    def debug_candidates(self, screen_off: bool = True) -> np.ndarray:
        n = ctypes.c_uint64(0)
        self._check(self._L.twkb_debug_candidates(self._ctx, int(screen_off), None, 0, ctypes.byref(n)))
        out = np.zeros(int(n.value), dtype=CAND_DTYPE)
        if n.value:
            self._check(self._L.twkb_debug_candidates(self._ctx, int(screen_off), out.ctypes.data, n.value, ctypes.byref(n)))
        return out[: int(n.value)]


class twk_ld:
    """Mirror of the reference's ``twk_ld`` (include/ld.h:40-69): ``Compute(settings)``
    reads ``settings.in`` (.twk), writes ``settings.out`` (.two) and returns a bool,
    printing errors to stderr like the reference does."""

    def __init__(self):
        self.last_stats = None

    def Compute(self, settings: Settings, in_path: str, out_path: str, ival_strings=()) -> bool:
        import sys

        L = lib()
        st = Stats()
        err = ctypes.create_string_buffer(1024)
        rc = L.twkb_calc_file_intervals(ctypes.byref(settings), in_path.encode(), out_path.encode(),
                                        _c_strings(ival_strings), len(ival_strings), ctypes.byref(st), err, 1024)
        self.last_stats = st
        if rc != 0:
            print(f"[ERROR] {err.value.decode()}", file=sys.stderr)
            return False
        return True
