"""ctypes binding of oracle/libldcore.so (the CPU restatement in ldcore.c) and
helpers to run the compiled reference binaries under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by tomahawk_b200/.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import numpy as np

from . import twk_format as tf

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_CALC = os.path.join(REF_DIR, "tomahawk_calc")
REF_VIEW = os.path.join(REF_DIR, "tomahawk_view")
REF_SORT = os.path.join(REF_DIR, "tomahawk_sort")
REF_FISHER = os.path.join(REF_DIR, "libref_fisher.so")

VARIANT_DTYPE = np.dtype(
    [("rid", "<u4"), ("pos", "<u4"), ("ac", "<u4"), ("an", "<u4"), ("hwe", "<f8"), ("gt_missing", "u1"), ("gt_phase", "u1"), ("pad", "u1", (6,))]
)
assert VARIANT_DTYPE.itemsize == 32


class Params(ctypes.Structure):
    """ld_params of ldcore.c: the twk_ld_settings fields calc reads
    (include/core.h:909-924; defaults lib/core.cpp:297-306)."""

    _fields_ = [
        ("minP", ctypes.c_double),
        ("minR2", ctypes.c_double),
        ("maxR2", ctypes.c_double),
        ("minDprime", ctypes.c_double),
        ("maxDprime", ctypes.c_double),
        ("force_phased", ctypes.c_int32),
        ("forced_unphased", ctypes.c_int32),
        ("window", ctypes.c_int32),
        ("l_window", ctypes.c_int32),
        ("emulate_quirks", ctypes.c_int32),
        ("block_size", ctypes.c_int32),
        ("skip_min_cell_rule", ctypes.c_int32),
        ("bitmaps", ctypes.c_int32),
    ]


def default_params(**kw) -> Params:
    p = Params(1.0, 0.1, 100.0, 0.0, 100.0, 0, 0, 0, 1000000, 1, 500, 0, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libldcore.so"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libldcore.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.ldcore_fisher.restype = ctypes.c_double
        L.ldcore_fisher.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
        L.ldcore_calc.restype = ctypes.c_int64
        L.ldcore_calc.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32,
            ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64),
        ]
        L.ldcore_phased_stats.restype = ctypes.c_int
        L.ldcore_phased_stats.argtypes = [ctypes.c_uint64] * 4 + [ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ldcore_unphased_stats.restype = ctypes.c_int
        L.ldcore_unphased_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Params), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


STATS_DTYPE = np.dtype(
    [("flags", "<u2"), ("pad", "u1", (6,)), ("cnt", "<f8", (4,)), ("D", "<f8"), ("Dprime", "<f8"), ("R", "<f8"), ("R2", "<f8"), ("P", "<f8"), ("chi_fisher", "<f8"), ("chi_model", "<f8")]
)
assert STATS_DTYPE.itemsize == 96


def fisher(n11, n12, n21, n22) -> float:
    return lib().ldcore_fisher(int(n11), int(n12), int(n21), int(n22), None, None)


def phased_stats(c0, c1, c4, c5, params=None, a=None, b=None):
    """counts in the reference's slot order (REFREF, slot1, slot4, ALTALT)."""
    params = params or default_params()
    va = np.zeros(1, VARIANT_DTYPE) if a is None else a
    vb = np.zeros(1, VARIANT_DTYPE) if b is None else b
    if a is None:
        va["hwe"] = 1.0
        va["ac"] = c1 + c5
    if b is None:
        vb["hwe"] = 1.0
        vb["ac"] = c4 + c5
    s = np.zeros(1, STATS_DTYPE)
    ok = lib().ldcore_phased_stats(int(c0), int(c1), int(c4), int(c5), ctypes.byref(params), va.ctypes.data, vb.ctypes.data, s.ctypes.data)
    return bool(ok), s[0]


def unphased_stats(table, params=None):
    params = params or default_params()
    t = np.ascontiguousarray(np.asarray(table, dtype=np.uint64).reshape(3, 3))
    va = np.zeros(1, VARIANT_DTYPE)
    vb = np.zeros(1, VARIANT_DTYPE)
    va["hwe"] = vb["hwe"] = 1.0
    va["ac"] = vb["ac"] = 100
    s = np.zeros(1, STATS_DTYPE)
    ok = lib().ldcore_unphased_stats(t.ctypes.data, ctypes.byref(params), va.ctypes.data, vb.ctypes.data, s.ctypes.data)
    return bool(ok), s[0]


from tomahawk_b200.synth import variant_meta  # noqa: E402,F401


def calc(s: tf.Synth, params: Params, cap: int | None = None):
    """Run the CPU restatement over all pairs -> (records[TWO_DTYPE], pairs_visited)."""
    data, mask = tf.pack_bits(s)
    meta = variant_meta(s)
    M = s.n_variants
    if cap is None:
        cap = M * (M - 1) // 2 + 1
    out = np.zeros(cap, tf.TWO_DTYPE)
    visited = ctypes.c_uint64(0)
    n = lib().ldcore_calc(
        data.ctypes.data, mask.ctypes.data if mask is not None else None, data.shape[1], s.n_samples, M,
        meta.ctypes.data, ctypes.byref(params), out.ctypes.data, cap, ctypes.byref(visited),
    )
    if n < 0:
        raise RuntimeError("ldcore_calc: output capacity too small")
    return out[:n].copy(), int(visited.value)


def scalc_select(s: tf.Synth, contig_rid: int, start: int, stop: int, l_surrounding: int, emulate_quirks: bool = True):
    """Variant selection of `scalc` (twk_ld_impl::LoadTargetSingle, lib/ld/ld.cpp:123-255): 1-based inclusive matching of
    pos + 1 against the target interval [start, stop] and the flanks [max(start - L, 0), max(start - 1, 0)], [stop, stop + L].
    Returns (Synth ordered [targets | neighbours], n_targets). "chr:pos" parses to [pos, pos + 1] (lib/intervals.cpp:113-114)."""
    p1 = s.pos.astype(np.int64) + 1
    on = s.rid == contig_rid
    t = on & (p1 >= start) & (p1 <= stop)
    left = on & (p1 >= max(start - l_surrounding, 0)) & (p1 <= max(start - 1, 0))
    right = on & (p1 >= stop) & (p1 <= stop + l_surrounding)
    others = np.flatnonzero((left | right) & ~t)
    if emulate_quirks:  # neighbours are gathered in blocks of 100; the last, partial block is dropped (ld.cpp:193-195, :241)
        others = others[: len(others) // 100 * 100]
    order = np.concatenate([np.flatnonzero(t), others])
    sub = tf.Synth(alleles=s.alleles[order], pos=s.pos[order].copy(), rid=s.rid[order].copy(), n_samples=s.n_samples)
    return sub, int(t.sum())


def calc_single(s: tf.Synth, params: Params, n_targets: int):
    """CPU restatement of `scalc` over a matrix ordered [targets | neighbours] -> (records, pairs_visited)."""
    L = lib()
    L.ldcore_calc_single.restype = ctypes.c_int64
    L.ldcore_calc_single.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.POINTER(Params), ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64)]
    data, mask = tf.pack_bits(s)
    meta = variant_meta(s)
    cap = n_targets * s.n_variants + 1
    out = np.zeros(cap, tf.TWO_DTYPE)
    visited = ctypes.c_uint64(0)
    n = L.ldcore_calc_single(data.ctypes.data, mask.ctypes.data if mask is not None else None, data.shape[1], s.n_samples, s.n_variants,
                             meta.ctypes.data, ctypes.byref(params), n_targets, out.ctypes.data, cap, ctypes.byref(visited))
    if n < 0:
        raise RuntimeError("ldcore_calc_single: output capacity too small")
    return out[:n].copy(), int(visited.value)


REF_SCALC = os.path.join(REF_DIR, "tomahawk_scalc")


def run_reference_scalc(twk_path: str, out_prefix: str, args: list[str], threads: int = 2, timeout=None):
    """`tomahawk scalc` of the reference itself (oracle/_ref/tomahawk_scalc)."""
    cmd = [REF_SCALC, "scalc", "-i", twk_path, "-o", out_prefix, "-t", str(threads)] + list(args)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference scalc failed ({r.returncode}): {r.stderr[-2000:]}")
    return {"stderr": r.stderr, "cmd": cmd}


# ---------------------------------------------------------- compiled reference
def have_reference() -> bool:
    return os.path.exists(REF_CALC) and os.access(REF_CALC, os.X_OK)


# The reference's progress thread and its final summary race on stderr, so short runs
# interleave the two lines; the patterns tolerate padding and fall back gracefully.
_PROGRESS_RE = re.compile(r"([\d,]+)\s+variants/s and\s+([\d,]+)")
_FINISHED_RE = re.compile(r"Finished in\s+(\S+)\s*\. Variants:\s*([\d,]+)\s*, genotypes:\s*\D*([\d,]+)\s*, output:\s*([\d,]+)")
_PERFORMING_RE = re.compile(r"Performing: ([\d,]+) variant comparisons")


def run_reference_calc(twk_path: str, out_prefix: str, args: list[str], threads: int | None = None, timeout=None):
    """`tomahawk calc` of the reference itself. Returns dict(records, pairs, pairs_per_s, stderr)."""
    threads = threads or os.cpu_count() or 1
    cmd = [REF_CALC, "calc", "-i", twk_path, "-o", out_prefix, "-t", str(threads)] + list(args)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference calc failed ({r.returncode}): {r.stderr[-2000:]}")
    info = {"stderr": r.stderr, "cmd": cmd}
    m = _PROGRESS_RE.search(r.stderr)
    if m:
        info["pairs_per_s"] = int(m.group(1).replace(",", ""))
        info["genotypes_per_s"] = int(m.group(2).replace(",", ""))
    m = _FINISHED_RE.search(r.stderr)
    if m:
        info["pairs"] = int(m.group(2).replace(",", ""))
        info["n_out"] = int(m.group(4).replace(",", ""))
        info["elapsed_str"] = m.group(1)
    m = _PERFORMING_RE.search(r.stderr)
    if m:
        info["planned_pairs"] = int(m.group(1).replace(",", ""))
    return info


_ref_fisher = None


def reference_fisher(n11, n12, n21, n22) -> float:
    """The reference's own kt_fisher_exact (lib/fisher_math.cpp:231) via oracle/_ref/libref_fisher.so."""
    global _ref_fisher
    if _ref_fisher is None:
        L = ctypes.CDLL(REF_FISHER)
        f = getattr(L, "_Z15kt_fisher_exactiiiiPdS_S_")
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_double)] * 3
        _ref_fisher = f
    l, r, t = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _ref_fisher(int(n11), int(n12), int(n21), int(n22), ctypes.byref(l), ctypes.byref(r), ctypes.byref(t))
    return t.value
