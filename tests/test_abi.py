"""CPU tests of the C-ABI boundary: the library loads, exports every symbol that
include/twkb.h declares, mirrors the reference's settings defaults, and refuses to
run (loudly) when no B200 is present -- there is no CPU fallback to fall into."""
import ctypes
import os
import re

import numpy as np
import pytest

import tomahawk_b200 as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "twkb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(twkb_[a-z0-9_]+)\s*\(", src)) - {"twkb_sink_fn"})


def test_every_declared_symbol_is_exported():
    L = tb.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libtwkb.so does not export {n}"
    assert set(tb.EXPORTS) == set(names)


def test_settings_defaults_match_reference():
    # reference lib/core.cpp:297-306
    s = tb.default_settings()
    assert (s.minP, s.minR2, s.maxR2, s.minDprime, s.maxDprime) == (1.0, 0.1, 100.0, 0.0, 100.0)
    assert (s.c_level, s.bl_size, s.b_size, s.l_window, s.l_surrounding) == (1, 500, 10000, 1000000, 500000)
    assert (s.n_chunks, s.c_chunk, s.window, s.force_phased, s.forced_unphased) == (1, 0, 0, 0, 0)
    assert s.twk_block_size == 500


def test_struct_layouts():
    assert ctypes.sizeof(tb.Settings) == 8 + 8 * 4 + 5 * 8 + 5 * 4 + 5 * 4
    assert tb.VARIANT_DTYPE.itemsize == 32
    assert tb.TWO_DTYPE.itemsize == tb.RECORD_BYTES == 106
    assert tb.CAND_DTYPE.itemsize == 48


def test_invalid_settings_are_rejected():
    L = tb.lib()
    ctx = ctypes.c_void_p()
    for kw in (dict(single=1, n_chunks=3), dict(single=1, window=1), dict(single=1, part_count=2, part_index=1),
               dict(force_phased=1, forced_unphased=1), dict(window=1, n_chunks=3), dict(n_chunks=3, c_chunk=5),
               dict(minR2=1.5), dict(part_count=2, part_index=2)):
        s = tb.default_settings(**kw)
        assert L.twkb_create(ctypes.byref(s), ctypes.byref(ctx)) == -1, kw
        assert len(L.twkb_last_error(None)) > 0


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_device_fails_loudly_and_never_falls_back():
    with pytest.raises(tb.TwkbError) as e:
        tb.Engine(force_phased=1)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_null_arguments():
    L = tb.lib()
    assert L.twkb_compute(None, tb.SINK_FN(lambda *a: 0), None) == -1
    assert L.twkb_get_stats(None, None) == -1
    assert L.twkb_version() >= 100


def test_missing_library_message(monkeypatch):
    monkeypatch.setattr(tb, "_lib", None)
    monkeypatch.setattr(tb, "LIB_PATH", "/nonexistent/libtwkb.so")
    with pytest.raises(ImportError) as e:
        tb.lib()
    assert "no CPU fallback" in str(e.value)
