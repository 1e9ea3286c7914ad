"""Tests of the measurement tooling (libtwkb_tools.so): the device-side synthetic generator that feeds the
biobank-scale bench configurations, and the peak probes. GPU tests; the CPU part only checks the exports."""
import ctypes
import os

import numpy as np
import pytest

import tomahawk_b200 as tb
from tomahawk_b200 import tools


def test_tools_library_loads_and_exports():
    L = tools.lib()
    for name in ("twkb_tools_synth", "twkb_tools_popc_rate", "twkb_tools_fp4_gemm", "twkb_tools_last_error"):
        assert hasattr(L, name)


def test_rows_to_alleles_roundtrip():
    from tomahawk_b200 import synth
    s = synth.synth_genotypes(70, 40, seed=5, missing_rate=0.1)
    data, mask = synth.pack_bits(s)
    assert np.array_equal(tools.rows_to_alleles(data, mask, 70), s.alleles)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(n_samples=2504, n_variants=3000), dict(n_samples=333, n_variants=700, missing_rate=0.05),
                                dict(n_samples=40000, n_variants=300, rare_fraction=0.8)])
def test_device_generator_is_consistent(kw):
    import torch

    n, m = kw["n_samples"], kw["n_variants"]
    data, mask, meta = tools.synth_device(seed=9, **kw)
    d = data.cpu().numpy().view(np.uint64)
    mk = mask.cpu().numpy().view(np.uint64) if mask is not None else None
    H = 2 * n
    bits = np.unpackbits(d.view(np.uint8), axis=1, bitorder="little")
    assert not bits[:, H:].any()                                  # padding bits stay zero
    assert np.array_equal(bits[:, :H].sum(1), meta["ac"])         # metadata matches the rows
    assert meta["ac"].min() >= 1
    if mk is not None:
        mb = np.unpackbits(mk.view(np.uint8), axis=1, bitorder="little")
        assert np.array_equal(mb[:, :H].sum(1), meta["an"])
        assert not (bits & mb).any()                              # data bits are 0 where missing (lib/core.cpp:377-380)
        assert np.array_equal(mb[:, 0:H:2], mb[:, 1:H:2])          # a sample is missing as a whole
        assert 0.03 < mb[:, :H].mean() < 0.07
        assert np.all((meta["ac"] + meta["an"]) <= H)
    assert np.array_equal(meta["pos"], np.arange(m) * 100)
    # any slice of the same (seed, shape) is the same data: what lets every rank generate its own rows
    d2, m2, meta2 = tools.synth_device(seed=9, first=m // 3, n_rows=m // 2, **kw)
    assert np.array_equal(d2.cpu().numpy().view(np.uint64), d[m // 3:m // 3 + m // 2])
    assert np.array_equal(meta2["ac"], meta["ac"][m // 3:m // 3 + m // 2])
    if kw.get("rare_fraction"):
        assert (meta["ac"] < 0.01 * H).mean() > 0.6


@pytest.mark.gpu
def test_device_generated_matrix_through_engine_matches_oracle():
    """Rows generated on the device, handed over by pointer (twkb_load_matrix_device), against the oracle on
    the same genotypes: the path bench.py uses for the configurations numpy cannot generate."""
    from oracle import ldcore as lc
    from oracle import twk_format as tf
    from tests.helpers import assert_records_bitexact

    n, m = 1200, 900
    data, mask, meta = tools.synth_device(n, m, seed=4)
    al = tools.rows_to_alleles(data.cpu().numpy().view(np.uint64), None, n)
    s = tf.Synth(alleles=al, pos=meta["pos"].copy(), rid=meta["rid"].copy(), n_samples=n)
    ref, visited = lc.calc(s, lc.default_params(force_phased=1, minR2=0.1))
    assert len(ref) > 200                                         # the generator really produces LD
    eng = tb.Engine(force_phased=1, minR2=0.1)
    eng.load_device(n, m, data.data_ptr(), None, data.shape[1], meta)
    got = eng.compute()
    assert eng.stats().pairs_visited == visited
    assert_records_bitexact(got, ref, p_rtol=1e-9)
    eng.close()


@pytest.mark.gpu
def test_peak_probes():
    rate, per_clk_sm, mhz = tools.popc_rate()
    assert 1e12 < rate < 2e13 and 4 < per_clk_sm < 70
    try:
        burst, sustained = tools.fp4_gemm_tflops(4096, 0.3)
    except RuntimeError as e:
        pytest.skip(f"cuBLASLt block-scaled e2m1 GEMM unavailable: {e}")
    assert 500 < burst < 20000 and 500 < sustained < 20000   # (back-to-back launches hide the launch gap: sustained may exceed burst at this size)
