"""Downstream consumer fed from the device-resident records: LD decay over distance, the reference's
two_reader::Decay (lib/two_reader.cpp:424-475), reduced on the GPU without writing / re-reading a .two file."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DECAY = os.path.join(ROOT, "oracle", "_ref", "tomahawk_decay")


def decay_of_records(recs, window_bp, n_bins):
    """numpy restatement of two_reader::Decay over forward + reverse records."""
    width = window_bp // n_bins
    posA, posB = (recs["packA"] >> 2).astype(np.int64), (recs["packB"] >> 2).astype(np.int64)
    keep = (recs["ridA"] == recs["ridB"]) & (posA < posB)
    b = np.minimum((posB[keep] - posA[keep]) // width, n_bins - 1)
    sums = np.bincount(b, weights=recs["R2"][keep], minlength=n_bins)
    cnt = np.bincount(b, minlength=n_bins)
    return sums, cnt


def test_decay_rejects_bad_arguments_without_a_gpu():
    L = tb.lib()
    assert hasattr(L, "twkb_compute_decay")
    assert L.twkb_compute_decay(None, 1000, 10, None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("skw,prm,window,bins", [
    (dict(n_samples=2504, n_variants=3000, seed=51), dict(force_phased=1, minR2=0.05), 100_000, 50),
    (dict(n_samples=300, n_variants=1500, seed=52, missing_rate=0.03), dict(forced_unphased=1, minR2=0.1), 40_000, 7),
    (dict(n_samples=500, n_variants=700, seed=53), dict(force_phased=1, minR2=0.0), 2_000_000, 3000),   # > 1,024 bins: global atomics
])
def test_device_decay_equals_decay_of_the_records(skw, prm, window, bins):
    s = tf.synth_genotypes(**skw)
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(**prm)
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    recs = eng.compute()
    sums, cnt = eng.compute_decay(window, bins)
    st = eng.stats()
    eng.close()
    want_sums, want_cnt = decay_of_records(recs, window, bins)
    assert st.bytes_d2h == 0 and st.records_out == len(recs)         # the records never left the device
    assert np.array_equal(cnt.astype(np.int64), want_cnt)
    np.testing.assert_allclose(sums, want_sums, rtol=1e-11, atol=1e-12)
    assert cnt.sum() == len(recs)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_DECAY), reason="oracle/_ref/tomahawk_decay not built")
def test_device_decay_equals_reference_decay_of_the_reference_file(tmpdir_repo):
    """The reference's calc writes a .two, the reference's Decay reads it back; the GPU path gives the same table
    (Mean to 1e-9, Frequency exactly) straight from the device."""
    s = tf.synth_genotypes(600, 2500, seed=54)
    twk = os.path.join(tmpdir_repo, "decay.twk")
    tf.write_twk(twk, s)
    lc.run_reference_calc(twk, os.path.join(tmpdir_repo, "decay_ref"), ["-p", "-r", "0.05"], threads=4)
    r = subprocess.run([REF_DECAY, "decay", "-i", os.path.join(tmpdir_repo, "decay_ref.two"), "-w", "150000", "-b", "30"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    rows = [ln.split("\t") for ln in r.stdout.strip().splitlines()[1:]]
    assert len(rows) == 30
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(force_phased=1, minR2=0.05)
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    sums, cnt = eng.compute_decay(150_000, 30)
    eng.close()
    for b, (lo, hi, mean, freq) in enumerate(rows):
        assert (int(lo), int(hi)) == (b * 5000, (b + 1) * 5000)
        assert int(freq) == int(cnt[b])
        assert abs(float(mean) - sums[b] / max(int(cnt[b]), 1)) <= 1e-5 * max(float(mean), 1e-12) + 1e-12   # printed with 6 digits
