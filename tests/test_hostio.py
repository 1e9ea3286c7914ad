"""CPU tests of the host-side formats: the .twk reader against the packer, and the
.two writer against an independent parser and the reference's own `view`."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf
from tests.helpers import load_golden


@pytest.mark.parametrize("kw", [
    dict(n_samples=2504, n_variants=1203, seed=1),
    dict(n_samples=333, n_variants=777, seed=2, missing_rate=0.07),
    dict(n_samples=31, n_variants=40, seed=3, missing_rate=0.2),
    dict(n_samples=64, n_variants=1, seed=4),
])
def test_twk_reader_roundtrip(kw, tmpdir_repo):
    s = tf.synth_genotypes(**kw)
    path = os.path.join(tmpdir_repo, "rt.twk")
    n_blocks = tf.write_twk(path, s)
    f = tb.TwkFile(path, n_threads=3)
    assert (f.n_samples, f.n_variants, f.n_blocks) == (s.n_samples, s.n_variants, n_blocks)
    data, mask, meta = f.matrix()
    want_data, want_mask = tf.pack_bits(s)
    assert f.stride == want_data.shape[1]
    assert np.array_equal(data, want_data)
    assert (mask is None) == (want_mask is None)
    if mask is not None:
        assert np.array_equal(mask, want_mask)
    want_meta = lc.variant_meta(s)
    for k in ("rid", "pos", "ac", "an", "hwe", "gt_missing", "gt_phase"):
        assert np.array_equal(meta[k], want_meta[k]), k
    f.close()


def test_twk_reader_multi_contig_blocks(tmpdir_repo):
    s = tf.synth_genotypes(100, 900, seed=5)
    s.rid[600:] = 1
    s.pos[600:] = (np.arange(300) * 100).astype(np.uint32)
    path = os.path.join(tmpdir_repo, "mc.twk")
    nb = tf.write_twk(path, s, contigs=[("1", 10**6), ("2", 10**6)])
    assert nb == 3  # 500 + 100 | 300: one contig per block
    f = tb.TwkFile(path)
    _, _, meta = f.matrix()
    assert np.array_equal(meta["rid"], s.rid) and np.array_equal(meta["pos"], s.pos)


def test_twk_reader_rejects_garbage(tmpdir_repo):
    p = os.path.join(tmpdir_repo, "bad.twk")
    open(p, "wb").write(b"NOTATWK" * 20)
    with pytest.raises(tb.TwkbError):
        tb.TwkFile(p)
    with pytest.raises(tb.TwkbError):
        tb.TwkFile(os.path.join(tmpdir_repo, "does_not_exist.twk"))


def test_two_writer_blocks_index_and_reverse_copies(tmpdir_repo):
    s, recs, prm, pairs, _ = load_golden("phased_r0")
    twk_path = os.path.join(tmpdir_repo, "w.twk")
    tf.write_twk(twk_path, s)
    twk = tb.TwkFile(twk_path)
    out = os.path.join(tmpdir_repo, "w.two")
    w = tb.TwoWriter(out, twk, "pytest", c_level=1, b_size=1000)
    w.add(recs[:2500])
    w.add(recs[2500:])
    w.close()
    back = tf.read_two(out)
    assert len(back) == 2 * len(recs)
    fwd = tf.canonical(back, forward_only=True)
    assert np.array_equal(fwd.view(np.uint8), tf.canonical(recs).view(np.uint8))
    # reverse copies: only (rid, pos) swapped (reference ld_engine.cpp:1292-1298)
    posA, posB = back["packA"] >> 2, back["packB"] >> 2
    rev = back[posA > posB]
    rev_sw = rev.copy()
    rev_sw["packA"], rev_sw["packB"] = rev["packB"], rev["packA"]
    rev_sw["ridA"], rev_sw["ridB"] = rev["ridB"], rev["ridA"]
    assert np.array_equal(tf.canonical(rev_sw).view(np.uint8), tf.canonical(recs).view(np.uint8))


@pytest.mark.skipif(not os.path.exists(lc.REF_VIEW), reason="oracle/_ref/tomahawk_view not built")
def test_reference_view_reads_our_two_file(tmpdir_repo):
    s, recs, prm, pairs, _ = load_golden("phased_r01")
    twk_path = os.path.join(tmpdir_repo, "v.twk")
    tf.write_twk(twk_path, s)
    twk = tb.TwkFile(twk_path)
    out = os.path.join(tmpdir_repo, "v.two")
    w = tb.TwoWriter(out, twk, "pytest")
    w.add(recs)
    w.close()
    r = subprocess.run([lc.REF_VIEW, "view", "-i", out, "-H"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln and not ln.startswith("#") and not ln.startswith("FLAG")]
    assert len(lines) == 2 * len(recs)
    first = lines[0].split("\t")
    assert len(first) == 16
