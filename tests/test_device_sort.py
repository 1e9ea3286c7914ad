"""Device-side sort of the resident records ("a GPU/CPU sorter equivalent to two_reader::Sort so output is
queryable"): twkb_compute_sorted orders forward + reverse records by twk1_two_t::operator< (lib/core.cpp:458-468) on the GPU,
SortedTwoWriter writes them as a sorted, indexed .two -- the file the reference gets from `calc` followed by `sort`."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf
from tests.test_hostio import _two_contigs, _unsorted_two

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def with_reverse(fwd):
    rev = fwd.copy()
    rev["ridA"], rev["ridB"] = fwd["ridB"], fwd["ridA"]
    rev["packA"], rev["packB"] = fwd["packB"], fwd["packA"]
    return np.concatenate([fwd, rev])


def file_order(recs):
    return recs[np.lexsort((recs["packB"], recs["packA"], recs["ridB"].astype(np.int32), recs["ridA"].astype(np.int32)))]


def strip(e):
    return (e[0], e[1], e[2], e[3], e[4], e[8])   # block index entry without the file offsets / compressed size


def test_sorted_writer_equals_the_sorter_and_rejects_disorder(tmpdir_repo):
    """Host side only: records in file order -> SortedTwoWriter == twkb_two_sort of the unsorted file."""
    s = _two_contigs(300, 1400, 900, seed=4)
    src, n = _unsorted_two(tmpdir_repo, "dsrt_src", s, dict(force_phased=1, minR2=0.02), contigs=[("1", 10**6), ("2", 10**6)])
    want = os.path.join(tmpdir_repo, "dsrt_want.two")
    assert tb.sort_two(src, want, c_level=1, n_threads=2) == n
    recs = file_order(tf.read_two(src))
    twk = tb.TwkFile(os.path.join(tmpdir_repo, "dsrt_src.twk"))
    got = os.path.join(tmpdir_repo, "dsrt_got.two")
    w = tb.SortedTwoWriter(got, twk, "pytest", c_level=1, n_threads=3)
    for k in range(0, n, 7001):                    # arbitrary chunking of the stream
        w.add(recs[k:k + 7001])
    w.close()
    assert np.array_equal(tf.read_two(got).view(np.uint8), tf.read_two(want).view(np.uint8))
    (sa, ea, ma), (sb, eb, mb) = tf.read_two_index(want), tf.read_two_index(got)
    assert sa == sb == 2
    assert [strip(e) for e in ea] == [strip(e) for e in eb]
    assert [(m[0], m[1], m[2], m[3], m[6]) for m in ma] == [(m[0], m[1], m[2], m[3], m[6]) for m in mb]
    # a record that sorts before its predecessor is refused, and an unfinished file does not stay behind
    bad = os.path.join(tmpdir_repo, "dsrt_bad.two")
    w = tb.SortedTwoWriter(bad, twk, "pytest")
    w.add(recs[:10])
    with pytest.raises(tb.TwkbError):
        w.add(recs[5:6])
    del w
    twk.close()


@pytest.mark.skipif(not os.path.exists(lc.REF_VIEW), reason="oracle/_ref/tomahawk_view not built")
def test_reference_view_seeks_in_the_sorted_writer_file(tmpdir_repo):
    s = _two_contigs(300, 1400, 900, seed=4)
    src, n = _unsorted_two(tmpdir_repo, "dsrt_v", s, dict(force_phased=1, minR2=0.02), contigs=[("1", 10**6), ("2", 10**6)])
    want = os.path.join(tmpdir_repo, "dsrt_v_want.two")
    tb.sort_two(src, want, c_level=1, n_threads=2)
    twk = tb.TwkFile(os.path.join(tmpdir_repo, "dsrt_v.twk"))
    got = os.path.join(tmpdir_repo, "dsrt_v_got.two")
    w = tb.SortedTwoWriter(got, twk, "pytest")
    w.add(file_order(tf.read_two(src)))
    w.close()
    twk.close()
    for q in (["-I", "1:20000-30000"], ["-I", "2:1000-9000"], []):
        outs = [subprocess.run([lc.REF_VIEW, "view", "-i", f, "-H"] + q, capture_output=True, text=True).stdout for f in (want, got)]
        assert outs[0] == outs[1] and len(outs[0]) > 0, q


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("skw,prm,split", [
    (dict(n_samples=2504, n_variants=3000, seed=71), dict(force_phased=1, minR2=0.05), 0),
    (dict(n_samples=300, n_variants=1400, seed=4), dict(force_phased=1, minR2=0.02), 900),                  # two contigs: rid digits take part
    (dict(n_samples=400, n_variants=760, seed=72), dict(force_phased=1, minR2=0.0), 0),                    # 288,420 pairs -> 5 gather chunks
    (dict(n_samples=300, n_variants=1500, seed=73, missing_rate=0.03), dict(minR2=0.1), 0),                # auto mode: two passes collected
])
def test_device_sorted_records_equal_host_sorted_records(skw, prm, split):
    s = tf.synth_genotypes(**skw)
    if split:
        s.rid[split:] = 1
        s.pos[split:] = (np.arange(s.n_variants - split) * 100).astype(np.uint32)
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(**prm)
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    plain = eng.compute()
    got = eng.compute_sorted()
    st = eng.stats()
    eng.close()
    assert len(plain) > 1000 and len(got) == 2 * len(plain) and st.records_out == len(plain)
    assert np.array_equal(got.view(np.uint8), file_order(with_reverse(plain)).view(np.uint8))


@pytest.mark.gpu
def test_cli_sorted_output_equals_calc_then_sort(tmpdir_repo):
    """twkb_calc calc --sorted == twkb_calc calc + twkb_sort: same records in the same order, same sorted-state index; the
    reference's view answers interval queries on it."""
    s = _two_contigs(300, 1400, 900, seed=4)
    twk = os.path.join(tmpdir_repo, "dsrt_cli.twk")
    tf.write_twk(twk, s, contigs=[("1", 10**6), ("2", 10**6)])
    exe = os.path.join(ROOT, "tomahawk_b200", "twkb_calc")
    plain, direct, sorted_ = (os.path.join(tmpdir_repo, f"dsrt_cli_{k}.two") for k in ("plain", "direct", "sorted"))
    for out, extra in ((plain, []), (direct, ["--sorted"])):
        r = subprocess.run([exe, "calc", "-p", "-r", "0.02", "-i", twk, "-o", out, "-t", "4", *extra], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
    n = tb.sort_two(plain, sorted_, c_level=1, n_threads=4)
    a, b = tf.read_two(sorted_), tf.read_two(direct)
    assert n == len(b) > 2000
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    (sa, ea, ma), (sb, eb, mb) = tf.read_two_index(sorted_), tf.read_two_index(direct)
    assert sa == sb == 2 and [strip(e) for e in ea] == [strip(e) for e in eb]
    assert [(m[0], m[1], m[2], m[3], m[6]) for m in ma] == [(m[0], m[1], m[2], m[3], m[6]) for m in mb]
    r = subprocess.run([exe, "calc", "-p", "-i", twk, "-o", direct, "--sorted", "-g", "0,0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "one device" in r.stderr
    if os.path.exists(lc.REF_VIEW):
        for q in (["-I", "1:20000-30000"], ["-I", "2:1000-9000"]):
            outs = [subprocess.run([lc.REF_VIEW, "view", "-i", f, "-H"] + q, capture_output=True, text=True).stdout for f in (sorted_, direct)]
            assert outs[0] == outs[1] and len(outs[0]) > 0, q
