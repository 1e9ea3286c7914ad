"""Downstream consumer fed from the device-resident records: `tomahawk aggregate`, the reference's
two_reader::Aggregate (lib/two_reader.cpp:543-853, lib/aggregation.h:127-175), rasterised on the GPU without writing /
re-reading a .two file. The checker is a numpy restatement of the reference's two passes, pinned on the CPU against the
reference's own binary (oracle/_ref/tomahawk_aggregate) reading a .two the reference's calc wrote."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_AGG = os.path.join(ROOT, "oracle", "_ref", "tomahawk_aggregate")
FIELDS = {"r2": lambda r: r["R2"], "r": lambda r: r["R"], "d": lambda r: r["D"], "dprime": lambda r: r["Dprime"], "p": lambda r: r["P"],
          "hets": lambda r: (r["cnt"][:, 1] + r["cnt"][:, 2]) / r["cnt"].sum(axis=1), "alts": lambda r: r["cnt"][:, 3] / r["cnt"].sum(axis=1)}


def with_reverse(fwd):
    rev = fwd.copy()                                  # ld_engine.cpp:1290-1298: (ridA:posA) <-> (ridB:posB), statistics unchanged
    rev["ridA"], rev["ridB"] = fwd["ridB"], fwd["ridA"]
    rev["packA"], rev["packB"] = fwd["packB"], fwd["packA"]
    return np.concatenate([fwd, rev])


def aggregate_of_records(recs, field, xbins, ybins, contig_n_bases, block_sizes=None, emulate_quirks=True):
    """numpy restatement of two_reader::Aggregate over records in FILE order (both orientations present).
    block_sizes: records per .two block -- blocks with fewer than 5 records are skipped (aggregation.h:131,152)."""
    if block_sizes is not None:
        keep = np.repeat(np.asarray(block_sizes) >= 5, block_sizes)
        recs = recs[keep]
    nc = len(contig_n_bases)
    posA, posB = (recs["packA"] >> 2).astype(np.int64), (recs["packB"] >> 2).astype(np.int64)
    cmin, cmax, isset = np.full(nc, 2**32 - 1, np.int64), np.zeros(nc, np.int64), np.zeros(nc, bool)
    for rid, pos in ((recs["ridA"], posA), (recs["ridB"], posB)):                       # FindRangesUnsorted
        np.minimum.at(cmin, rid, pos); np.maximum.at(cmax, rid, pos); isset[rid] = True
    n_set = int(isset.sum()) + (1 if emulate_quirks and isset[0] else 0)                # two_reader.cpp:737-740 counts contig 0 twice
    if n_set == 1:
        omin, omax, span = cmin, cmax, np.where(isset, cmax - cmin + 1, 0)
    else:
        omin, omax = np.zeros(nc, np.int64), np.asarray(contig_n_bases, np.int64) % 2**32
        span = np.where(isset, np.asarray(contig_n_bases, np.int64), 0)
    cum = np.cumsum(span)
    rng = int(cum[-1])
    bpx, bpy = int(np.ceil(np.float32(rng) / np.float32(xbins))), int(np.ceil(np.float32(rng) / np.float32(ybins)))   # :801-802, float arithmetic
    base = cum - (omax - omin)
    ca, cb = base[recs["ridA"]] + (posA - omin[recs["ridA"]]), base[recs["ridB"]] + (posB - omin[recs["ridB"]])
    x, y = np.minimum(ca // bpx, xbins - 1), np.minimum(cb // bpy, ybins - 1)            # the clamp is ours (the reference indexes out of range)
    v = FIELDS[field](recs)
    bins = np.zeros((xbins, ybins), dtype=tb.AGG_BIN_DTYPE)
    np.add.at(bins["n"], (x, y), 1)
    np.add.at(bins["total"], (x, y), v)
    np.add.at(bins["total_squared"], (x, y), v * v)
    np.minimum.at(bins["min"], (x, y), v)                                               # twk_sstats starts min and max at 0
    np.maximum.at(bins["max"], (x, y), v)
    return bins, dict(range=rng, bpx=bpx, bpy=bpy), dict(range=cum, min=omin, max=omax)


def run_reference_aggregate(two, field, red, x, y, cutoff, threads=1):
    r = subprocess.run([REF_AGG, "aggregate", "-i", two, "-f", field, "-r", red, "-x", str(x), "-y", str(y), "-c", str(cutoff), "-t", str(threads)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    lines = r.stdout.strip().splitlines()
    rng, bpx, bpy, n_orig = (int(t) for t in lines[0].split())
    nc = int(lines[1])
    offs = np.array([[int(t) for t in ln.split()] for ln in lines[2:2 + nc]], dtype=np.int64)
    mat = np.array([[float(t) for t in ln.split()] for ln in lines[2 + nc:2 + nc + x]])
    return dict(range=rng, bpx=bpx, bpy=bpy, n_original=n_orig), offs, mat


def two_contigs(s, split):
    s.rid[split:] = 1
    s.pos[split:] -= s.pos[split] - 100
    return s


def test_aggregate_rejects_bad_arguments_without_a_gpu():
    L = tb.lib()
    assert L.twkb_compute_aggregate(None, 0, 10, 10, None, 0, None, None, None, None, None) == -1
    bins = np.zeros((3, 3), dtype=tb.AGG_BIN_DTYPE)
    bins["n"][0, 0], bins["total"][0, 0], bins["total_squared"][0, 0] = 4, 2.0, 1.5
    bins["n"][1, 1], bins["total"][1, 1], bins["total_squared"][1, 1] = 10, 5.0, 4.0
    assert tb.aggregate_reduce(bins, "mean", 5)[0, 0] == 0 and tb.aggregate_reduce(bins, "mean", 5)[1, 1] == 0.5   # GetMean: n < cutoff -> 0
    assert tb.aggregate_reduce(bins, "mean", 0)[1, 1] == 0                                                          # ... and cutoff 0 -> 0 (core.h:958)
    assert tb.aggregate_reduce(bins, "count", 5)[0, 0] == 0 and tb.aggregate_reduce(bins, "n", 3)[0, 0] == 4
    assert abs(tb.aggregate_reduce(bins, "sd", 2)[1, 1] - np.sqrt(4.0 / 10 - 0.25)) < 1e-15
    with pytest.raises(tb.TwkbError):
        tb.aggregate_reduce(bins, "median")


@pytest.mark.skipif(not (os.path.exists(REF_AGG) and lc.have_reference()), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["one_contig_not_first", "first_contig_only", "two_contigs"])
def test_numpy_restatement_equals_reference_aggregate(case, tmpdir_repo):
    """Pins the checker: reference calc -> .two -> reference Aggregate == aggregate_of_records(read_two(.two))."""
    s = tf.synth_genotypes(300, 1400, seed=61)
    contigs = [("1", 400_000), ("2", 300_000)]
    if case == "two_contigs":
        s = two_contigs(s, 800)
    elif case == "one_contig_not_first":
        s.rid[:] = 1
    twk = os.path.join(tmpdir_repo, f"agg_{case}.twk")
    tf.write_twk(twk, s, contigs=contigs)
    lc.run_reference_calc(twk, os.path.join(tmpdir_repo, f"agg_{case}_ref"), ["-p", "-r", "0.05"], threads=2)
    two = os.path.join(tmpdir_repo, f"agg_{case}_ref.two")
    recs = tf.read_two(two)
    sizes = [e[1] for e in tf.read_two_index(two)[1]]
    assert len(recs) > 2000 and sum(sizes) == len(recs)
    for field, red, x, y, cutoff in (("r2", "mean", 40, 40, 5), ("d", "min", 25, 25, 1), ("hets", "sd", 30, 30, 2), ("p", "count", 12, 12, 0), ("dprime", "total", 17, 17, 3)):
        lay, offs, mat = run_reference_aggregate(two, field, red, x, y, cutoff)
        bins, mylay, myoffs = aggregate_of_records(recs, field, x, y, [c[1] for c in contigs], block_sizes=sizes)
        assert (lay["range"], lay["bpx"], lay["bpy"]) == (mylay["range"], mylay["bpx"], mylay["bpy"])
        assert np.array_equal(offs[:, 0], myoffs["range"])
        if case != "two_contigs" and not (case == "first_contig_only"):
            assert np.array_equal(offs[s.rid[0], 1:], [myoffs["min"][s.rid[0]], myoffs["max"][s.rid[0]]])
        np.testing.assert_allclose(mat, tb.aggregate_reduce(bins, red, cutoff), rtol=1e-12, atol=1e-300)
    if case == "first_contig_only":
        assert lay["range"] == 400_000        # the reference counts contig 0 twice: whole-contig coordinates although one contig has data


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("skw,prm,split,field,x,y", [
    (dict(n_samples=2504, n_variants=3000, seed=62), dict(force_phased=1, minR2=0.05), 0, "r2", 64, 64),
    (dict(n_samples=400, n_variants=2200, seed=63), dict(force_phased=1, minR2=0.0), 1300, "d", 33, 21),            # two contigs, x != y, negative values
    (dict(n_samples=300, n_variants=1500, seed=64, missing_rate=0.03), dict(forced_unphased=1, minR2=0.1), 0, "hets", 10, 10),
    (dict(n_samples=500, n_variants=900, seed=65), dict(force_phased=1, minR2=0.02), -1, "p", 1000, 1000),          # only contig 1: data-range coordinates
])
def test_device_aggregate_equals_aggregate_of_the_records(skw, prm, split, field, x, y):
    s = tf.synth_genotypes(**skw)
    if split > 0:
        s = two_contigs(s, split)
    elif split < 0:
        s.rid[:] = 1
    n_bases = [int(s.pos.max()) + 1000, 400_000]
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(**prm)
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    recs = eng.compute()
    bins, lay, offs = eng.compute_aggregate(field, x, y, n_bases)
    st = eng.stats()
    eng.close()
    want, wlay, woffs = aggregate_of_records(with_reverse(recs), field, x, y, n_bases)
    assert st.bytes_d2h == 0 and st.records_out == len(recs)             # the records never left the device
    assert (lay["range"], lay["bpx"], lay["bpy"], lay["n_records"]) == (wlay["range"], wlay["bpx"], wlay["bpy"], 2 * len(recs))
    assert np.array_equal(offs["range"], woffs["range"].astype(np.uint64))
    assert np.array_equal(bins["n"], want["n"]) and bins["n"].sum() == 2 * len(recs)
    assert np.array_equal(bins["min"], want["min"]) and np.array_equal(bins["max"], want["max"])
    np.testing.assert_allclose(bins["total"], want["total"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(bins["total_squared"], want["total_squared"], rtol=1e-11, atol=1e-13)
    if x == y:
        assert np.array_equal(bins["n"], bins["n"].T)                    # forward + reverse copies: a symmetric raster


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_AGG) and lc.have_reference()), reason="oracle/_ref not built")
def test_device_aggregate_equals_reference_aggregate_of_the_reference_file(tmpdir_repo):
    """The reference's calc writes a .two, the reference's Aggregate reads it back; the GPU path gives the same raster
    straight from the device (file read through the .twk reader, contig lengths from its header)."""
    s = two_contigs(tf.synth_genotypes(600, 2500, seed=66), 1500)
    contigs = [("1", 500_000), ("2", 250_000)]
    twk = os.path.join(tmpdir_repo, "agg_gpu.twk")
    tf.write_twk(twk, s, contigs=contigs)
    lc.run_reference_calc(twk, os.path.join(tmpdir_repo, "agg_gpu_ref"), ["-p", "-r", "0.05"], threads=4)
    two = os.path.join(tmpdir_repo, "agg_gpu_ref.two")
    sizes = [e[1] for e in tf.read_two_index(two)[1]]
    f = tb.TwkFile(twk)
    n_bases = f.contigs()
    assert list(n_bases) == [500_000, 250_000]
    data, mask, meta = f.matrix()
    eng = tb.Engine(force_phased=1, minR2=0.05)
    eng.load(f.n_samples, data, mask, meta)
    for field, red, cutoff in (("r2", "mean", 5), ("r", "max", 0), ("alts", "count", 1)):
        lay, offs, mat = run_reference_aggregate(two, field, red, 50, 50, cutoff, threads=3)
        bins, mylay, _ = eng.compute_aggregate(field, 50, 50, n_bases)
        assert (lay["range"], lay["bpx"], lay["bpy"]) == (mylay["range"], mylay["bpx"], mylay["bpy"])
        got = tb.aggregate_reduce(bins, red, cutoff)
        if min(sizes) >= 5:                      # otherwise the reference dropped the records of its short blocks
            assert lay["n_original"] == mylay["n_records"]
            np.testing.assert_allclose(got, mat, rtol=1e-9, atol=1e-300)
        else:
            assert np.mean(np.isclose(got, mat, rtol=1e-9, atol=1e-300)) > 0.99
    eng.close()
    f.close()
