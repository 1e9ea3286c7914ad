"""Position shards of a -w run ("position-sharding with +-window halo" for a matrix larger than one GPU):
twkb_plan_shards cuts the .twk blocks into consecutive own ranges + halo, a context with settings.shard_blocks computes
only the pairs whose earlier member it owns. CPU tests pin the planner against the reference's window semantics through
the oracle (ldcore.c follows ld_balancing.h:176-233 and ld_engine.cpp:2553-2560); the GPU tests compare the union of the
shards with the whole-matrix run record for record."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def uniform_blocks(s, bs):
    """First variant of every .twk block (blocks of <= bs variants, never across contigs: lib/importer.cpp:196-236) + M."""
    first, v, M = [], 0, s.n_variants
    while v < M:
        e = v + 1
        while e < M and e - v < bs and s.rid[e] == s.rid[v]:
            e += 1
        first.append(v)
        v = e
    return np.array(first + [M], dtype=np.uint32)


def reference_prune(first, pos, w):
    """ld_balancing.h:189-196: first block column the balancer drops from block row bi (positions only, uint32 wrap)."""
    nb = len(first) - 1
    prune = np.full(nb, nb, dtype=np.int64)
    for bi in range(nb):
        last = int(pos[first[bi + 1] - 1])
        for bj in range(bi + 1, nb):
            if (int(pos[first[bj]]) - last) % (1 << 32) > w:
                prune[bi] = bj
                break
    return prune


def sub_synth(s, v0, v1):
    return tf.Synth(alleles=s.alleles[v0:v1], pos=s.pos[v0:v1], rid=s.rid[v0:v1], n_samples=s.n_samples)


@pytest.mark.parametrize("n_shards", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("w", [700, 15_000, 10_000_000])
def test_plan_shards_partitions_blocks_and_halo_covers_the_window(n_shards, w):
    s = tf.synth_genotypes(8, 4100, seed=11)
    s.rid[2500:] = 1                      # two contigs: the prune rule looks at positions only
    s.pos[2500:] -= s.pos[2500] - 100
    first = uniform_blocks(s, 130)
    meta = lc.variant_meta(s)
    own, halo = tb.plan_shards(first, meta, w, n_shards)
    nb = len(first) - 1
    assert own[0] == 0 and own[-1] == nb and np.all(np.diff(own.astype(np.int64)) >= 1)   # consecutive, non-empty, complete
    prune = reference_prune(first, s.pos, w)
    for k in range(n_shards):
        assert own[k + 1] <= halo[k] <= nb
        assert halo[k] == max(own[k + 1], prune[own[k]:own[k + 1]].max())                 # exactly the reachable blocks
    if w == 15_000 and n_shards > 1:      # equal pair work, not equal block counts
        n = np.diff(first.astype(np.int64))
        work = np.array([n[b] * (n[b] - 1) / 2 + n[b] * (first[prune[b]] - first[b + 1]) for b in range(nb)])
        per = np.array([work[own[k]:own[k + 1]].sum() for k in range(n_shards)])
        assert per.max() < 1.35 * per.mean()


def test_plan_shards_more_shards_than_blocks_and_bad_input():
    s = tf.synth_genotypes(8, 300, seed=12)
    first = uniform_blocks(s, 100)
    meta = lc.variant_meta(s)
    own, halo = tb.plan_shards(first, meta, 1000, 5)          # 3 blocks, 5 shards: the extra shards own nothing
    assert list(own) == [0, 1, 2, 3, 3, 3] and list(halo[3:]) == [3, 3]
    with pytest.raises(tb.TwkbError):
        tb.plan_shards(first[:-1], meta, 1000, 2)             # last entry must be n_variants
    with pytest.raises(tb.TwkbError):
        tb.plan_shards(first, meta, 1000, 0)
    with pytest.raises(tb.TwkbError):
        tb.Engine(shard_blocks=2)                             # needs window mode (validated before any device call)


@pytest.mark.parametrize("mode", ["phased", "auto"])
def test_oracle_union_of_shards_equals_whole_window_run(mode):
    """The reference's -w semantics (oracle) on every shard's sub-matrix, keeping the pairs whose earlier member the
    shard owns, reproduce the whole-matrix -w run: the halo of twkb_plan_shards is sufficient and nothing is doubled."""
    bs, w = 50, 16_000     # the block span must stay below the window or the reference abandons the block pair (Q7)
    s = tf.synth_genotypes(300, 2300, seed=13, missing_rate=0.03 if mode == "auto" else 0.0)
    prm = dict(window=1, l_window=w, block_size=bs, minR2=0.02, force_phased=1 if mode == "phased" else 0)
    whole, visited = lc.calc(s, lc.default_params(**prm))
    first = uniform_blocks(s, bs)
    own, halo = tb.plan_shards(first, lc.variant_meta(s), w, 4)
    parts = []
    for k in range(4):
        v0, v1, ve = first[own[k]], first[halo[k]], first[own[k + 1]]
        recs, _ = lc.calc(sub_synth(s, v0, v1), lc.default_params(**prm))
        pos_a = recs["packA"] >> 2
        parts.append(recs[pos_a < s.pos[ve - 1] + 1])         # one contig, increasing positions: own <=> posA <= last own position
    got = np.concatenate(parts)
    assert len(whole) > 500
    assert np.array_equal(tf.canonical(got, False).view(np.uint8), tf.canonical(whole, False).view(np.uint8))


# ------------------------------------------------------------------ GPU
SHARD_CASES = [   # name, synth, settings, extra settings, shards, .twk block length (the block span must stay below the window, Q7)
    ("phased_sparse", dict(n_samples=900, n_variants=3100, seed=94, rare_fraction=0.8), dict(force_phased=1, minR2=0.05, window=1, l_window=40000),
     dict(sparse_max_words=6), 3, 300),
    ("phased_dense_tight", dict(n_samples=300, n_variants=2600, seed=38), dict(force_phased=1, minR2=0.0, window=1, l_window=15000), {}, 4, 100),
    ("unphased_missing", dict(n_samples=300, n_variants=2100, seed=77, missing_rate=0.05), dict(forced_unphased=1, minR2=0.1, window=1, l_window=30000), {}, 2, 250),
    ("auto_missing", dict(n_samples=200, n_variants=1900, seed=78, missing_rate=0.02), dict(minR2=0.1, window=1, l_window=20000), {}, 3, 150),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,skw,prm,extra,n_shards,bs", SHARD_CASES, ids=[c[0] for c in SHARD_CASES])
def test_engine_union_of_shards_equals_whole(name, skw, prm, extra, n_shards, bs):
    s = tf.synth_genotypes(**skw)
    data, mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    first = uniform_blocks(s, bs)
    eng = tb.Engine(**prm, **extra)
    eng.load(s.n_samples, data, mask, meta)
    eng.set_blocks(first[:-1])
    whole = eng.compute()
    st = eng.stats()
    eng.close()
    own, halo = tb.plan_shards(first, meta, prm["l_window"], n_shards)
    parts, visited, loaded = [], 0, 0
    for k in range(n_shards):
        v0, v1 = int(first[own[k]]), int(first[halo[k]])
        e = tb.Engine(shard_blocks=int(own[k + 1] - own[k]), **prm, **extra)
        e.load(s.n_samples, data[v0:v1], mask[v0:v1] if mask is not None else None, meta[v0:v1])
        e.set_blocks(first[own[k]:halo[k]] - v0)
        parts.append(e.compute())
        visited += e.stats().pairs_visited
        loaded += v1 - v0
        e.close()
    got = np.concatenate(parts)
    assert len(whole) > 100
    assert visited == st.pairs_visited
    assert np.array_equal(tf.canonical(got, False).view(np.uint8), tf.canonical(whole, False).view(np.uint8))
    if name == "phased_dense_tight":
        assert loaded < 1.5 * s.n_variants        # own + halo, not n_shards copies of the matrix


@pytest.mark.gpu
def test_cli_window_on_several_devices_runs_position_shards(tmpdir_repo):
    """twkb_calc -w -g a,b,c: the C++ mirror loads one shard per device context (runs decoded on the device, and the
    host-unpack arrangement); file content equals the one-device run."""
    s = tf.synth_genotypes(400, 3300, seed=79, rare_fraction=0.5)
    twk = os.path.join(tmpdir_repo, "shard_cli.twk")
    tf.write_twk(twk, s, block_size=300)
    exe = os.path.join(ROOT, "tomahawk_b200", "twkb_calc")
    outs = {}
    for name, extra in (("one", ["-g", "0"]), ("shards", ["-g", "0,0,0"]), ("shards_host", ["-g", "0,0", "--host-unpack"])):
        out = os.path.join(tmpdir_repo, f"shard_cli_{name}.two")
        r = subprocess.run([exe, "calc", "-p", "-r", "0.05", "-w", "50000", "-i", twk, "-o", out, "-t", "4", *extra],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert ("position shards" in r.stderr) == (name != "one")
        outs[name] = tf.canonical(tf.read_two(out), forward_only=False)
    assert len(outs["one"]) > 100
    assert np.array_equal(outs["shards"].view(np.uint8), outs["one"].view(np.uint8))
    assert np.array_equal(outs["shards_host"].view(np.uint8), outs["one"].view(np.uint8))
