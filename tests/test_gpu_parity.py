"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the
oracle on seeded inputs, against the committed golden vectors of the reference, and
through size-independent properties at larger sizes.

Bars (BASELINE.json north_star): contingency counts bit-exact; phased D, D', R, R2,
chi-squared bit-exact (same IEEE operations in the same order); Fisher P to 1e-9
relative (stated tolerance 1e-4; only exp() differs from glibc); unphased statistics to
1e-6 relative (CUDA acos/cos/pow differ from glibc by a few ulp), with pass/fail
disagreements allowed only on enumerated decision boundaries."""
import os

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf
from tomahawk_b200 import synth as synth_mod
from tests.helpers import (GOLDEN_CASES, TOL_P, TOL_STAT, assert_records_bitexact, keyset, load_golden,
                           unphased_pair_is_boundary)

pytestmark = pytest.mark.gpu

KERNELS = [tb.KERNEL_POPC, tb.KERNEL_AUTO]


def gpu_run(s, prm, kernel=tb.KERNEL_POPC, **extra):
    data, mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    if prm.get("bitmaps"):   # the reference takes its bitmap slaves only for -p -m -M (ld_engine.cpp:1832)
        extra = dict(extra, low_memory=1)
    eng = tb.Engine(kernel=kernel, **prm, **extra)
    eng.load(s.n_samples, data, mask, meta)
    recs = eng.compute()
    st = eng.stats()
    return eng, recs, st


def exact_tables(s):
    """numpy ground truth of the counts: phased 2x2 (masked) and unphased 3x3 for all pairs."""
    a = s.alleles
    alt = (a == 1)
    valid = (a != 2)
    g = a.reshape(a.shape[0], -1, 2)
    sv = (g != 2).all(axis=2)
    gt = np.where(sv, (g == 1).sum(axis=2), -1)
    return alt, valid, sv, gt


def check_unphased(s, got, ref, prm):
    got = tf.canonical(got, forward_only=False)
    ref = tf.canonical(ref, forward_only=False)
    kg, kr = keyset(got), keyset(ref)
    step = int(s.pos[1] - s.pos[0]) if s.n_variants > 1 else 1
    disputed = sorted(kg ^ kr)
    for (_, pa, _, pb) in disputed:   # enumerated and explained: decision boundaries only
        assert unphased_pair_is_boundary(s, pa // step, pb // step, prm), f"unexplained pass/fail disagreement at {(pa, pb)}"
    assert len(disputed) <= max(3, 0.005 * len(ref))
    common = kg & kr
    gi = np.array([k in common for k in zip(got["ridA"].tolist(), (got["packA"] >> 2).tolist(), got["ridB"].tolist(), (got["packB"] >> 2).tolist())], dtype=bool)
    ri = np.array([k in common for k in zip(ref["ridA"].tolist(), (ref["packA"] >> 2).tolist(), ref["ridB"].tolist(), (ref["packB"] >> 2).tolist())], dtype=bool)
    g, r = got[gi], ref[ri]
    assert np.array_equal(g["packA"], r["packA"]) and np.array_equal(g["packB"], r["packB"])
    phased_math = (r["controller"] & 1) == 1
    # pairs without het/het samples go through the phased math: bit-exact
    for f in ("controller", "cnt", "D", "Dprime", "R", "R2", "ChiSqFisher", "ChiSqModel"):
        assert np.array_equal(g[f][phased_math], r[f][phased_math]), f
    u = ~phased_math
    assert np.array_equal(g["controller"][u] & ~np.uint16(32), r["controller"][u] & ~np.uint16(32))
    T2 = r["cnt"][u].sum(axis=1, keepdims=True)
    assert np.all(np.abs(g["cnt"][u] - r["cnt"][u]) <= 1e-6 * T2)      # estimated counts: 1e-6 of 2T
    for f in ("D", "Dprime", "R", "R2"):
        np.testing.assert_allclose(g[f][u], r[f][u], rtol=TOL_STAT, atol=1e-12, err_msg=f)
    np.testing.assert_allclose(g["ChiSqFisher"][u], r["ChiSqFisher"][u], rtol=TOL_P, atol=1e-9)
    # Fisher runs on round()-ed estimated counts: identical unless a count sits on x.5
    same_int = np.all(np.round(g["cnt"][u]) == np.round(r["cnt"][u]), axis=1)
    assert same_int.mean() > 0.999
    big = r["P"][u][same_int] > 1e-300
    np.testing.assert_allclose(g["P"][u][same_int][big], r["P"][u][same_int][big], rtol=TOL_P)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if n != "phased_miss_quirks"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_reference_vectors(name, kernel):
    s, ref, prm, pairs, cli = load_golden(name)
    eng, got, st = gpu_run(s, prm, kernel)
    assert st.pairs_visited == pairs
    if "unphased" in name or name.startswith("auto"):
        check_unphased(s, got, ref, prm)
    else:
        assert_records_bitexact(got, ref, p_rtol=1e-9)
    eng.close()


def test_golden_phased_missing_unaligned_counts_documented_divergence():
    """2N % 128 != 0 with missing data: the reference's scalar tail is defective (Q1,
    ld_engine.cpp:594-609). The device computes the CORRECT masked counts; it must agree
    with the oracle run with quirk emulation off, and the divergence from the golden
    reference output is confined to the count fields touched by Q1."""
    s, ref, prm, pairs, _ = load_golden("phased_miss_quirks")
    eng, got, st = gpu_run(s, prm, tb.KERNEL_POPC, emulate_quirks=0)
    want, _ = lc.calc(s, lc.default_params(**prm, emulate_quirks=0))
    assert_records_bitexact(got, want, p_rtol=1e-9)
    assert len(keyset(got) ^ keyset(ref)) > 0  # the reference really is different here
    eng.close()


# ------------------------------------------------------------- seeded vs the oracle
CASES = [
    ("phased", dict(n_samples=2504, n_variants=900, seed=31), dict(force_phased=1, minR2=0.1)),
    ("phased_all", dict(n_samples=2504, n_variants=260, seed=32), dict(force_phased=1, minR2=0.0)),
    ("phased_tiny_n", dict(n_samples=3, n_variants=50, seed=33), dict(force_phased=1, minR2=0.0)),
    ("phased_ragged", dict(n_samples=1001, n_variants=129, seed=34), dict(force_phased=1, minR2=0.01)),
    ("phased_rare", dict(n_samples=2000, n_variants=700, seed=35, rare_fraction=0.8), dict(force_phased=1, minR2=0.3)),
    ("phased_missing", dict(n_samples=1024, n_variants=500, seed=36, missing_rate=0.08), dict(force_phased=1, minR2=0.05)),
    ("filters", dict(n_samples=600, n_variants=400, seed=37), dict(force_phased=1, minR2=0.05, maxR2=0.9, minDprime=0.2, maxDprime=0.95, minP=1e-3)),
    ("window", dict(n_samples=300, n_variants=2600, seed=38), dict(force_phased=1, minR2=0.1, window=1, l_window=40000)),
    ("window_tight", dict(n_samples=300, n_variants=1200, seed=39), dict(force_phased=1, minR2=0.0, window=1, l_window=700)),
]


@pytest.mark.parametrize("name,skw,prm", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("kernel", KERNELS)
def test_phased_vs_oracle(name, skw, prm, kernel):
    s = tf.synth_genotypes(**skw)
    ref, visited = lc.calc(s, lc.default_params(**prm))
    eng, got, st = gpu_run(s, prm, kernel)
    assert st.pairs_visited == visited
    assert_records_bitexact(got, ref, p_rtol=1e-9)
    eng.close()


@pytest.mark.parametrize("skw,prm", [
    (dict(n_samples=1000, n_variants=600, seed=51, missing_rate=0.05), dict(forced_unphased=1, minR2=0.1)),
    (dict(n_samples=500, n_variants=300, seed=52), dict(forced_unphased=1, minR2=0.0)),
    (dict(n_samples=37, n_variants=200, seed=53, missing_rate=0.3), dict(forced_unphased=1, minR2=0.2)),
])
@pytest.mark.parametrize("kernel", KERNELS)
def test_unphased_vs_oracle(skw, prm, kernel):
    s = tf.synth_genotypes(**skw)
    ref, visited = lc.calc(s, lc.default_params(**prm))
    eng, got, st = gpu_run(s, prm, kernel)
    assert st.pairs_visited == visited
    check_unphased(s, got, ref, prm)
    eng.close()


# ------------------------------------------------- exact counts for EVERY pair (no screen)
ALL_KERNELS = [tb.KERNEL_POPC, tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4, tb.KERNEL_AUTO]


@pytest.mark.parametrize("kernel", ALL_KERNELS)
@pytest.mark.parametrize("missing", [0.0, 0.1])
def test_phased_counts_bit_exact_all_pairs(missing, kernel):
    """Raw 2x2 counts of EVERY pair (screen off) from every count kernel -- the tcgen05 ones included,
    not only the records that survive their fp32 screen -- against the numpy ground truth."""
    if missing and kernel == tb.KERNEL_UMMA:
        pytest.skip("the int8 tensor kernel serves complete data only")
    s = tf.synth_genotypes(512 if missing else 515, 300, seed=61, missing_rate=missing)
    eng, _, st0 = gpu_run(s, dict(force_phased=1, minR2=0.5), kernel)
    c = eng.debug_candidates(True)
    if kernel != tb.KERNEL_POPC:
        assert eng.stats().kernel_used in (tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)
    alt, valid, _, _ = exact_tables(s)
    ac = s.ac
    keep = {(i, j) for i in range(s.n_variants) for j in range(i + 1, s.n_variants) if ac[i] + ac[j] > 2}
    assert {(int(x["i"]), int(x["j"])) for x in c} == keep
    i, j = c["i"].astype(int), c["j"].astype(int)
    v = valid[i] & valid[j]
    A, B = alt[i] & v, alt[j] & v
    want = np.stack([(~alt[i] & ~alt[j] & v).sum(1), (A & ~B).sum(1), (~A & B & v).sum(1), (A & B).sum(1)], axis=1)
    assert np.array_equal(c["c"][:, :4].astype(np.int64), want)
    assert np.all(c["mode"] == 0)
    eng.close()


@pytest.mark.parametrize("kernel", [tb.KERNEL_POPC, tb.KERNEL_AUTO])
@pytest.mark.parametrize("missing", [0.0, 0.15])
def test_unphased_tables_bit_exact_all_pairs(missing, kernel):
    s = tf.synth_genotypes(333, 220, seed=62, missing_rate=missing)
    eng, _, _ = gpu_run(s, dict(forced_unphased=1, minR2=0.5), kernel)
    c = eng.debug_candidates(True)
    if kernel != tb.KERNEL_POPC:
        assert eng.stats().kernel_used == tb.KERNEL_UMMA_FP4
    _, _, sv, gt = exact_tables(s)
    i, j = c["i"].astype(int), c["j"].astype(int)
    want = np.zeros((len(c), 9), dtype=np.int64)
    both = sv[i] & sv[j]
    for x in range(3):
        for y in range(3):
            want[:, 3 * x + y] = ((gt[i] == x) & (gt[j] == y) & both).sum(1)
    assert np.array_equal(c["c"].astype(np.int64), want)
    assert np.all(c["mode"] == 1)
    eng.close()


# ------------------------------------------------------------------- all count kernels
def test_tensor_and_popc_kernels_agree_bit_for_bit():
    """POPC, int8 (kind::i8) and e2m1 (kind::mxf4, the AUTO choice) kernels: identical bytes."""
    s = tf.synth_genotypes(2504, 1100, seed=71)
    prm = dict(force_phased=1, minR2=0.02)
    out = {}
    for k in (tb.KERNEL_POPC, tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4, tb.KERNEL_AUTO):
        e, r, st = gpu_run(s, prm, k)
        assert st.kernel_used == (tb.KERNEL_UMMA_FP4 if k == tb.KERNEL_AUTO else k)
        out[k] = tf.canonical(r, False)
        e.close()
    for k in out:
        assert np.array_equal(out[tb.KERNEL_POPC].view(np.uint8), out[k].view(np.uint8)), k


PLANES_CASES = [
    # masked phased 2x2 (2 operand rows per variant), unphased 3x3 without / with missing (2 / 3 rows)
    ("phased_miss", dict(n_samples=1024, n_variants=1300, seed=72, missing_rate=0.07), dict(force_phased=1, minR2=0.02)),
    ("phased_miss_all", dict(n_samples=640, n_variants=333, seed=73, missing_rate=0.2), dict(force_phased=1, minR2=0.0)),
    ("unphased_nomiss", dict(n_samples=1000, n_variants=1100, seed=74), dict(forced_unphased=1, minR2=0.05)),
    ("unphased_miss", dict(n_samples=1250, n_variants=1000, seed=75, missing_rate=0.05), dict(forced_unphased=1, minR2=0.1)),
    ("unphased_miss_all", dict(n_samples=77, n_variants=250, seed=76, missing_rate=0.3), dict(forced_unphased=1, minR2=0.0)),
    ("unphased_window", dict(n_samples=300, n_variants=2100, seed=77, missing_rate=0.05), dict(forced_unphased=1, minR2=0.1, window=1, l_window=30000)),
    ("auto_mixed", dict(n_samples=60, n_variants=900, seed=78, missing_rate=0.01), dict(minR2=0.05)),  # ~half the variants complete
]


@pytest.mark.parametrize("name,skw,prm", PLANES_CASES, ids=[c[0] for c in PLANES_CASES])
def test_planes_tensor_kernels_agree_with_popc_bit_for_bit(name, skw, prm):
    """Masked-phased and unphased tables on the tensor pipe (NP operand rows per variant, e2m1)
    must give the very bytes of the LOP3+POPC kernel: same counts -> same candidates -> same records."""
    s = tf.synth_genotypes(**skw)
    out, used = {}, {}
    for k in (tb.KERNEL_POPC, tb.KERNEL_AUTO):
        e, r, st = gpu_run(s, prm, k)
        out[k], used[k] = tf.canonical(r, False), st.kernel_used
        e.close()
    assert used[tb.KERNEL_POPC] == tb.KERNEL_POPC
    assert used[tb.KERNEL_AUTO] == tb.KERNEL_UMMA_FP4
    assert len(out[tb.KERNEL_POPC]) > 0
    assert np.array_equal(out[tb.KERNEL_POPC].view(np.uint8), out[tb.KERNEL_AUTO].view(np.uint8))


def test_planes_tensor_parts_union_equals_whole():
    s = tf.synth_genotypes(400, 700, seed=79, missing_rate=0.05)
    prm = dict(forced_unphased=1, minR2=0.05)
    e, whole, _ = gpu_run(s, prm, tb.KERNEL_AUTO)
    e.close()
    parts = []
    for r in range(3):
        e, got, st = gpu_run(s, prm, tb.KERNEL_AUTO, part_index=r, part_count=3)
        assert st.kernel_used == tb.KERNEL_UMMA_FP4
        parts.append(got)
        e.close()
    union = tf.canonical(np.concatenate(parts), False)
    assert np.array_equal(union.view(np.uint8), tf.canonical(whole, False).view(np.uint8))


def _dense_matrix(n_samples, n_variants, seed, chain=False):
    """Adversarial inputs for the exactness of the e2m1 kernel's fp32 accumulation: nearly every
    haplotype carries the alt allele (counts close to 2N), or -- chain=True -- every variant is its
    predecessor with 2 % of the haplotypes flipped (half-dense rows in strong LD, so records
    survive an R2 cut and counts are ~N)."""
    rng = np.random.default_rng(seed)
    nb = 2 * n_samples
    words = (nb + 127) // 128 * 2
    data = np.zeros((n_variants, words), np.uint64)
    ac = np.zeros(n_variants, np.uint32)
    prev = None
    for v in range(n_variants):
        bits = np.zeros(words * 64, np.uint8)
        if chain and prev is not None:
            bits[:nb] = prev[:nb] ^ (rng.random(nb) < 0.02)
        else:
            bits[:nb] = rng.random(nb) < (0.5 if chain or v % 3 == 0 else 0.97)
        bits[0] = 0
        prev = bits
        ac[v] = bits.sum()
        data[v] = np.packbits(bits, bitorder="little").view(np.uint64)
    meta = np.zeros(n_variants, tb.VARIANT_DTYPE)
    meta["pos"] = 100 * (1 + np.arange(n_variants)); meta["ac"] = ac; meta["hwe"] = 1.0; meta["gt_phase"] = 1
    return data, meta


@pytest.mark.parametrize("n_samples,n_variants,minR2,kernels", [
    (2504, 700, 0.0, (tb.KERNEL_POPC, tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)),
    (60000, 600, 0.0, (tb.KERNEL_POPC, tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)),      # counts up to ~116,000
    (500000, 520, 0.3, (tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)),                     # 1M haplotypes, counts ~500,000
])
def test_e2m1_accumulation_exact_on_dense_data(n_samples, n_variants, minR2, kernels):
    data, meta = _dense_matrix(n_samples, n_variants, seed=n_samples, chain=n_samples >= 500000)
    out = []
    for k in kernels:
        eng = tb.Engine(force_phased=1, minR2=minR2, kernel=k)
        eng.load(n_samples, data, None, meta)
        out.append(tf.canonical(eng.compute(), False))
        assert eng.stats().kernel_used == k
        eng.close()
    assert len(out[0]) > 0
    for o in out[1:]:
        assert np.array_equal(out[0].view(np.uint8), o.view(np.uint8))
    # ground truth for a few pairs straight from the bits
    bits = np.unpackbits(data[:8].view(np.uint8), axis=1, bitorder="little")[:, :2 * n_samples].astype(np.int64)
    n11 = bits @ bits.T
    r = out[-1]
    step = 100
    ia, ib = (r["packA"] >> 2) // step - 1, (r["packB"] >> 2) // step - 1
    sel = (ia < 8) & (ib < 8)
    assert sel.any()
    assert np.array_equal(r["cnt"][sel][:, 3], n11[ia[sel], ib[sel]].astype(np.float64))


# ------------------------------------------------------------ rare-variant (list) kernel
def _cands_sorted(eng):
    c = eng.debug_candidates(True)
    return c[np.lexsort((c["j"], c["i"]))]


@pytest.mark.parametrize("T", [3, 12, 10_000])
def test_sparse_kernel_counts_bit_exact_all_pairs(T):
    """Variants with <= T non-zero words go through the list kernel (count_sparse.cuh), the rest
    through the dense kernels on the [dense | sparse]-ordered resident matrix: every pair exactly
    once, oriented by file order, with the very counts of the all-dense run."""
    s = tf.synth_genotypes(1500, 900, seed=91, rare_fraction=0.7)
    e0, _, _ = gpu_run(s, dict(force_phased=1, minR2=0.5), sparse_max_words=-1)
    e1, _, st = gpu_run(s, dict(force_phased=1, minR2=0.5), sparse_max_words=T)
    assert st.sparse_variants > 100 and st.sparse_launches > 0
    if T >= 10_000:
        assert st.sparse_variants == s.n_variants   # everything sparse: no dense phase at all
    else:
        assert st.sparse_variants < s.n_variants
    a, b = _cands_sorted(e0), _cands_sorted(e1)
    assert len(a) == len(b)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    e0.close(); e1.close()


SPARSE_CASES = [
    ("r2_002", dict(n_samples=2504, n_variants=2300, seed=92, rare_fraction=0.8), dict(force_phased=1, minR2=0.02), 20),
    ("r2_0_all_pairs", dict(n_samples=700, n_variants=500, seed=93, rare_fraction=0.6), dict(force_phased=1, minR2=0.0), 5),
    ("window", dict(n_samples=900, n_variants=3100, seed=94, rare_fraction=0.8), dict(force_phased=1, minR2=0.05, window=1, l_window=40000), 6),
    ("auto_mode_complete_data", dict(n_samples=1000, n_variants=1200, seed=95, rare_fraction=0.5), dict(minR2=0.05), 6),
]


@pytest.mark.parametrize("name,skw,prm,T", SPARSE_CASES, ids=[c[0] for c in SPARSE_CASES])
def test_sparse_path_records_identical_to_dense(name, skw, prm, T):
    s = tf.synth_genotypes(**skw)
    e, ref, st0 = gpu_run(s, prm, tb.KERNEL_POPC, sparse_max_words=-1)
    e.close()
    e, got, st = gpu_run(s, prm, tb.KERNEL_AUTO, sparse_max_words=T)
    e.close()
    assert st.sparse_variants > 0.1 * s.n_variants and st.sparse_word_ops > 0
    assert st.kernel_used == tb.KERNEL_UMMA_FP4          # dense x dense stays on the tensor pipe
    assert st.pairs_visited == st0.pairs_visited
    assert len(ref) > 0
    assert np.array_equal(tf.canonical(got, False).view(np.uint8), tf.canonical(ref, False).view(np.uint8))


def test_sparse_path_parts_union_equals_whole():
    s = tf.synth_genotypes(800, 2600, seed=96, rare_fraction=0.75)
    prm = dict(force_phased=1, minR2=0.05)
    e, whole, st = gpu_run(s, prm, tb.KERNEL_POPC, sparse_max_words=-1)
    e.close()
    chunks, visited = [], 0
    for r in range(3):
        e, recs, stp = gpu_run(s, prm, tb.KERNEL_AUTO, sparse_max_words=6, part_index=r, part_count=3)
        assert stp.sparse_variants > 0
        chunks.append(recs)
        visited += stp.pairs_visited
        e.close()
    assert visited == st.pairs_visited
    allr = np.concatenate(chunks)
    assert len(keyset(allr)) == len(allr)
    assert np.array_equal(tf.canonical(allr, False).view(np.uint8), tf.canonical(whole, False).view(np.uint8))


_biobank_matrix = synth_mod.biobank_matrix


def test_biobank_scale_window_sparse_auto_threshold():
    """1,000,000 haplotypes, -w window, rare-variant class chosen by the AUTOMATIC threshold
    (sparse_max_words = 0 -> ceil(2N/32)/64 = 488 words): records identical to the all-dense run,
    counts checked against the bits, and the list kernel must do far less work than dense rows."""
    n_samples, n_variants = 500_000, 1200
    data, meta = _biobank_matrix(n_samples, n_variants, seed=5)
    prm = dict(force_phased=1, minR2=0.2, window=1, l_window=50_000)
    out, sts = [], []
    for smw in (-1, 0):
        eng = tb.Engine(kernel=tb.KERNEL_AUTO, sparse_max_words=smw, **prm)
        eng.load(n_samples, data, None, meta)
        out.append(tf.canonical(eng.compute(), False))
        sts.append(eng.stats())
        eng.close()
    dense, auto = sts
    assert dense.sparse_variants == 0 and auto.sparse_variants > 0.3 * n_variants
    assert auto.kernel_used == tb.KERNEL_UMMA_FP4 and auto.sparse_launches > 0
    assert auto.pairs_visited == dense.pairs_visited
    assert len(out[0]) > 100
    assert np.array_equal(out[0].view(np.uint8), out[1].view(np.uint8))
    # the list kernel touches only the non-zero words of the rare rows
    K32 = (2 * n_samples + 31) // 32
    assert auto.sparse_word_ops < 0.05 * auto.sparse_variants * n_variants * K32
    r = out[1]
    assert np.all(r["cnt"].sum(axis=1) == 2 * n_samples)
    ia, ib = (r["packA"] >> 2) // 100 - 1, (r["packB"] >> 2) // 100 - 1
    assert np.all(r["cnt"][:, 1] + r["cnt"][:, 3] == meta["ac"][ia])
    assert np.all(r["cnt"][:, 2] + r["cnt"][:, 3] == meta["ac"][ib])
    for k in np.linspace(0, len(r) - 1, 40).astype(int):      # ALTALT straight from the bits
        n11 = int(np.unpackbits((data[ia[k]] & data[ib[k]]).view(np.uint8)).sum())
        assert r["cnt"][k, 3] == n11


def test_sparse_path_rejects_mode_change_without_reload():
    s = tf.synth_genotypes(300, 400, seed=97, rare_fraction=0.8)
    e, _, st = gpu_run(s, dict(force_phased=1, minR2=0.1), tb.KERNEL_AUTO, sparse_max_words=4)
    assert st.sparse_variants > 0
    e.update(force_phased=0, forced_unphased=1)
    with pytest.raises(tb.TwkbError):
        e.compute()
    e.close()


# ----------------------------------------------------------- multi-part = whole (no GPU-GPU traffic)
@pytest.mark.parametrize("parts", [2, 3])
def test_parts_union_equals_whole(parts):
    s = tf.synth_genotypes(800, 1500, seed=81)
    prm = dict(force_phased=1, minR2=0.05)
    eng, whole, st = gpu_run(s, prm)
    eng.close()
    chunks, visited = [], 0
    for r in range(parts):
        e, recs, stp = gpu_run(s, prm, part_index=r, part_count=parts)
        chunks.append(recs)
        visited += stp.pairs_visited
        e.close()
    assert visited == st.pairs_visited
    allr = np.concatenate(chunks)
    assert len(keyset(allr)) == len(allr)
    assert np.array_equal(tf.canonical(allr, False).view(np.uint8), tf.canonical(whole, False).view(np.uint8))


def test_chunked_runs_cover_the_triangle():
    s = tf.synth_genotypes(400, 2100, seed=82)
    prm = dict(force_phased=1, minR2=0.1)
    eng, whole, st = gpu_run(s, prm)
    eng.close()
    chunks = []
    for c in range(3):
        e, recs, _ = gpu_run(s, prm, n_chunks=3, c_chunk=c)
        chunks.append(recs)
        e.close()
    allr = np.concatenate(chunks)
    assert np.array_equal(tf.canonical(allr, False).view(np.uint8), tf.canonical(whole, False).view(np.uint8))


# ------------------------------------------------------------------------- file to file
def test_calc_file_end_to_end_matches_reference_golden(tmpdir_repo):
    s, ref, prm, pairs, cli = load_golden("phased_r01")
    twk = os.path.join(tmpdir_repo, "e2e.twk")
    tf.write_twk(twk, s)
    ld = tb.twk_ld()
    st = tb.default_settings(**prm)
    assert ld.Compute(st, twk, os.path.join(tmpdir_repo, "e2e_out"))
    back = tf.read_two(os.path.join(tmpdir_repo, "e2e_out.two"))
    assert len(back) == 2 * len(ref)
    assert_records_bitexact(tf.canonical(back, forward_only=True), ref, p_rtol=1e-9)
    assert ld.last_stats.pairs_visited == pairs
    assert not ld.Compute(st, os.path.join(tmpdir_repo, "missing.twk"), os.path.join(tmpdir_repo, "x"))


@pytest.mark.parametrize("name", ["interval_one", "interval_two"])
def test_calc_file_interval_mode_matches_reference_golden(name, tmpdir_repo):
    """`calc -I`: block-granular interval selection + LD, file to file, against the reference's records."""
    s, ref, prm, pairs, cli = load_golden(name)
    t = cli.split()
    ivals = [t[i + 1] for i in range(len(t)) if t[i] == "-I"]
    twk = os.path.join(tmpdir_repo, f"{name}_gpu.twk")
    tf.write_twk(twk, s)
    ld = tb.twk_ld()
    assert ld.Compute(tb.default_settings(**prm), twk, os.path.join(tmpdir_repo, f"{name}_gpu_out"), ival_strings=ivals)
    back = tf.read_two(os.path.join(tmpdir_repo, f"{name}_gpu_out.two"))
    assert len(back) == 2 * len(ref)
    assert_records_bitexact(tf.canonical(back, forward_only=True), ref, p_rtol=1e-9)
    assert ld.last_stats.pairs_visited == pairs
    assert not ld.Compute(tb.default_settings(**prm), twk, os.path.join(tmpdir_repo, "x"), ival_strings=["nochr:1-2"])


# ---------------------------------------------------------- properties at larger sizes
def test_properties_at_scale():
    """20,000 x 5,008 haplotypes (2e8 pairs): too big for the scalar oracle, so check
    invariants: table sums, marginals, symmetry under variant reversal, idempotence."""
    s = tf.synth_genotypes(2504, 20000, seed=91)
    prm = dict(force_phased=1, minR2=0.2)
    eng, recs, st = gpu_run(s, prm, tb.KERNEL_AUTO)
    assert st.pairs_visited == 20000 * 19999 // 2
    assert len(recs) > 1000
    assert np.all(recs["cnt"].sum(axis=1) == 5008)
    step = int(s.pos[1] - s.pos[0])
    ia, ib = (recs["packA"] >> 2) // step, (recs["packB"] >> 2) // step
    ac = s.ac
    assert np.all(recs["cnt"][:, 1] + recs["cnt"][:, 3] == ac[ia])      # A alt marginal
    assert np.all(recs["cnt"][:, 2] + recs["cnt"][:, 3] == ac[ib])      # B alt marginal
    assert np.all((recs["R2"] >= 0.2) & (recs["R2"] <= 1.0 + 1e-12))
    assert np.all(np.abs(recs["Dprime"]) <= 1.0 + 1e-9)
    assert np.all((recs["P"] >= 0) & (recs["P"] <= 1))
    again = eng.compute()
    assert np.array_equal(tf.canonical(again, False).view(np.uint8), tf.canonical(recs, False).view(np.uint8))
    eng.close()
    # reversed variant order: the same unordered pairs pass, R2 identical
    rev = tf.Synth(alleles=s.alleles[::-1].copy(), pos=s.pos.copy(), rid=s.rid.copy(), n_samples=s.n_samples)
    eng2, recs2, _ = gpu_run(rev, prm, tb.KERNEL_AUTO)
    n = s.n_variants
    ja, jb = (recs2["packA"] >> 2) // step, (recs2["packB"] >> 2) // step
    k1 = np.sort(ia.astype(np.int64) * n + ib)
    k2 = np.sort((n - 1 - jb).astype(np.int64) * n + (n - 1 - ja))
    assert np.array_equal(k1, k2)
    o1 = np.argsort(ia.astype(np.int64) * n + ib)
    o2 = np.argsort((n - 1 - jb).astype(np.int64) * n + (n - 1 - ja))
    np.testing.assert_allclose(recs["R2"][o1], recs2["R2"][o2], rtol=1e-12)
    eng2.close()


# ------------------------------------------- BASELINE.json configs, at size or on a stated sub-sample
def test_config1_full_size_tensor_records_equal_popc_records():
    """BASELINE configs[1] AT SIZE (2,504 samples x 200,000 SNVs, R2 >= 0.1, 2.0e10 pairs): the default e2m1
    tcgen05 kernel (fp32 5-instruction screen -> survivor ring -> exact decision) must hand over the very
    same records, byte for byte, as the LOP3+POPC kernel (exact integer counts, fp64 screen)."""
    s = tf.synth_genotypes(2504, 200_000, seed=20)
    data, mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    del s
    out = {}
    for k in (tb.KERNEL_AUTO, tb.KERNEL_POPC):
        eng = tb.Engine(kernel=k, force_phased=1, minR2=0.1)
        eng.load(2504, data, mask, meta)
        recs = eng.compute()
        st = eng.stats()
        assert st.pairs_visited == 200_000 * 199_999 // 2
        assert st.kernel_used == (tb.KERNEL_UMMA_FP4 if k == tb.KERNEL_AUTO else tb.KERNEL_POPC)
        out[k] = tf.canonical(recs, forward_only=False).view(np.uint8)   # (ridA, posA, ridB, posB) order: keys are unique
        eng.close()
    assert len(out[tb.KERNEL_AUTO]) == len(out[tb.KERNEL_POPC]) > 500_000 * 106
    assert np.array_equal(out[tb.KERNEL_AUTO], out[tb.KERNEL_POPC])


def test_small_work_buffers_overflow_retry_and_record_rotation(monkeypatch):
    """Candidate buffer of little more than one tile and a record buffer of the same size (TWKB_CAND_CAP / TWKB_REC_CAP are
    read when a context allocates them): (a) R2 >= 0, every pair a record -- one tile per batch, the record buffers rotate
    through the drain thread a dozen times; (b) a matrix with little LD, then one of the SAME shape in long LD blocks --
    the survivor rate kept from the first run sizes a first batch that overflows the candidate buffer, which must be
    retried with fewer tiles. Records identical to runs with the default 16 M-entry buffers."""
    def records(eng, s, **kw):
        data, mask = tf.pack_bits(s)
        eng.load(s.n_samples, data, mask, lc.variant_meta(s))
        return tf.canonical(eng.compute(), forward_only=False).view(np.uint8)

    s0 = tf.synth_genotypes(600, 1500, seed=11)
    sa = tf.synth_genotypes(600, 12_000, seed=12, p_copy=0.0)
    sb = tf.synth_genotypes(600, 12_000, seed=13, p_copy=0.98, redraw=0.01)
    ref = {}
    eng = tb.Engine(force_phased=1, minR2=0.0)
    ref["s0"] = records(eng, s0)
    eng.close()
    eng = tb.Engine(force_phased=1, minR2=0.1)
    ref["sa"], ref["sb"] = records(eng, sa), records(eng, sb)
    eng.close()
    assert len(ref["s0"]) > 4 * 70_000 * 106 and len(ref["sb"]) > 70_000 * 106 > len(ref["sa"])

    monkeypatch.setenv("TWKB_CAND_CAP", "70000")
    monkeypatch.setenv("TWKB_REC_CAP", "70000")
    eng = tb.Engine(force_phased=1, minR2=0.0)
    got = records(eng, s0)
    st = eng.stats()
    assert st.count_launches >= 20                       # 27 tiles of 256 x 240, one per batch
    assert np.array_equal(got, ref["s0"])
    eng.close()
    eng = tb.Engine(force_phased=1, minR2=0.1)
    assert np.array_equal(records(eng, sa), ref["sa"])
    n_sa = eng.stats().count_launches
    assert np.array_equal(records(eng, sb), ref["sb"])   # first batch sized by sa's survivor rate: overflow, retry
    assert eng.stats().count_launches > n_sa + 2
    eng.close()


def test_400k_variant_load_equals_union_of_two_overlapping_loads():
    """Loads of >= 400,000 variants run the host-side metadata passes (the scheduler's copy of contig / position, the
    device records) chunk-wise on a thread team (engine.cu: host_meta_work); smaller loads use the two-thread arrangement.
    A -w run over 400,000 variants (window 60 kb = 600 variants, above the 50 kb span of a .twk block) must give exactly the
    union of the same run over two overlapping, block-aligned halves of the matrix."""
    s = tf.synth_genotypes(300, 400_000, seed=41, p_copy=0.9)
    data, mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    prm = dict(force_phased=1, minR2=0.3, window=1, l_window=60_000)

    def run(lo, hi):
        eng = tb.Engine(**prm)
        eng.load(300, np.ascontiguousarray(data[lo:hi]), None if mask is None else np.ascontiguousarray(mask[lo:hi]),
                 np.ascontiguousarray(meta[lo:hi]))
        recs = eng.compute()
        eng.close()
        return tf.canonical(recs, forward_only=False)

    whole = run(0, 400_000)
    parts = np.concatenate([run(0, 250_000), run(150_000, 400_000)])
    parts = np.unique(parts.view(np.dtype((np.void, 106)))).view(np.uint8).reshape(-1, 106)
    whole_u = np.unique(whole.view(np.dtype((np.void, 106)))).view(np.uint8).reshape(-1, 106)
    assert len(whole_u) == len(whole) > 100_000           # keys are unique; plenty of records
    assert np.array_equal(whole_u, parts)


@pytest.mark.timeout(300)
def test_two_contexts_on_one_device_run_concurrently_without_hanging():
    """Two contexts on ONE device, each launching full-grid (148-CTA) persistent count kernels from its own host thread
    and stream at the same time. The CTA pairs of a launch pace each other through global counters; if the two grids
    ever shared the SMs, pairs spinning for non-resident siblings would hold the SMs those need -- the wait is bounded
    (count_umma.cuh, UMMA3_PACE_MAX_SPINS) so the worst case is an un-paced launch. Both runs must finish and give the
    records of a run that had the device to itself."""
    import threading

    s = tf.synth_genotypes(2504, 40_000, seed=7)
    data, mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    del s
    engines = [tb.Engine(force_phased=1, minR2=0.1) for _ in range(2)]
    for e in engines:
        e.load(2504, data, mask, meta)
    alone = tf.canonical(engines[0].compute(), forward_only=False).view(np.uint8)
    assert engines[0].stats().kernel_used == tb.KERNEL_UMMA_FP4
    results, errors = [None, None], []
    gate = threading.Barrier(2)

    def worker(k):
        try:
            gate.wait()
            for _ in range(6):
                engines[k].compute_resident()          # ctypes releases the GIL: the two launch streams overlap
            results[k] = tf.canonical(engines[k].compute(), forward_only=False).view(np.uint8)
        except Exception as ex:                        # noqa: BLE001 -- reported by the main thread
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in threads), "a concurrent run did not finish"
    assert not errors, errors
    for k in range(2):
        assert np.array_equal(results[k], alone)
    for e in engines:
        e.close()


def _live_reference(s, cli, tmpdir, name, threads=None):
    twk = os.path.join(tmpdir, f"{name}.twk")
    tf.write_twk(twk, s)
    info = lc.run_reference_calc(twk, os.path.join(tmpdir, name), cli, threads=threads)
    return tf.read_two(os.path.join(tmpdir, f"{name}.two")), info


@pytest.mark.skipif(not lc.have_reference(), reason="oracle/_ref/tomahawk_calc not built")
@pytest.mark.parametrize("kernel", KERNELS)
def test_config0_subsample_vs_reference_binary(kernel, tmpdir_repo):
    """BASELINE configs[0] shape (2,504 samples, -p, R2 >= 0: EVERY pair is a record and goes through Fisher)
    on its first 2,000 variants, against the reference's own `tomahawk calc` run on this host."""
    s = tf.synth_genotypes(2504, 2000, seed=20)
    ref, info = _live_reference(s, ["-p", "-r", "0"], tmpdir_repo, "c0_sub")
    eng, got, st = gpu_run(s, dict(force_phased=1, minR2=0.0), kernel)
    if "pairs" in info:
        assert st.pairs_visited == info["pairs"]
    assert len(got) > 1_000_000
    assert_records_bitexact(got, tf.canonical(ref, forward_only=True), p_rtol=1e-9)
    eng.close()


@pytest.mark.skipif(not lc.have_reference(), reason="oracle/_ref/tomahawk_calc not built")
def test_config2_subsample_vs_reference_binary(tmpdir_repo):
    """BASELINE configs[2] shape (10,000 samples, -u, 5 % missing genotypes, Fisher + chi2) on its first 1,500
    variants, against the reference's own `tomahawk calc`; tolerances of north_star, flips enumerated."""
    s = tf.synth_genotypes(10_000, 1500, seed=20, missing_rate=0.05)
    ref, info = _live_reference(s, ["-u", "-r", "0.1"], tmpdir_repo, "c2_sub")
    prm = dict(forced_unphased=1, minR2=0.1)
    eng, got, st = gpu_run(s, prm, tb.KERNEL_AUTO)
    assert st.kernel_used == tb.KERNEL_UMMA_FP4 and st.n_planes == 3
    if "pairs" in info:
        assert st.pairs_visited == info["pairs"]
    check_unphased(s, got, tf.canonical(ref, forward_only=True), prm)
    eng.close()


def test_calc_file_follows_the_files_block_structure(tmpdir_repo):
    """`import -b 300`: the reference's window rules (balancer row prune over blocks, abort of a block pair at its first
    out-of-window pair) are defined on the file's own blocks, which twkb_calc_file takes from the index (twkb_set_blocks);
    assuming 500-variant blocks gives a different record set."""
    s, ref, prm, pairs, cli = load_golden("window_blocks300")
    twk = os.path.join(tmpdir_repo, "blk300.twk")
    assert tf.write_twk(twk, s, block_size=300) == 7
    assert list(tb.TwkFile(twk).blocks()) == [0, 300, 600, 900, 1200, 1500, 1800]
    st = tb.default_settings(force_phased=1, minR2=0.1, window=1, l_window=45000)
    ld = tb.twk_ld()
    out = os.path.join(tmpdir_repo, "blk300.two")
    assert ld.Compute(st, twk, out)
    assert ld.last_stats.pairs_visited == pairs
    assert_records_bitexact(tf.canonical(tf.read_two(out), forward_only=True), ref, p_rtol=1e-9)
    # the same matrix without the file's blocks: 500-variant blocks are assumed, and the result differs
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(force_phased=1, minR2=0.1, window=1, l_window=45000)
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    other = eng.compute()
    assert len(keyset(other) ^ keyset(ref)) > 0
    eng.set_blocks(np.arange(0, 1900, 300))
    assert_records_bitexact(eng.compute(), ref, p_rtol=1e-9)
    eng.close()


def test_edge_cases():
    # a single pair
    s = tf.synth_genotypes(50, 2, seed=95)
    ref, _ = lc.calc(s, lc.default_params(force_phased=1, minR2=0.0))
    eng, got, st = gpu_run(s, dict(force_phased=1, minR2=0.0))
    assert st.pairs_visited == 1
    assert_records_bitexact(got, ref, p_rtol=1e-9)
    eng.close()
    # singletons only: every pair has ac_i + ac_j <= 2 and is skipped (ld_engine.cpp:1918)
    al = np.zeros((20, 200), dtype=np.uint8)
    for v in range(20):
        al[v, v] = 1
    s = tf.Synth(alleles=al, pos=(np.arange(20) * 100).astype(np.uint32), rid=np.zeros(20, np.uint32), n_samples=100)
    eng, got, st = gpu_run(s, dict(force_phased=1, minR2=0.0))
    assert len(got) == 0 and st.pairs_visited == 190
    eng.close()
    # compute before load
    e = tb.Engine(force_phased=1)
    with pytest.raises(tb.TwkbError):
        e.compute()
    e.close()


# ---------------------------------------------------------- device-side .twk decode
@pytest.mark.parametrize("kw,widths", [
    (dict(n_samples=2504, n_variants=1500, seed=31), None),
    (dict(n_samples=333, n_variants=700, seed=32, missing_rate=0.07), [1, 2, 4]),
    (dict(n_samples=31, n_variants=41, seed=33, missing_rate=0.2), [1, 2, 4]),
    (dict(n_samples=1, n_variants=9, seed=34), [1]),
    (dict(n_samples=40000, n_variants=64, seed=35), [4]),        # interiors > 64 words: warp-cooperative stores
    (dict(n_samples=40000, n_variants=64, seed=36, missing_rate=0.01), [2, 4]),
])
def test_device_run_decode_rows_bit_identical(kw, widths):
    """decode_runs_kernel == twk_igt_vec::Build (lib/core.cpp:349-383): the resident rows and mask
    rows after twkb_load_runs equal the host-packed ones word for word."""
    s = tf.synth_genotypes(**kw)
    if kw["n_samples"] == 40000:   # long alt/alt and missing runs
        s.alleles[3, 1000:60000] = 1
        s.alleles[5, :] = 1
        if kw.get("missing_rate"):
            s.alleles[7, 20000:70000] = 2
    raw, desc = tf.encode_runs(s, widths, seed=kw["seed"])
    want_data, want_mask = tf.pack_bits(s)
    meta = lc.variant_meta(s)
    eng = tb.Engine(force_phased=1, sparse_max_words=-1)
    eng.load_runs(s.n_samples, raw, desc, meta)
    data, mask = eng.rows(s.n_variants, want_data.shape[1], with_mask=True)
    assert np.array_equal(data, want_data)
    assert np.array_equal(mask, want_mask if want_mask is not None else np.zeros_like(want_data))
    eng.close()


def test_device_run_decode_records_equal_matrix_load():
    s = tf.synth_genotypes(700, 900, seed=37, missing_rate=0.05)
    prm = dict(forced_unphased=1, minR2=0.05)
    _, want, st0 = gpu_run(s, prm, tb.KERNEL_AUTO)
    raw, desc = tf.encode_runs(s, [1, 2, 4], seed=1)
    eng = tb.Engine(kernel=tb.KERNEL_AUTO, **prm)
    eng.load_runs(s.n_samples, raw, desc, lc.variant_meta(s))
    got = eng.compute()
    assert eng.stats().pairs_visited == st0.pairs_visited
    assert np.array_equal(tf.canonical(got, forward_only=False).view(np.uint8), tf.canonical(want, forward_only=False).view(np.uint8))


def test_device_run_decode_rejects_bad_runs():
    s = tf.synth_genotypes(100, 20, seed=38)
    raw, desc = tf.encode_runs(s, [2])
    meta = lc.variant_meta(s)
    eng = tb.Engine(force_phased=1)
    bad = desc.copy()
    bad["n_runs"][4] -= 1                     # runs no longer cover all samples
    with pytest.raises(tb.TwkbError, match="variant 4"):
        eng.load_runs(s.n_samples, raw, bad, meta)
    bad = desc.copy()
    bad["offset"][6] = len(raw)               # points past the buffer
    with pytest.raises(tb.TwkbError, match="truncated"):
        eng.load_runs(s.n_samples, raw, bad, meta)
    bad = desc.copy()
    bad["width"][0] = 3
    with pytest.raises(tb.TwkbError):
        eng.load_runs(s.n_samples, raw, bad, meta)
    eng.load_runs(s.n_samples, raw, desc, meta)   # the context is still usable
    assert eng.stats() is not None


@pytest.mark.parametrize("name", ["phased_r01", "unphased_miss", "auto_mixed"])
def test_calc_file_device_decode_equals_host_unpack(name, tmpdir_repo):
    """twkb_calc_file: device-side decode (default) and host unpack (host_unpack=1) write the same records."""
    s, ref, prm, pairs, cli = load_golden(name)
    twk = os.path.join(tmpdir_repo, f"dd_{name}.twk")
    tf.write_twk(twk, s)
    outs = []
    for host_unpack in (0, 1):
        ld = tb.twk_ld()
        out = os.path.join(tmpdir_repo, f"dd_{name}_{host_unpack}")
        assert ld.Compute(tb.default_settings(host_unpack=host_unpack, n_threads=4, **prm), twk, out)
        assert ld.last_stats.pairs_visited == pairs
        outs.append(tf.canonical(tf.read_two(out + ".two"), forward_only=False))
    assert np.array_equal(outs[0].view(np.uint8), outs[1].view(np.uint8))
    if prm.get("force_phased"):
        assert len(outs[0]) == 2 * len(ref)
    else:   # unphased pairs may flip on enumerated decision boundaries
        assert abs(len(outs[0]) - 2 * len(ref)) <= max(6, 0.01 * len(ref))


# ---------------------------------------------------------- survivor ring under pressure
@pytest.mark.timeout(180)
@pytest.mark.parametrize("name,skw,prm", [
    # the in-kernel screen flags most pairs, so the shared-memory ring between the epilogue warps and
    # the drain warp is full most of the time (the 1-plane ring holds 1024 entries, the planes ring 341)
    ("phased_ring", dict(n_samples=800, n_variants=2600, seed=81), dict(force_phased=1, minR2=1e-4)),
    ("unphased_nomiss_ring", dict(n_samples=900, n_variants=1500, seed=82), dict(forced_unphased=1, minR2=1e-3)),
    ("unphased_miss_ring", dict(n_samples=1250, n_variants=1500, seed=75, missing_rate=0.05), dict(forced_unphased=1, minR2=1e-3)),
    ("phased_miss_ring", dict(n_samples=700, n_variants=1500, seed=83, missing_rate=0.05), dict(force_phased=1, minR2=1e-3)),
])
def test_survivor_ring_full_is_drained_and_exact(name, skw, prm):
    """Regression for a hang: the drain warp's control flow must stay warp-uniform and a full ring must
    only delay the epilogue warps. Results still equal the LOP3+POPC kernel byte for byte."""
    s = tf.synth_genotypes(**skw)
    out = {}
    for k in (tb.KERNEL_POPC, tb.KERNEL_AUTO):
        e, r, st = gpu_run(s, prm, k)
        out[k] = tf.canonical(r, False)
        if k == tb.KERNEL_AUTO:
            assert st.kernel_used == tb.KERNEL_UMMA_FP4
            assert st.pairs_screened > 0.05 * st.pairs_visited   # the ring really was under pressure
        e.close()
    assert np.array_equal(out[tb.KERNEL_POPC].view(np.uint8), out[tb.KERNEL_AUTO].view(np.uint8))
