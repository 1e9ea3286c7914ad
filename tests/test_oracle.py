"""CPU tests: the oracle (oracle/ldcore.c) against the reference's golden vectors,
its known-answer rows and -- when the compiled reference is available -- the
reference itself. These pin the oracle."""
import numpy as np
import pytest

from oracle import ldcore as lc
from oracle import twk_format as tf
from tests.helpers import BLOCK_CASES, GOLDEN_CASES, assert_records_bitexact, load_golden

# docs/tutorial.md:608-612 of the reference: counts -> D, D', R, R2, P, T*R2
TUTORIAL_ROWS = [
    ((5002, 5, 0, 1), 0.00019944127, 1, 0.4080444, 0.16650023, 0.0011980831, 833.83313),
    ((5005, 2, 0, 1), 0.00019956089, 1, 0.57723492, 0.33320019, 0.00059904153, 1668.6665),
    ((5004, 0, 3, 1), 0.00019952102, 1, 0.49985018, 0.2498502, 0.00079872204, 1251.2498),
    ((5006, 0, 1, 1), 0.00019960076, 1, 0.70703614, 0.49990013, 0.00039936102, 2503.4999),
    ((5004, 3, 0, 1), 0.00019952102, 1, 0.49985018, 0.2498502, 0.00079872204, 1251.2498),
]


@pytest.mark.parametrize("row", TUTORIAL_ROWS)
def test_tutorial_known_answers(row):
    (c0, c1, c4, c5), D, Dp, R, R2, P, chi = row
    # four of the five rows have < 5 minor haplotypes: the doc predates the "< 5" rule
    # (ld_engine.cpp:1174-1186), which is bypassed here and asserted separately below
    ok, st = lc.phased_stats(c0, c1, c4, c5, lc.default_params(minR2=0.0, skip_min_cell_rule=1))
    assert ok
    minor = c5 + c4 + c1 if c0 >= c5 else c4 + c1 + c0
    assert lc.phased_stats(c0, c1, c4, c5, lc.default_params(minR2=0.0))[0] == (minor >= 5)
    # the doc prints 8 significant digits of float-era records
    for got, want in ((st["D"], D), (st["Dprime"], Dp), (st["R"], R), (st["R2"], R2), (st["P"], P), (st["chi_fisher"], chi)):
        assert got == pytest.approx(want, rel=2e-7)
    assert list(st["cnt"]) == [c0, c1, c4, c5]
    assert st["flags"] & 1


def test_fisher_golden():
    z = np.load("tests/golden/fisher.npz")
    for t, want in zip(z["tables"], z["two_sided"]):
        assert lc.fisher(*t) == want  # bit-exact: same algorithm, same libm lgamma


def test_fisher_edge_cases():
    assert lc.fisher(0, 0, 0, 0) == 1.0
    assert lc.fisher(10, 0, 0, 0) == 1.0          # min == max: no test
    assert lc.fisher(3, 0, 0, 3) == pytest.approx(0.1)
    assert lc.fisher(1, 1, 1, 1) == 1.0
    big = lc.fisher(2000, 10, 12, 2986)           # underflows towards 0, never negative/NaN
    assert 0.0 <= big < 1e-300


@pytest.mark.parametrize("name", GOLDEN_CASES + BLOCK_CASES)
def test_oracle_matches_reference_golden(name):
    s, ref, prm, pairs, _ = load_golden(name)
    got, visited = lc.calc(s, lc.default_params(**prm))
    assert visited == pairs
    assert_records_bitexact(got, ref)


def test_phased_filters_in_reference_order():
    prm = lc.default_params(minR2=0.0)
    assert not lc.phased_stats(2, 1, 1, 0, prm)[0]            # T < 5
    assert not lc.phased_stats(1000, 2, 2, 0, prm)[0]         # minor cells < 5
    assert not lc.phased_stats(50, 50, 50, 50, prm)[0]        # D == 0
    ok, st = lc.phased_stats(900, 50, 40, 10, lc.default_params(minR2=0.0, maxR2=0.0001))
    assert not ok                                             # maxR2
    ok, st = lc.phased_stats(900, 50, 40, 10, prm)
    assert ok and st["chi_model"] == 0.0 and st["chi_fisher"] == 1000 * st["R2"]


def test_unphased_without_hets_uses_phased_math():
    t = [[80, 5, 1], [6, 0, 2], [1, 3, 2]]
    ok, st = lc.unphased_stats(t, lc.default_params(minR2=0.0))
    assert ok and (st["flags"] & 1)
    c0 = 2 * 80 + 5 + 6
    assert st["cnt"][0] == c0


def test_unphased_cubic_sets_no_phased_flag():
    t = [[60, 10, 2], [9, 12, 3], [1, 2, 1]]
    ok, st = lc.unphased_stats(t, lc.default_params(minR2=0.0))
    assert ok and not (st["flags"] & 1)
    assert st["chi_model"] == 0.0
    assert abs(st["cnt"].sum() - 2 * 100) < 1e-6


@pytest.mark.skipif(not lc.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize(
    "skw,cli,prm",
    [
        (dict(n_samples=300, n_variants=700, seed=41), ["-p", "-r", "0.2"], dict(force_phased=1, minR2=0.2)),
        (dict(n_samples=250, n_variants=260, seed=42, missing_rate=0.1), ["-u", "-r", "0.05"], dict(forced_unphased=1, minR2=0.05)),
        (dict(n_samples=333, n_variants=600, seed=43, missing_rate=0.03), ["-r", "0.1"], dict(minR2=0.1)),
        (dict(n_samples=200, n_variants=1300, seed=44), ["-p", "-r", "0.1", "-w", "30000"], dict(force_phased=1, minR2=0.1, window=1, l_window=30000)),
    ],
)
def test_oracle_matches_live_reference(skw, cli, prm, tmpdir_repo):
    s = tf.synth_genotypes(**skw)
    twk = f"{tmpdir_repo}/live.twk"
    tf.write_twk(twk, s)
    lc.run_reference_calc(twk, f"{tmpdir_repo}/live", cli, threads=4)
    ref = tf.canonical(tf.read_two(f"{tmpdir_repo}/live.two"), forward_only=True)
    got, _ = lc.calc(s, lc.default_params(**prm))
    assert_records_bitexact(got, ref)


@pytest.mark.skipif(not lc.have_reference(), reason="oracle/_ref not built")
def test_fisher_matches_live_reference():
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.choice([30, 500, 5008]))
        a = int(rng.integers(0, n // 2 + 1)); b = int(rng.integers(0, n - a + 1)); c = int(rng.integers(0, n - a - b + 1))
        d = n - a - b - c
        assert lc.fisher(a, b, c, d) == lc.reference_fisher(a, b, c, d)
