"""`scalc` / twk_ld::ComputeSingle: one target site against its neighbourhood. Golden vectors cut from the
reference's own scalc binary pin the oracle restatement (CPU) and the engine (GPU): selection of the variants, pair set,
comparator choice, records."""
import os
import subprocess

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["scalc_one", "scalc_missing", "scalc_multi_partial"]


def load(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    s = tf.Synth(alleles=z["alleles"], pos=z["pos"], rid=z["rid"], n_samples=int(z["n_samples"]))
    recs = np.frombuffer(z["records"].tobytes(), dtype=tf.TWO_DTYPE)
    return s, recs, str(z["interval"]), int(z["start"]), int(z["stop"]), int(z["l_surrounding"]), int(z["n_targets"]), int(z["n_neighbours"])


def with_reverse(fwd):
    rev = fwd.copy()
    rev["ridA"], rev["ridB"] = fwd["ridB"], fwd["ridA"]
    rev["packA"], rev["packB"] = fwd["packB"], fwd["packA"]
    return np.concatenate([fwd, rev])


def multiset(recs):
    from numpy.lib import recfunctions as rfn

    recs = rfn.repack_fields(np.ascontiguousarray(recs))   # a multi-field index is a view WITH the dropped fields' bytes
    return sorted(bytes(x) for x in recs.view(np.uint8).reshape(len(recs), -1))


@pytest.mark.parametrize("name", CASES)
def test_oracle_scalc_matches_reference_golden(name):
    s, ref, ival, a, b, L, nt, nn = load(name)
    sub, n_targets = lc.scalc_select(s, 0, a, b, L)
    assert (n_targets, sub.n_variants - n_targets) == (nt, nn)
    got, visited = lc.calc_single(sub, lc.default_params(minR2=0.0), n_targets)
    assert visited == nt * (nt - 1) // 2 + nt * nn
    assert multiset(with_reverse(got)) == multiset(ref)      # every field of every record, forward and reverse copies


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("runs", [False, True])
def test_reader_scalc_selection(name, runs, tmpdir_repo):
    """twkb_twk_open_single == twk_ld_impl::LoadTargetSingle: [targets | neighbours] in file order, the reference's
    dropped partial block of 100 included (emulate_quirks), every neighbour without it."""
    s, ref, ival, a, b, L, nt, nn = load(name)
    path = os.path.join(tmpdir_repo, f"{name}.twk")
    tf.write_twk(path, s)
    sub, _ = lc.scalc_select(s, 0, a, b, L)
    f = tb.TwkFile(path, intervals=[ival], single_surrounding=L, runs=runs)
    assert (f.n_targets, f.n_variants) == (nt, nt + nn)
    meta = f.runs()[2] if runs else f.matrix()[2]
    assert np.array_equal(meta["pos"], sub.pos)
    if not runs:
        data, mask, _ = f.matrix()
        want, wmask = tf.pack_bits(sub)
        assert np.array_equal(data, want)
        if wmask is not None:
            assert np.array_equal(mask, wmask)
    f.close()
    full, _ = lc.scalc_select(s, 0, a, b, L, emulate_quirks=False)
    f2 = tb.TwkFile(path, intervals=[ival], single_surrounding=L, emulate_quirks=False)
    assert f2.n_variants == full.n_variants >= nt + nn
    f2.close()


def test_reader_scalc_errors(tmpdir_repo):
    s = tf.synth_genotypes(50, 700, seed=3)
    path = os.path.join(tmpdir_repo, "sc_err.twk")
    tf.write_twk(path, s)
    with pytest.raises(tb.TwkbError, match="no surrounding variants"):      # < 100 neighbours: the reference's own message
        tb.TwkFile(path, intervals=["1:30001"], single_surrounding=3000)
    with pytest.raises(tb.TwkbError, match="no data found for reference"):  # no variant at the target
        tb.TwkFile(path, intervals=["1:30051"], single_surrounding=30000)
    with pytest.raises(tb.TwkbError, match="Contig does not exist"):
        tb.TwkFile(path, intervals=["7:30001"], single_surrounding=30000)
    with pytest.raises(tb.TwkbError):
        tb.Engine(single=1, n_chunks=3)                                   # "Cannot use chunking in single mode!"


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_scalc_matches_reference_golden(name):
    s, ref, ival, a, b, L, nt, nn = load(name)
    sub, n_targets = lc.scalc_select(s, 0, a, b, L)
    data, mask = tf.pack_bits(sub)
    eng = tb.Engine(single=1, single_targets=n_targets, minR2=0.0)
    eng.load(sub.n_samples, data, mask, lc.variant_meta(sub))
    got = eng.compute()
    st = eng.stats()
    eng.close()
    assert st.pairs_visited == nt * (nt - 1) // 2 + nt * nn
    want, _ = lc.calc_single(sub, lc.default_params(minR2=0.0), n_targets)
    from tests.test_gpu_parity import check_unphased
    from tests.helpers import assert_records_bitexact
    if name == "scalc_one":      # complete data: phased math, bit-exact (P to 1e-9)
        assert_records_bitexact(got, want, p_rtol=1e-9)
        keep = ["controller", "ridA", "ridB", "packA", "packB", "cnt", "D", "Dprime", "R", "R2"]
        assert multiset(with_reverse(tf.canonical(got, False))[keep]) == multiset(tf.canonical(ref, False)[keep])
    else:                        # auto mode picks the unphased comparator for pairs with missing alleles: cubic tolerances
        assert len(got) == len(want)
        g, w = got[np.lexsort((got["packB"], got["packA"]))], want[np.lexsort((want["packB"], want["packA"]))]
        assert np.array_equal(g["packA"], w["packA"]) and np.array_equal(g["packB"], w["packB"])
        np.testing.assert_allclose(g["R2"], w["R2"], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(g["D"], w["D"], rtol=1e-6, atol=1e-12)
        phased = (w["controller"] & 1) == 1
        assert np.array_equal(g["cnt"][phased], w["cnt"][phased])


@pytest.mark.gpu
def test_cli_scalc_end_to_end(tmpdir_repo):
    """twkb_calc scalc == the reference's scalc: same records (forward + reverse) in the .two file; -r is overridden."""
    s, ref, ival, a, b, L, nt, nn = load("scalc_one")
    twk = os.path.join(tmpdir_repo, "sc_cli.twk")
    tf.write_twk(twk, s)
    out = os.path.join(tmpdir_repo, "sc_cli_out.two")
    exe = os.path.join(ROOT, "tomahawk_b200", "twkb_calc")
    r = subprocess.run([exe, "scalc", "-i", twk, "-o", out, "-I", ival, "-w", str(L), "-r", "0.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = tf.read_two(out)
    assert len(got) == len(ref)
    keep = ["controller", "ridA", "ridB", "packA", "packB", "cnt", "D", "Dprime", "R", "R2", "ChiSqFisher"]
    assert multiset(got[keep]) == multiset(ref[keep])
    gs, rs = got[np.lexsort((got["packB"], got["packA"]))], ref[np.lexsort((ref["packB"], ref["packA"]))]
    np.testing.assert_allclose(gs["P"], rs["P"], rtol=1e-9)
    # the host-unpack arrangement gives the same file content
    out2 = os.path.join(tmpdir_repo, "sc_cli_out2.two")
    s2 = tb.default_settings(single=1, l_surrounding=L, minR2=0.0, host_unpack=1)
    assert tb.twk_ld().Compute(s2, twk, out2, [ival])
    assert multiset(tf.read_two(out2)[keep]) == multiset(ref[keep])
