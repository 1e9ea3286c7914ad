"""Shared helpers of the parity tests."""
import ast
import os

import numpy as np

from oracle import ldcore as lc
from oracle import twk_format as tf

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["phased_r01", "phased_r0", "phased_odd_n", "unphased_miss", "unphased_nomiss_r0",
                "phased_miss_aligned", "phased_miss_quirks", "window", "minp_filter", "auto_mixed", "auto_window",
                "bitmap_window"]
BLOCK_CASES = ["window_blocks300"]   # block structure comes from the FILE: run through twkb_calc_file (block_size is an oracle parameter)
INTERVAL_CASES = ["interval_one", "interval_two"]   # need the file reader: run through twkb_calc_file_intervals

# Tolerances of BASELINE.json's north_star
TOL_STAT = 1e-6   # relative, D / D' / R / R2
TOL_P = 1e-4      # relative, Fisher P and chi-squared


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    s = tf.Synth(alleles=z["alleles"], pos=z["pos"], rid=z["rid"], n_samples=int(z["n_samples"]))
    recs = np.frombuffer(z["records"].tobytes(), dtype=tf.TWO_DTYPE)
    prm = ast.literal_eval(str(z["params"]))
    return s, recs, prm, int(z["pairs"]), str(z["cli"])


def keyset(recs):
    return set(zip(recs["ridA"].tolist(), (recs["packA"] >> 2).tolist(), recs["ridB"].tolist(), (recs["packB"] >> 2).tolist()))


def assert_records_bitexact(got, ref, p_rtol=0.0):
    """Same pass set; every field identical (P optionally to a relative tolerance)."""
    got = tf.canonical(got, forward_only=False)
    ref = tf.canonical(ref, forward_only=False)
    assert len(got) == len(ref), f"record count {len(got)} != {len(ref)}: only_ref={sorted(keyset(ref) - keyset(got))[:5]} only_got={sorted(keyset(got) - keyset(ref))[:5]}"
    for f in tf.TWO_DTYPE.names:
        if f == "P" and p_rtol > 0:
            # P sums pmf terms that underflow; below ~1e-290 (denormal territory) a last-ulp
            # difference in exp() is an O(1) relative change, so only require "both tiny" there
            big = ref[f] > 1e-290
            np.testing.assert_allclose(got[f][big], ref[f][big], rtol=p_rtol, atol=0, err_msg=f)
            assert np.all(got[f][~big] <= 1e-289), "P underflow region"
        else:
            assert np.array_equal(got[f], ref[f]), f"field {f} differs"


def rel_close(a, b, rtol, atol=0.0):
    return np.abs(a - b) <= atol + rtol * np.abs(b)


def unphased_pair_is_boundary(s, i, j, prm, eps=1e-7):
    """True when pair (i, j) of an unphased run sits on a decision boundary of the
    reference's math: the estimated-count `< 5` rule, the R2 cut
    or the haplotype-frequency bounds are met to within eps, so a last-ulp difference
    in acos/cos/pow decides pass/fail."""
    a = s.alleles[i].reshape(-1, 2)
    b = s.alleles[j].reshape(-1, 2)
    v = (a != 2).all(1) & (b != 2).all(1)
    ga, gb = a.sum(1), b.sum(1)
    t = np.array([[((ga == x) & (gb == y) & v).sum() for y in range(3)] for x in range(3)], dtype=np.uint64)
    loose = dict(prm)
    loose.update(minR2=0.0, maxR2=100.0, minDprime=-100.0, maxDprime=100.0, minP=1.0)
    ok, st = lc.unphased_stats(t, lc.default_params(**{k: v for k, v in loose.items() if k.startswith(("min", "max"))}))
    if not ok:
        return True  # rejected on the root bounds / `<5` rule even with loose thresholds: boundary by construction
    c = st["cnt"]
    low3 = (c[2] + c[1] + c[0]) if c[0] < c[3] else (c[3] + c[2] + c[1])
    if abs(low3 - 5.0) < 1e-6:
        return True
    if abs(st["R2"] - prm.get("minR2", 0.1)) <= eps * max(1.0, abs(st["R2"])):
        return True
    return False
