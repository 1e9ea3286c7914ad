"""Generates tests/golden/scalc_*.npz from the REFERENCE's own `scalc` (oracle/_ref/tomahawk_scalc, built from
/root/reference by oracle/build_ref.sh). Run in the build container only; the fixtures are committed.

    python tests/golden/make_golden_scalc.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ldcore as lc  # noqa: E402
from oracle import twk_format as tf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TMP = os.path.join(ROOT, "tests", "_tmp")
os.makedirs(TMP, exist_ok=True)

CASES = {
    # name: (synth kwargs, -I string, (start, stop) it parses to, -w)
    # one target, complete data, 400 neighbours
    "scalc_one": (dict(n_samples=300, n_variants=2000, seed=5), "1:100001", (100001, 100002), 20000),
    # missing genotypes: the comparator is picked per pair (auto mode), -r is overridden to 0 by the reference CLI
    "scalc_missing": (dict(n_samples=300, n_variants=2000, seed=6, missing_rate=0.02), "1:100001", (100001, 100002), 30000),
    # three targets (target x target pairs too) and 245 neighbours of which the reference drops the last 45 (blocks of 100)
    "scalc_multi_partial": (dict(n_samples=64, n_variants=900, seed=8, missing_rate=0.1), "1:40001-40201", (40001, 40201), 12345),
}


def main():
    assert os.path.exists(lc.REF_SCALC), "build the reference first: bash oracle/build_ref.sh"
    for name, (skw, ival, (a, b), L) in CASES.items():
        s = tf.synth_genotypes(**skw)
        twk = os.path.join(TMP, f"g_{name}.twk")
        tf.write_twk(twk, s)
        lc.run_reference_scalc(twk, os.path.join(TMP, f"g_{name}"), ["-I", ival, "-w", str(L)])
        recs = tf.read_two(os.path.join(TMP, f"g_{name}.two"))
        sub, nt = lc.scalc_select(s, 0, a, b, L)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), alleles=s.alleles, pos=s.pos, rid=s.rid, n_samples=np.int64(s.n_samples),
                            records=recs.view(np.uint8), interval=np.array(ival), start=np.int64(a), stop=np.int64(b),
                            l_surrounding=np.int64(L), n_targets=np.int64(nt), n_neighbours=np.int64(sub.n_variants - nt))
        print(name, "records", len(recs), "targets", nt, "neighbours", sub.n_variants - nt)


if __name__ == "__main__":
    main()
