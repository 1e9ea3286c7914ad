"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref/tomahawk_calc and
libref_fisher.so, built from /root/reference by oracle/build_ref.sh). Run in the build
container only; the fixtures are committed so the pin travels to the GPU box.

    python tests/golden/make_golden.py
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
    # name: (synth kwargs, reference CLI args, oracle/engine params)
    "phased_r01": (dict(n_samples=2504, n_variants=400, seed=101), ["-p", "-r", "0.1"], dict(force_phased=1, minR2=0.1)),
    "phased_r0": (dict(n_samples=2504, n_variants=110, seed=102), ["-p", "-r", "0"], dict(force_phased=1, minR2=0.0)),
    "phased_odd_n": (dict(n_samples=777, n_variants=300, seed=103), ["-p", "-r", "0.02"], dict(force_phased=1, minR2=0.02)),
    "unphased_miss": (dict(n_samples=1000, n_variants=300, seed=104, missing_rate=0.05), ["-u", "-r", "0.1"], dict(forced_unphased=1, minR2=0.1)),
    "unphased_nomiss_r0": (dict(n_samples=400, n_variants=120, seed=105), ["-u", "-r", "0"], dict(forced_unphased=1, minR2=0.0)),
    "phased_miss_aligned": (dict(n_samples=1024, n_variants=300, seed=106, missing_rate=0.05), ["-p", "-r", "0.05"], dict(force_phased=1, minR2=0.05)),
    "phased_miss_quirks": (dict(n_samples=1000, n_variants=300, seed=107, missing_rate=0.05), ["-p", "-r", "0.05"], dict(force_phased=1, minR2=0.05)),
    "window": (dict(n_samples=500, n_variants=1700, seed=108), ["-p", "-r", "0.1", "-w", "60000"], dict(force_phased=1, minR2=0.1, window=1, l_window=60000)),
    # the reference CLI only exposes -r and -P (lib/calc.h:99-220); maxR2/D' keep their defaults
    # auto mode (neither -p nor -u): a pair is unphased iff either variant has missing alleles (Q4)
    "auto_mixed": (dict(n_samples=60, n_variants=300, seed=110, missing_rate=0.01), ["-r", "0.05"], dict(minR2=0.05)),
    # -I: block-granular selection (lib/ld/ld.cpp:257-365). One interval -> the overlapping blocks;
    # two disjoint intervals -> the reference reads n consecutive blocks from the first overlap.
    "interval_one": (dict(n_samples=300, n_variants=2300, seed=111), ["-p", "-r", "0.1", "-I", "1:60000-110000"],
                     dict(force_phased=1, minR2=0.1)),
    "interval_two": (dict(n_samples=300, n_variants=2300, seed=112), ["-p", "-r", "0.1", "-I", "1:60000-110000", "-I", "1:200000-210000"],
                     dict(force_phased=1, minR2=0.1)),
    # -w in AUTO mode: twk_ld_slave::Calculate (ld_engine.cpp:2737-2838) has no per-pair window test, only the
    # balancer's row prune (ld_balancing.h:176-211) applies -- far more pairs than "-p -w"
    "auto_window": (dict(n_samples=200, n_variants=1700, seed=121, missing_rate=0.01), ["-r", "0.05", "-w", "60000"],
                    dict(minR2=0.05, window=1, l_window=60000)),
    # -p -m -M -w: CalculatePhasedBitmapWindow (:2441-2524) skips a pair only when the contigs differ and the
    # wrapping position difference exceeds the window; masked pairs always take PhasedRunlength (Q3 slots)
    "bitmap_window": (dict(n_samples=200, n_variants=1700, seed=123, missing_rate=0.01, two_contigs=1100),
                      ["-p", "-m", "-M", "-r", "0.05", "-w", "60000"],
                      dict(force_phased=1, bitmaps=1, minR2=0.05, window=1, l_window=60000)),
    # a .twk imported with -b 300: the window rules (row prune, block-pair abort) follow the FILE's blocks
    "window_blocks300": (dict(n_samples=300, n_variants=1900, seed=131, twk_block=300), ["-p", "-r", "0.1", "-w", "45000"],
                         dict(force_phased=1, minR2=0.1, window=1, l_window=45000, block_size=300)),
    "minp_filter": (dict(n_samples=600, n_variants=250, seed=109), ["-p", "-r", "0.05", "-P", "1e-3"],
                    dict(force_phased=1, minR2=0.05, minP=1e-3)),
}


def main():
    assert lc.have_reference(), "build the reference first: bash oracle/build_ref.sh"
    only = set(sys.argv[1:])
    for name, (skw, cli, prm) in CASES.items():
        if only and name not in only:
            continue
        skw = dict(skw)
        split = skw.pop("two_contigs", None)
        twk_block = skw.pop("twk_block", 500)
        s = tf.synth_genotypes(**skw)
        contigs = None
        if split:  # second contig restarts its positions
            s.rid[split:] = 1
            s.pos[split:] = (np.arange(s.n_variants - split) * 100).astype(np.uint32)
            contigs = [("1", 10**6), ("2", 10**6)]
        twk = os.path.join(TMP, f"g_{name}.twk")
        tf.write_twk(twk, s, contigs=contigs, block_size=twk_block)
        info = lc.run_reference_calc(twk, os.path.join(TMP, f"g_{name}"), cli, threads=4)
        recs = tf.canonical(tf.read_two(os.path.join(TMP, f"g_{name}.two")), forward_only=True)
        # pairs visited: the reference's own figure when its (racy) stderr summary parses,
        # else the restatement's count (identical whenever both are available)
        if "-I" in cli:
            # the restatement takes the variant set the reference loaded (its own LOG line)
            import re
            m = re.search(r"([\d,]+) variants from ([\d,]+) blocks", info["stderr"])
            nv = int(m.group(1).replace(",", ""))
            first = int(np.searchsorted(s.pos, (recs["packA"] >> 2).min())) // 500 * 500
            sub = tf.Synth(alleles=s.alleles[first:first + nv], pos=s.pos[first:first + nv], rid=s.rid[first:first + nv], n_samples=s.n_samples)
            _, visited = lc.calc(sub, lc.default_params(**prm))
        else:
            _, visited = lc.calc(s, lc.default_params(**prm))
        if "pairs" in info:
            assert info["pairs"] == visited, (info["pairs"], visited)
        info["pairs"] = visited
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            alleles=s.alleles, pos=s.pos, rid=s.rid, n_samples=np.int64(s.n_samples),
            records=recs.view(np.uint8), pairs=np.int64(info["pairs"]),
            cli=np.array(" ".join(cli)), params=np.array(repr(prm)),
        )
        print(name, "records", len(recs), "pairs", info["pairs"])
    if only and "fisher" not in only:
        return
    # Fisher known answers straight from the reference's kt_fisher_exact
    rng = np.random.default_rng(7)
    tabs = []
    for _ in range(400):
        n = int(rng.choice([20, 200, 2000, 5008, 20000]))
        a = rng.integers(0, n // 2 + 1)
        b = rng.integers(0, n - a + 1)
        c = rng.integers(0, n - a - b + 1)
        d = n - a - b - c
        tabs.append((a, b, c, d))
    tabs += [(5002, 5, 0, 1), (5005, 2, 0, 1), (5004, 0, 3, 1), (5006, 0, 1, 1), (5004, 3, 0, 1), (0, 0, 0, 0), (3, 0, 0, 5), (1, 1, 1, 1)]
    tabs = np.array(tabs, dtype=np.int64)
    vals = np.array([lc.reference_fisher(*t) for t in tabs])
    np.savez_compressed(os.path.join(HERE, "fisher.npz"), tables=tabs, two_sided=vals)
    print("fisher", len(tabs))


if __name__ == "__main__":
    main()
