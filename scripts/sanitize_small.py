"""compute-sanitizer target: one small run of every count-kernel mode (tensor and POPC), the rare-variant path, the
R2 >= 0 direct path, the device run-length decoder and the statistics kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
CASES = [("phased", dict(n_samples=300, n_variants=700, seed=1), dict(force_phased=1, minR2=0.1)),
         ("phased_r0", dict(n_samples=200, n_variants=300, seed=2), dict(force_phased=1, minR2=0.0)),
         ("phased_missing", dict(n_samples=256, n_variants=500, seed=3, missing_rate=0.05), dict(force_phased=1, minR2=0.1)),
         ("unphased", dict(n_samples=300, n_variants=500, seed=4), dict(forced_unphased=1, minR2=0.1)),
         ("unphased_missing", dict(n_samples=300, n_variants=500, seed=5, missing_rate=0.05), dict(forced_unphased=1, minR2=0.1)),
         ("sparse", dict(n_samples=1000, n_variants=600, seed=6, rare_fraction=0.8), dict(force_phased=1, minR2=0.2, sparse_max_words=6))]
kernels = [tb.KERNEL_AUTO, tb.KERNEL_POPC] if "--popc" in sys.argv else [tb.KERNEL_AUTO]
only = [a for a in sys.argv[1:] if not a.startswith("--")]   # optional: names of the cases to run
for name, skw, prm in CASES:
    if only and name not in only:
        continue
    s = synth.synth_genotypes(**skw)
    data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
    for k in kernels:
        eng = tb.Engine(kernel=k, **prm)
        eng.load(s.n_samples, data, mask, meta)
        recs = eng.compute()
        print(name, "kernel", eng.stats().kernel_used, "records", len(recs), flush=True)
        eng.close()
