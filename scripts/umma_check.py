"""Development aid: tensor-core count kernel vs the LOP3+POPC kernel (exact candidate dump)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth

def run(n_samples, n_variants, seed, minR2):
    s = synth.synth_genotypes(n_samples, n_variants, seed=seed)
    data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
    out = {}
    for name, k in (("popc", tb.KERNEL_POPC), ("umma", tb.KERNEL_UMMA)):
        eng = tb.Engine(force_phased=1, minR2=minR2, kernel=k)
        eng.load(s.n_samples, data, mask, meta)
        t = time.time(); recs = eng.compute(); dt = time.time() - t
        st = eng.stats()
        order = np.lexsort((recs["packB"], recs["packA"]))
        out[name] = recs[order]
        print(f"  {name}: kernel_used={st.kernel_used} records={len(recs)} screened={st.pairs_screened} count_ms={st.ms_count_kernel:.3f} stats_ms={st.ms_stats_kernel:.3f} wall={dt:.3f}")
        eng.close()
    same = len(out["popc"]) == len(out["umma"]) and np.array_equal(out["popc"].view(np.uint8), out["umma"].view(np.uint8))
    print(f"[{n_samples}x{n_variants} r2>={minR2}] identical={same}")
    if not same:
        a, b = out["popc"], out["umma"]
        print("   first popc", a[:2]); print("   first umma", b[:2])
    return same

ok = True
ok &= run(2504, 300, 1, 0.0)
ok &= run(2504, 1500, 2, 0.1)
ok &= run(100, 700, 3, 0.05)
ok &= run(777, 1000, 4, 0.02)
print("UMMA OK" if ok else "UMMA MISMATCH")
