"""Development aid: planes tensor kernels (masked phased / unphased) vs the POPC kernel, and C3-shaped timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth

def run(n_samples, n_variants, seed, prm, missing=0.0, kernels=(tb.KERNEL_POPC, tb.KERNEL_AUTO), reps=2):
    s = synth.synth_genotypes(n_samples, n_variants, seed=seed, missing_rate=missing)
    data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
    out = {}
    for k in kernels:
        eng = tb.Engine(kernel=k, **prm)
        eng.load(n_samples, data, mask, meta)
        for _ in range(reps):
            eng.compute_resident()
            st = eng.stats()
            print(f"  kernel={k} used={st.kernel_used} np={st.n_planes} pairs={st.pairs_visited} screened={st.pairs_screened} records={st.records_out} "
                  f"count_ms={st.ms_count_kernel:.2f} stats_ms={st.ms_stats_kernel:.2f} total_ms={st.ms_device_total:.2f} launches={st.count_launches}", flush=True)
        if n_variants <= 20000:
            recs = eng.compute()
            order = np.lexsort((recs["packB"], recs["packA"]))
            out[k] = recs[order]
        eng.close()
    if len(out) > 1:
        ks = list(out)
        same = all(len(out[ks[0]]) == len(out[k]) and np.array_equal(out[ks[0]].view(np.uint8), out[k].view(np.uint8)) for k in ks)
        print(f"[{n_samples}x{n_variants} {prm} miss={missing}] identical={same}", flush=True)
        return same
    return True

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        ok = run(1000, 1500, 1, dict(forced_unphased=1, minR2=0.1), missing=0.05)
        ok &= run(1000, 1500, 2, dict(forced_unphased=1, minR2=0.1))
        ok &= run(1024, 1500, 3, dict(force_phased=1, minR2=0.1), missing=0.05)
        print("PLANES OK" if ok else "PLANES MISMATCH")
    elif which == "c3":
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
        ks = (tb.KERNEL_AUTO,) if len(sys.argv) <= 3 else (tb.KERNEL_POPC, tb.KERNEL_AUTO)
        run(10000, M, 20, dict(forced_unphased=1, minR2=0.1), missing=0.05, kernels=ks)
