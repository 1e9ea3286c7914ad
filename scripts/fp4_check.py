"""Development aid: e2m1 (kind::mxf4) tensor-core count kernel vs the int8 and POPC kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth

KS = (("popc", tb.KERNEL_POPC), ("i8", tb.KERNEL_UMMA), ("fp4", tb.KERNEL_UMMA_FP4))

def run(n_samples, n_variants, seed, minR2, kernels=KS, dense=False):
    if dense:  # adversarial: nearly every haplotype carries the alt allele -> counts near 2N
        rng = np.random.default_rng(seed)
        nb = 2 * n_samples
        words = (nb + 127) // 128 * 2
        data = np.zeros((n_variants, words), np.uint64)
        ac = np.zeros(n_variants, np.uint32)
        for v in range(n_variants):
            bits = np.zeros(words * 64, np.uint8)
            bits[:nb] = rng.random(nb) < (0.97 if v % 3 else 0.5)
            bits[0] = 0
            ac[v] = bits.sum()
            data[v] = np.packbits(bits, bitorder="little").view(np.uint64)
        mask = None
        meta = np.zeros(n_variants, tb.VARIANT_DTYPE)
        meta["pos"] = 100 * (1 + np.arange(n_variants)); meta["ac"] = ac; meta["hwe"] = 1.0; meta["gt_phase"] = 1
    else:
        s = synth.synth_genotypes(n_samples, n_variants, seed=seed)
        data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
    out = {}
    for name, k in kernels:
        eng = tb.Engine(force_phased=1, minR2=minR2, kernel=k)
        eng.load(n_samples, data, mask, meta)
        recs = eng.compute()
        t = time.time(); recs = eng.compute(); dt = time.time() - t
        st = eng.stats()
        order = np.lexsort((recs["packB"], recs["packA"]))
        out[name] = recs[order]
        print(f"  {name}: kernel_used={st.kernel_used} records={len(recs)} screened={st.pairs_screened} count_ms={st.ms_count_kernel:.3f} stats_ms={st.ms_stats_kernel:.3f} wall={dt:.3f}", flush=True)
        eng.close()
    ref = out[kernels[0][0]]
    same = all(len(ref) == len(o) and np.array_equal(ref.view(np.uint8), o.view(np.uint8)) for o in out.values())
    print(f"[{n_samples}x{n_variants} r2>={minR2} dense={dense}] identical={same}", flush=True)
    return same

ok = True
ok &= run(2504, 300, 1, 0.0)
ok &= run(2504, 1500, 2, 0.1)
ok &= run(100, 700, 3, 0.05)
ok &= run(777, 1000, 4, 0.02)
ok &= run(2504, 700, 5, 0.0, dense=True)
ok &= run(60000, 600, 6, 0.0, dense=True)      # counts up to ~116,000
ok &= run(500000, 520, 7, 0.3, kernels=KS[1:], dense=True)   # counts up to ~970,000 (1M haplotypes)
print("FP4 OK" if ok else "FP4 MISMATCH", flush=True)
if len(sys.argv) > 1:
    run(2504, int(sys.argv[1]), 20, 0.1, kernels=KS[1:])
