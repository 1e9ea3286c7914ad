"""Quick GPU parity check (development aid): CUDA path vs the CPU restatement."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import twk_format as tf, ldcore as lc
import tomahawk_b200 as tb

def cmp_records(name, ref, got, exact_stats=True):
    ref = tf.canonical(ref, forward_only=False); got = tf.canonical(got, forward_only=False)
    ok = True
    if len(ref) != len(got):
        ka = set(zip(ref['packA'].tolist(), ref['packB'].tolist())); kb = set(zip(got['packA'].tolist(), got['packB'].tolist()))
        print(f"[{name}] COUNT MISMATCH ref={len(ref)} got={len(got)} only_ref={len(ka-kb)} only_got={len(kb-ka)}", list(ka-kb)[:3], list(kb-ka)[:3])
        return False
    msgs = []
    for f in tf.TWO_DTYPE.names:
        a, b = ref[f], got[f]
        if a.dtype.kind == 'f':
            ex = np.array_equal(a, b)
            with np.errstate(all='ignore'):
                rel = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))) if len(a) else 0.0
            msgs.append(f"{f}:{'=' if ex else f'{rel:.2g}'}")
            if f == 'cnt' and not ex and exact_stats: ok = False
        else:
            eq = np.array_equal(a, b)
            msgs.append(f"{f}:{'=' if eq else 'DIFF'}")
            ok = ok and eq
    print(f"[{name}] n={len(ref)} " + " ".join(msgs))
    return ok

def run(name, s, kernel=tb.KERNEL_POPC, **kw):
    pk = {k: v for k, v in kw.items() if k in ('force_phased','forced_unphased','minR2','window','l_window','minP','maxR2','minDprime','maxDprime')}
    t = time.time(); ref, visited = lc.calc(s, lc.default_params(**pk)); t_cpu = time.time() - t
    data, mask = tf.pack_bits(s); meta = lc.variant_meta(s)
    eng = tb.Engine(kernel=kernel, **pk)
    eng.load(s.n_samples, data, mask, meta)
    t = time.time(); got = eng.compute(); t_gpu = time.time() - t
    st = eng.stats()
    ok = cmp_records(name, ref, got)
    print(f"   visited cpu={visited} gpu={st.pairs_visited} screened={st.pairs_screened} records={st.records_out} "
          f"count_ms={st.ms_count_kernel:.3f} stats_ms={st.ms_stats_kernel:.3f} cpu_s={t_cpu:.2f} gpu_s={t_gpu:.3f} kernel={st.kernel_used}")
    # exact counts for every pair
    cands = eng.debug_candidates(True)
    print(f"   debug candidates: {len(cands)}")
    if not ok or os.environ.get("VERBOSE"):
        key = {(int(c['i']), int(c['j'])): c for c in cands}
        refc = tf.canonical(ref, forward_only=False); gotc = tf.canonical(got, forward_only=False)
        gk = {(int(r['packA']), int(r['packB'])): r for r in gotc}
        shown = 0
        step = int(s.pos[1] - s.pos[0])
        for r in refc:
            k = (int(r['packA']), int(r['packB']))
            g = gk.get(k)
            bad = g is None or not np.array_equal(r['cnt'], g['cnt']) or abs(r['R2']-g['R2']) > 1e-9*abs(r['R2'])
            if bad and shown < 6:
                i, j = (k[0] >> 2)//step, (k[1] >> 2)//step
                c = key.get((i, j))
                print("   MISMATCH pair", i, j, "table", None if c is None else c['c'].tolist(), "mode", None if c is None else int(c['mode']))
                print("      ref", r['cnt'].tolist(), r['D'], r['R2'], r['P'], int(r['controller']))
                if g is not None: print("      got", g['cnt'].tolist(), g['D'], g['R2'], g['P'], int(g['controller']))
                shown += 1
    eng.close()
    return ok

if __name__ == "__main__":
    allok = True
    if os.environ.get("ONLY_UNPHASED"):
        allok &= run("unphased miss", tf.synth_genotypes(1000, 700, seed=5, missing_rate=0.05), forced_unphased=1, minR2=0.1)
        allok &= run("unphased nomiss r0", tf.synth_genotypes(1000, 700, seed=6), forced_unphased=1, minR2=0.0)
        sys.exit(0)
    allok &= run("phased r0.1", tf.synth_genotypes(2504, 1500, seed=3), force_phased=1, minR2=0.1)
    allok &= run("phased r0", tf.synth_genotypes(2504, 300, seed=4), force_phased=1, minR2=0.0)
    allok &= run("unphased miss", tf.synth_genotypes(1000, 700, seed=5, missing_rate=0.05), forced_unphased=1, minR2=0.1)
    allok &= run("unphased nomiss r0", tf.synth_genotypes(1000, 700, seed=6), forced_unphased=1, minR2=0.0)
    allok &= run("phased miss 2N%128==0", tf.synth_genotypes(1024, 700, seed=8, missing_rate=0.05), force_phased=1, minR2=0.05)
    allok &= run("window", tf.synth_genotypes(500, 3000, seed=9), force_phased=1, minR2=0.1, window=1, l_window=60000)
    print("ALL OK" if allok else "SOME FAILED")
