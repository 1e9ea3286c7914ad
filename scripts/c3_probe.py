"""Development aid: the planes (unphased / masked) tensor kernel on a C3-shaped problem.
  python scripts/c3_probe.py [variants] [samples] [missing] [minR2]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomahawk_b200 as tb
from tomahawk_b200 import synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
r2 = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
s = synth.synth_genotypes(N, M, seed=20, missing_rate=miss)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
eng = tb.Engine(forced_unphased=1, minR2=r2)
eng.load(N, data, mask, meta)
for it in range(3):
    eng.compute_resident()
    st = eng.stats()
    print(f"run {it}: pairs={st.pairs_visited} screened={st.pairs_screened} ({st.pairs_screened / st.pairs_visited:.2e}) records={st.records_out} "
          f"count_ms={st.ms_count_kernel:.2f} stats_ms={st.ms_stats_kernel:.2f} launches={st.count_launches}", flush=True)
