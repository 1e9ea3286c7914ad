"""Development aid: why is the count kernel slower in the N = 2 bench? Single GPU; separates the data (device
generator vs numpy) from the partition (part 0 of 2) and the load path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomahawk_b200 as tb
from tomahawk_b200 import tools

def run(name, n, m, parts, part, data_t, meta):
    eng = tb.Engine(force_phased=1, minR2=0.1, part_index=part, part_count=parts)
    eng.load_device(n, m, data_t.data_ptr(), None, data_t.shape[1], meta)
    ms = []
    for _ in range(4):
        eng.compute_resident(); st = eng.stats(); ms.append(st.ms_count_kernel)
    print(f"{name:40s} count_ms={min(ms[1:]):7.2f} pairs={st.pairs_visited:.4g} screened={st.pairs_screened} records={st.records_out} launches={st.count_launches}", flush=True)
    eng.close()

n = 2504
d, _, meta = tools.synth_device(n, 200000, seed=20)
run("device data 200k, whole", n, 200000, 1, 0, d, meta)
ac = meta["ac"]; print("device AF: frac ac<=2", (ac <= 2).mean(), "median ac", np.median(ac), "frac ac>2500", (ac > 2500).mean())
d2, _, meta2 = tools.synth_device(n, 282843, seed=20)
run("device data 282843, whole", n, 282843, 1, 0, d2, meta2)
run("device data 282843, part 0/2", n, 282843, 2, 0, d2, meta2)
run("device data 282843, part 1/2", n, 282843, 2, 1, d2, meta2)
from tomahawk_b200 import synth
s = synth.synth_genotypes(n, 100000, seed=20)
print("numpy AF: frac ac<=2", (s.ac <= 2).mean(), "median ac", np.median(s.ac))
