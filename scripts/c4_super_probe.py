"""Development aid: super-tile edge of the tile order at biobank row length (1 M haplotypes), one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomahawk_b200 as tb
from tomahawk_b200 import tools
n, m = 500_000, int(sys.argv[1]) if len(sys.argv) > 1 else 24_000
d, _, meta = tools.synth_device(n, m, seed=20)
for sup in (32, 16, 12, 9, 8, 6, 0):
    if sup: os.environ["TWKB_SUPER"] = str(sup)
    else: os.environ.pop("TWKB_SUPER", None)
    eng = tb.Engine(force_phased=1, minR2=0.1, sparse_max_words=-1, profiling=True)
    eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
    ms = []
    for _ in range(2):
        eng.compute_resident(); st = eng.stats(); ms.append(st.ms_count_kernel)
    tf = st.pairs_visited * 2e6 / (min(ms) * 1e-3) / 1e12
    print(f"super={sup:2d} count_ms={min(ms):8.2f}  {tf:7.0f} TFLOP/s ({tf / 9000:.3f} of nominal) records={st.records_out}", flush=True)
    eng.close()
