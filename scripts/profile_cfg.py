"""Development aid: one resident pass of a named configuration, for ncu captures (device-generated data).
   python scripts/profile_cfg.py c3|c4|c1 [variants]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomahawk_b200 as tb
from tomahawk_b200 import tools
cfg = sys.argv[1]
if cfg == "c3":
    n, m = 10_000, int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    d, mk, meta = tools.synth_device(n, m, seed=20, missing_rate=0.05)
    eng = tb.Engine(forced_unphased=1, minR2=0.1)
    eng.load_device(n, m, d.data_ptr(), mk.data_ptr(), d.shape[1], meta)
elif cfg == "c4":
    n, m = 500_000, int(sys.argv[2]) if len(sys.argv) > 2 else 16_000
    d, mk, meta = tools.synth_device(n, m, seed=20)
    eng = tb.Engine(force_phased=1, minR2=0.1)
    eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
else:
    n, m = 2504, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    d, mk, meta = tools.synth_device(n, m, seed=20)
    eng = tb.Engine(force_phased=1, minR2=0.0)
    eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
for _ in range(2):
    eng.compute_resident()
    st = eng.stats()
    print(cfg, "count_ms %.2f sparse_ms %.2f stats_ms %.2f pairs %.4g records %d" % (st.ms_count_kernel, st.ms_sparse_kernel, st.ms_stats_kernel, st.pairs_visited, st.records_out), flush=True)
eng.close()
