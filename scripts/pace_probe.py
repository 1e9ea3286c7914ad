"""Development aid: sweep of the K-sweep pacing parameters of count_umma3_kernel (profiling build reads TWKB_PACE_KB /
TWKB_PACE_DEPTH; TWKB_SUPER = super-tile edge) on a biobank-shaped matrix (1,000,000 haplotypes, device-generated).
   python scripts/pace_probe.py [variants]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomahawk_b200 as tb
from tomahawk_b200 import tools

n, m = 500_000, int(sys.argv[1]) if len(sys.argv) > 1 else 16_000
d, mk, meta = tools.synth_device(n, m, seed=20)
eng = tb.Engine(force_phased=1, minR2=0.1, profiling=True)
eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
ref = None
grid = [(0, 0, 8), (16, 2, 8), (16, 1, 8), (16, 4, 8), (8, 2, 8), (32, 2, 8), (64, 2, 8), (16, 2, 9), (16, 2, 12), (16, 2, 6), (32, 1, 8), (0, 0, 8)]
if len(sys.argv) > 2:
    grid = [tuple(int(x) for x in g.split(",")) for g in sys.argv[2:]]
for kb, depth, sup in grid:
    os.environ["TWKB_PACE_KB"] = str(max(kb, 1))
    os.environ["TWKB_PACE_DEPTH"] = str(depth)
    os.environ["TWKB_SUPER"] = str(sup)
    ms = []
    for _ in range(3):
        eng.compute_resident()
        st = eng.stats()
        ms.append(st.ms_count_kernel)
    if ref is None:
        ref = st.records_out
    print("pace_kb %3d depth %d super %2d: count_ms %s sparse_ms %.2f records %d %s" % (
        kb, depth, sup, " ".join("%.2f" % x for x in ms), st.ms_sparse_kernel, st.records_out, "OK" if st.records_out == ref else "MISMATCH"), flush=True)
eng.close()
