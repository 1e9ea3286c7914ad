"""Development aid: does bounding the drift between the CTA pairs of count_umma3_kernel change its DRAM traffic at C2?
Per-tile pacing (profiling build: TWKB_PACE_MIN_KB=1, TWKB_PACE_KB=20 = one chunk per tile, TWKB_PACE_DEPTH = waves a pair
may run ahead of the slowest one) against the free-running kernel, for a few tile orders.
   python scripts/drift_probe.py [reps]     (under ncu: one launch pair per variant, in the order printed)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomahawk_b200 as tb
from tomahawk_b200 import tools

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, m = 2504, 200_000
d, mk, meta = tools.synth_device(n, m, seed=20)
# (super, inner_i, inner_j, pace_depth; 0 = free running)
grid = [(32, 0, 0, 0), (32, 0, 0, 1), (32, 0, 0, 2), (32, 0, 0, 3), (32, 0, 0, 4), (32, 0, 0, 6), (32, 0, 0, 8), (32, 0, 0, 16), (32, 0, 0, 32),
        (32, 8, 9, 2), (32, 8, 9, 4), (16, 0, 0, 2), (24, 0, 0, 2), (48, 0, 0, 2), (48, 0, 0, 4), (64, 8, 9, 4), (32, 0, 0, 0)]
ref = None
for sup, ii, ij, depth in grid:
    os.environ.update(TWKB_SUPER=str(sup), TWKB_INNER_I=str(ii), TWKB_INNER_J=str(ij), TWKB_PACE_KB="20", TWKB_PACE_DEPTH=str(depth),
                      TWKB_PACE_MIN_KB="1" if depth else "100000")
    eng = tb.Engine(force_phased=1, minR2=0.1, kernel=tb.KERNEL_UMMA_FP4, profiling=True)
    eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
    ms = []
    for _ in range(reps):
        eng.compute_resident()
        st = eng.stats()
        ms.append(st.ms_count_kernel)
    if ref is None:
        ref = st.records_out
    print("super %2d inner %dx%-2d pace depth %d: count_ms %s records %d %s" % (
        sup, ii, ij, depth, " ".join("%.2f" % x for x in ms), st.records_out, "OK" if st.records_out == ref else "MISMATCH"), flush=True)
    eng.close()
