"""Development aid: DRAM traffic / time of the C2 count kernel against the tile order and the L2 eviction hints of the
operand loads (profiling build: TWKB_SUPER, TWKB_INNER_I/J, TWKB_L2_HINT_A/B with 0 normal, 1 evict_first, 2 evict_last).
   python scripts/l2_sweep.py [reps] [variants]            # times
   ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:count_umma3 --csv \\
       --log-file out.csv python scripts/l2_sweep.py 1     # one launch per variant, in the order printed"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomahawk_b200 as tb
from tomahawk_b200 import tools

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
n = 2504
d, mk, meta = tools.synth_device(n, m, seed=20)
# (super, inner_i, inner_j, hint_a, hint_b)
grid = [(32, 0, 0, 0, 0), (16, 0, 0, 0, 0), (24, 0, 0, 0, 0), (48, 0, 0, 0, 0), (64, 0, 0, 0, 0),
        (32, 8, 9, 0, 0), (32, 6, 12, 0, 0), (32, 4, 16, 0, 0), (48, 8, 9, 0, 0), (64, 8, 9, 0, 0), (64, 8, 8, 0, 0), (96, 8, 9, 0, 0),
        (32, 0, 0, 1, 0), (32, 0, 0, 0, 2), (32, 0, 0, 1, 2), (32, 0, 0, 2, 0), (32, 0, 0, 2, 1),
        (32, 8, 9, 1, 0), (32, 8, 9, 2, 0), (48, 8, 9, 2, 0), (64, 8, 9, 2, 0), (64, 8, 9, 2, 1),
        (32, 0, 0, 0, 0)]
ref = None
for sup, ii, ij, ha, hb in grid:
    os.environ.update(TWKB_SUPER=str(sup), TWKB_INNER_I=str(ii), TWKB_INNER_J=str(ij), TWKB_L2_HINT_A=str(ha), TWKB_L2_HINT_B=str(hb))
    eng = tb.Engine(force_phased=1, minR2=0.1, kernel=tb.KERNEL_UMMA_FP4, profiling=True)
    eng.load_device(n, m, d.data_ptr(), None, d.shape[1], meta)
    ms = []
    for _ in range(reps):
        eng.compute_resident()
        st = eng.stats()
        ms.append(st.ms_count_kernel)
    if ref is None:
        ref = st.records_out
    print("super %2d inner %dx%-2d hintA %d hintB %d: count_ms %s records %d %s" % (
        sup, ii, ij, ha, hb, " ".join("%.2f" % x for x in ms), st.records_out, "OK" if st.records_out == ref else "MISMATCH"), flush=True)
    eng.close()
