"""Host timing of the .two sorter against the reference's `sort` binary (oracle/_ref/tomahawk_sort) on a
C2-sized result file (1.1 M records, shuffled).  python scripts/sort_bench.py [records] [threads]"""
import json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf
from tests.helpers import load_golden

n_want = int(sys.argv[1]) if len(sys.argv) > 1 else 1_109_434
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 4)
s, recs, prm, pairs, _ = load_golden("phased_r0")
tmp = tempfile.mkdtemp(prefix="twkb_sort_")
twk = os.path.join(tmp, "g.twk")
tf.write_twk(twk, s, contigs=[("1", 2**30)])
reps = (n_want + len(recs) - 1) // len(recs)
big = np.tile(recs, reps)[:n_want].copy()
shift = (np.arange(n_want) // len(recs)).astype(np.uint32) * np.uint32(1_000_000)   # distinct keys per copy
big["packA"] = ((big["packA"] >> 2) + shift) << 2
big["packB"] = ((big["packB"] >> 2) + shift) << 2
big = big[np.random.default_rng(1).permutation(n_want)]
src = os.path.join(tmp, "u.two")
w = tb.TwoWriter(src, tb.TwkFile(twk), "sort_bench", c_level=1, b_size=10000, n_threads=threads)
w.add(big); w.close()
t0 = time.perf_counter()
n = tb.sort_two(src, os.path.join(tmp, "ours.two"), 1, threads)
ours = time.perf_counter() - t0
row = {"records": int(n), "threads": threads, "ours_s": ours, "records_per_s": n / ours, "in_bytes": os.path.getsize(src)}
if os.path.exists(lc.REF_SORT):
    t0 = time.perf_counter()
    r = subprocess.run([lc.REF_SORT, "sort", "-i", src, "-o", os.path.join(tmp, "ref.two"), "-t", str(threads)], capture_output=True, text=True)
    row["reference_s"] = time.perf_counter() - t0
    a, b = tf.read_two(os.path.join(tmp, "ref.two")), tf.read_two(os.path.join(tmp, "ours.two"))
    row["identical_to_reference"] = bool(r.returncode == 0 and np.array_equal(a.view(np.uint8), b.view(np.uint8)))
print(json.dumps(row))
for f in os.listdir(tmp):
    os.unlink(os.path.join(tmp, f))
os.rmdir(tmp)
