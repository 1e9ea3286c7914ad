"""Development aid: SM clock / power while the C2 count kernel runs back to back, with and without its
epilogue (TWKB_DEBUG_FLAGS=2) -- separates the clock cost of the epilogue under the power cap from
tensor-pipe bubbles."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
s = synth.synth_genotypes(2504, 200000, seed=20)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
for flags in (0, 2, 0, 2):
    os.environ["TWKB_DEBUG_FLAGS"] = str(flags)
    eng = tb.Engine(force_phased=1, minR2=0.1)
    eng.load(2504, data, mask, meta)
    eng.compute_resident()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                         stdout=subprocess.PIPE, text=True)
    lines = []
    th = threading.Thread(target=lambda: lines.extend(p.stdout), daemon=True); th.start()
    ms = []
    t0 = time.time()
    while time.time() - t0 < 3.0:
        eng.compute_resident(); ms.append(eng.stats().ms_count_kernel)
    p.terminate(); time.sleep(0.2)
    vals = [tuple(float(x) for x in ln.split(",")) for ln in lines if "," in ln]
    clk = np.array([v[0] for v in vals[5:]]); pw = np.array([v[1] for v in vals[5:]])
    print(f"flags={flags}: count_ms median {np.median(ms):.2f} (n={len(ms)}), sm clock median {np.median(clk):.0f} MHz "
          f"(min {clk.min():.0f}), power median {np.median(pw):.0f} W (max {pw.max():.0f}), samples {len(clk)}", flush=True)
    eng.close()
