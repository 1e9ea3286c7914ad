"""Development aid: decomposition of the persistent tensor-core kernel's time at C2 with the profiling build
(libtwkb_prof.so, TWKB_DEBUG_FLAGS; results are invalid with any switch on):
  1 = no operand traffic after the first ring fill   2 = no epilogue (no TMEM loads, no screen)
  8 = epilogue: TMEM loads only                      16 = epilogue: screen arithmetic only (no TMEM loads)
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
s = synth.synth_genotypes(2504, M, seed=20)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
del s
names = {0: "full kernel", 1: "no operand traffic", 2: "no epilogue", 3: "MMA only", 8: "epilogue = TMEM loads only",
         16: "epilogue = screen arithmetic only", 9: "TMEM loads only, no operand traffic"}
for flags in (0, 2, 3, 8, 16, 9, 0):
    os.environ["TWKB_DEBUG_FLAGS"] = str(flags)
    eng = tb.Engine(force_phased=1, minR2=0.1, kernel=tb.KERNEL_UMMA_FP4, profiling=True)
    eng.load(2504, data, mask, meta)
    ms = []
    for _ in range(4):
        eng.compute_resident(); ms.append(eng.stats().ms_count_kernel)
    print(f"flags={flags:2d} {names[flags]:38s}: count_ms={min(ms[1:]):.2f} (runs {['%.1f' % x for x in ms]})", flush=True)
    eng.close()
