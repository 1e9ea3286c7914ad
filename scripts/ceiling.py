"""Development aid: decomposition of the persistent tensor-core kernel's time at C2
(TWKB_DEBUG_FLAGS: 1 = no operand traffic after the first ring fill, 2 = no epilogue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
s = synth.synth_genotypes(2504, M, seed=20)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
for name, k in (("fp4", tb.KERNEL_UMMA_FP4),):
    for flags in (0, 1, 2, 3):
        os.environ["TWKB_DEBUG_FLAGS"] = str(flags)
        eng = tb.Engine(force_phased=1, minR2=0.1, kernel=k)
        eng.load(2504, data, mask, meta)
        ms = []
        for _ in range(4):
            eng.compute_resident(); ms.append(eng.stats().ms_count_kernel)
        print(f"{name} flags={flags}: count_ms={min(ms[1:]):.2f} (runs {['%.1f' % x for x in ms]})", flush=True)
        eng.close()
