"""Development aid: where does the screen arithmetic's cost come from? (profiling build, C2)
 16 = screen only; 16|32 = screen arithmetic without its shared-memory loads; 16|64 = the loads without the arithmetic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
s = synth.synth_genotypes(2504, M, seed=20)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
del s
for flags in (0, 2, 16, 48, 80, 32, 64):
    os.environ["TWKB_DEBUG_FLAGS"] = str(flags)
    eng = tb.Engine(force_phased=1, minR2=0.1, kernel=tb.KERNEL_UMMA_FP4, profiling=True)
    eng.load(2504, data, mask, meta)
    ms = []
    for _ in range(4):
        eng.compute_resident(); ms.append(eng.stats().ms_count_kernel)
    print(f"flags={flags:3d}: count_ms={min(ms[1:]):.2f} (runs {['%.1f' % x for x in ms]})", flush=True)
    eng.close()
