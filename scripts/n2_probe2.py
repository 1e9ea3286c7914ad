"""Development aid: ablations of the count kernel on DEVICE-generated data (profiling build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tomahawk_b200 as tb
from tomahawk_b200 import tools, synth
n = 2504
d, _, meta = tools.synth_device(n, 200000, seed=20)
s = synth.synth_genotypes(n, 200000, seed=20)
data, _ = synth.pack_bits(s); metan = synth.variant_meta(s)
dn = torch.from_numpy(data.view(np.int64)).cuda()
for name, dt, mt in (("device", d, meta), ("numpy", dn, metan)):
    for flags in (0, 2, 3, 8, 16):
        os.environ["TWKB_DEBUG_FLAGS"] = str(flags)
        eng = tb.Engine(force_phased=1, minR2=0.1, profiling=True)
        eng.load_device(n, 200000, dt.data_ptr(), None, dt.shape[1], mt)
        ms = []
        for _ in range(3):
            eng.compute_resident(); st = eng.stats(); ms.append(st.ms_count_kernel)
        print(f"{name} flags={flags:2d} count_ms={min(ms[1:]):7.2f} screened={st.pairs_screened}", flush=True)
        eng.close()
# row statistics that could matter to the operand pipeline / epilogue
bits_d = d.cpu().numpy().view(np.uint64)
print("device rows: distinct rows", len(np.unique(bits_d[:20000], axis=0)), "of 20000; numpy:", len(np.unique(data[:20000], axis=0)))
