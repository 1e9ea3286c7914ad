"""compute-sanitizer target (round 2 additions): the device-side consumers of the resident records -- decay, aggregate (range
+ raster passes), the radix sorter -- and a position-sharded -w run with the rare-variant path, all on small inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from tomahawk_b200 import synth
s = synth.synth_genotypes(300, 700, seed=1)
s.rid[400:] = 1
s.pos[400:] = (np.arange(300) * 100).astype(np.uint32)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
eng = tb.Engine(force_phased=1, minR2=0.02)
eng.load(s.n_samples, data, mask, meta)
print("decay", eng.compute_decay(50_000, 20)[1].sum(), flush=True)
bins, lay, _ = eng.compute_aggregate("r2", 40, 25, [100_000, 80_000])
print("aggregate", int(bins["n"].sum()), lay, flush=True)
recs = eng.compute_sorted()
print("sorted", len(recs), bool(np.all(np.diff(recs["ridA"].astype(np.int64)) >= 0)), flush=True)
eng.close()
s = synth.synth_genotypes(900, 1500, seed=6, rare_fraction=0.8)
data, mask = synth.pack_bits(s); meta = synth.variant_meta(s)
first = np.arange(0, 1500 + 150, 150, dtype=np.uint32); first[-1] = 1500
own, halo = tb.plan_shards(first, meta, 30_000, 2)
for k in range(2):
    v0, v1 = int(first[own[k]]), int(first[halo[k]])
    e = tb.Engine(force_phased=1, minR2=0.05, window=1, l_window=30_000, sparse_max_words=6, shard_blocks=int(own[k + 1] - own[k]))
    e.load(s.n_samples, data[v0:v1], None, meta[v0:v1])
    e.set_blocks(first[own[k]:halo[k]] - v0)
    print("shard", k, len(e.compute()), flush=True)
    e.close()
