#!/bin/bash
# round 2: bench at N = 8 with the final build (per-tile pacing): weak-scaled configs[1] + configs[3] strong scaling + configs[4] slice
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/d8_bench_n8.json 2> gpurun_out/d8_bench_n8.err
echo "n8 rc=$?"; grep real gpurun_out/d8_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d8_bench_n8.json').read().strip().splitlines()[-1])
print('primary', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['parts_last_step'], 'frac', d['roofline']['frac'])
for x in d.get('extra_configs',[]): print(x['baseline_config'], 'value %.4g ms %.1f count %.1f sparse %.1f e2e %.4g (%.1f ms) frac %.3f' % (x['value'], x['ms_per_step'], x['ms_count_kernel_per_step'], x['ms_sparse_kernel_per_step'], x['e2e']['value'], x['e2e']['ms_per_step'], x['roofline']['frac']), x.get('parity_spot'))
PY
