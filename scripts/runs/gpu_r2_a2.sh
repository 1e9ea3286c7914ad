#!/bin/bash
# round 2: bench at N = 2 (weak-scaled configs[1], configs[3] strong-scaled with K-sweep pacing, configs[4] slice position-sharded)
# + the multi-GPU tests (sliced loads, CLI -g 0,1, shards over two devices)
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/a2_bench_n2.json 2> gpurun_out/a2_bench_n2.err
echo "n2 rc=$?"; tail -n 4 gpurun_out/a2_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/a2_bench_n2.json'))
print('primary', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
for x in d.get('extra_configs',[]): print(json.dumps(x)[:2500]); print()
PY
( time timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_shards.py -m gpu -q ) > gpurun_out/a2_pytest_mg.log 2>&1
tail -n 5 gpurun_out/a2_pytest_mg.log
