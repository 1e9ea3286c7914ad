#!/bin/bash
# round 2: position-shard tests + configs[4] slice at N = 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_shards.py tests/test_cli.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -n 12 gpurun_out/t_pytest.log
( time timeout 600 python bench.py --c5 --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/t_bench_c5.json 2> gpurun_out/t_bench_c5.err
echo "bench rc=$?"; tail -n 5 gpurun_out/t_bench_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/t_bench_c5.json'))
print('primary', d['value'], d['ms_per_step'], d['roofline']['frac'])
for x in d.get('extra_configs',[]): print(json.dumps(x)[:3000])
PY
