#!/bin/bash
# round 2: incremental record drain + host metadata team: GPU suite, bench N=1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -n 6 gpurun_out/c2_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/c2_bench_n1.json 2> gpurun_out/c2_bench_n1.err
echo "bench rc=$?"; tail -n 3 gpurun_out/c2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c2_bench_n1.json').read().strip().splitlines()[-1])
print('primary value %.4g ms %.3f count %.3f stats %.3f e2e %.4g (%.3f ms) frac %.4f traffic %s clocks %s' % (d['value'], d['ms_per_step'], d['config']['ms_count_kernel_per_step'], d['config']['ms_stats_kernel_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks']))
print(d['e2e'])
print('parity', d.get('parity',{}).get('checked'), d.get('parity',{}).get('only_ref'), d.get('parity',{}).get('only_gpu'))
for x in d.get('extra_configs',[]): print(x['baseline_config'], 'value %.4g ms %.2f count %.2f stats %.2f e2e %.4g (%.1f ms) frac %.3f parity %s' % (x['value'], x['ms_per_step'], x['ms_count_kernel_per_step'], x['ms_stats_kernel_per_step'], x['e2e']['value'], x['e2e']['ms_per_step'], x['roofline']['frac'], (x.get('parity') or {}).get('checked')))
PY
