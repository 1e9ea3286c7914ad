#!/bin/bash
# N-GPU bench run; $1 = N
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err
echo "n$N rc=$?"; grep -v "^$" gpurun_out/g_bench_n$N.err | tail -6
python - <<PY
import json
f="g_bench_n$N"
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[0])
    print(f, "value %.4g ms %.2f e2e %.4g (%.1f ms) frac %.3f count_ms %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"]["ms_count_kernel_per_step"]))
    print("  e2e", d["e2e"]); print("  clocks", d["clocks"])
    for x in d.get("extra_configs", []):
        print("  extra", json.dumps(x)[:2500])
except Exception as e:
    print(f, "FAILED", e)
PY
