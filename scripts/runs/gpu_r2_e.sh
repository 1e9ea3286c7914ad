#!/bin/bash
# round 2, call E (2 GPUs): multi-GPU tests, bench at N = 2 (NCCL sliced loads + configs[3] strong scaling), bench at N = 1 with extra_configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/e_pytest_mg.log 2>&1
tail -3 gpurun_out/e_pytest_mg.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
echo "n2 rc=$?"; tail -5 gpurun_out/e_bench_n2.err
( time timeout 900 python bench.py ) > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
echo "n1 rc=$?"; tail -5 gpurun_out/e_bench_n1.err
python - <<'PY'
import json
for f in ("e_bench_n1","e_bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[0])
        print(f, "value %.4g ms %.2f e2e %.4g (%.1f ms) frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
        print("  parity", d.get("parity")); print("  cpu", d.get("cpu_baseline")); print("  roof", {k:v for k,v in d["roofline"].items() if k not in ("note","peak_source")})
        for x in d.get("extra_configs", []):
            print("  extra", json.dumps(x)[:1500])
    except Exception as e:
        print(f, "FAILED", e)
PY
