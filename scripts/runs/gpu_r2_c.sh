#!/bin/bash
# round 2, GPU call C: converged-warp mainloop + division-free Fisher: parity suite, ablations, C1 / C2 bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
timeout 600 python scripts/ceiling2.py > gpurun_out/c_ceiling2.log 2>&1
for occ in 2 3; do
  TWKB_STATS_OCC=$occ timeout 300 python bench.py --variants 10000 --min-r2 0 --steps 5 --no-cpu-baseline --no-mma-ceiling > gpurun_out/c_bench_c1_occ$occ.json 2> gpurun_out/c_bench_c1_occ$occ.err
done
timeout 600 python bench.py > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err
tail -4 gpurun_out/c_pytest.log; cat gpurun_out/c_ceiling2.log
python - <<'PY'
import json
for f in ("c_bench_c1_occ2","c_bench_c1_occ3","c_bench_c2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.4g ms %.2f count %.2f stats %.2f e2e %.4g (%.1f ms) frac %.3f ceil %s cpu %s" % (d["value"], d["ms_per_step"], d["config"]["ms_count_kernel_per_step"], d["config"]["ms_stats_kernel_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("mma_only_ceiling"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "FAILED", e)
PY
