#!/bin/bash
# round 2: full GPU suite + default bench (N = 1)
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -15 gpurun_out/q_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/q_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/q_bench_ref.json 2> gpurun_out/q_bench_ref.err
echo "ref rc=$?"; tail -3 gpurun_out/q_bench_ref.err; cat gpurun_out/q_bench_ref.json | cut -c1-600
