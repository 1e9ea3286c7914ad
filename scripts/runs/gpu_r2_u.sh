#!/bin/bash
# round 2: full GPU suite after position shards + aggregate consumer
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/u_pytest.log
tail -n 25 gpurun_out/u_pytest.log
