#!/bin/bash
# round 2: K-sweep pacing probe at biobank scale + full GPU suite
mkdir -p gpurun_out
( time timeout 400 python scripts/pace_probe.py 16000 ) > gpurun_out/s_pace_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/s_pace_probe.log
cat gpurun_out/s_pace_probe.log
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -n 15 gpurun_out/s_pytest.log
