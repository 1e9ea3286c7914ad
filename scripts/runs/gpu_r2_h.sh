#!/bin/bash
# round 2, 8-GPU call: bench at N = 8 / 4 / 2 (weak scaling of configs[1] + configs[3] strong scaling), multi-GPU tests
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
bash scripts/runs/gpu_r2_g.sh 8
bash scripts/runs/gpu_r2_g.sh 4
bash scripts/runs/gpu_r2_g.sh 2
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/h_pytest_mg.log 2>&1
tail -3 gpurun_out/h_pytest_mg.log
