#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_ab_env.sh "1 2 3" "X=0" 2>&1 | tee gpurun_out/ab_grad.txt
