#!/bin/bash
# one GPU visit: parity tests, then A/B of the rho+grad variants, then the perf matrix
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_ab_env.sh "1 2 3" "X=0" "OKB_PLAIN_CHUNKS=1" "OKB_VARIANT=MB11xBN1xWM2xWN4xNPW8" "OKB_VARIANT=MB11xBN2xWM2xWN2xNPW8" 2>&1 | tee gpurun_out/ab_grad.txt
bash scripts/gpu_ab_env.sh "4 5 6" "X=0" "OKB_PLAIN_CHUNKS=1" 2>&1 | tee gpurun_out/ab_lap.txt
bash scripts/gpu_ab_env.sh "" "X=0" "OKB_PLAIN_CHUNKS=1" 2>&1 | tee gpurun_out/ab_val.txt
timeout 900 python scripts/perf_matrix.py 2>&1 | tee gpurun_out/perf_matrix.txt
