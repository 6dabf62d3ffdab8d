#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_ab_env.sh "4 5 6" "X=0" "OKB_LAP_ONE_PASS=1" 2>&1 | tee gpurun_out/ab_lap.txt
AB_NMO=1000 AB_N=100 bash scripts/gpu_ab_env.sh "4 5 6" "X=0" "OKB_LAP_ONE_PASS=1" 2>&1 | tee -a gpurun_out/ab_lap.txt
PROBE_PEAKS=0 python scripts/e2e_probe2.py 2>&1 | tail -4 | tee gpurun_out/e2e_probe2.txt
