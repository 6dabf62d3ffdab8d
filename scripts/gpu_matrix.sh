#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python scripts/perf_matrix.py 2>&1 | tee gpurun_out/perf_matrix.txt | grep -v "n_mo=1000\|e2e pieces"
