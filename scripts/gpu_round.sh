#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/perf_ci.txt
import sys; sys.path.insert(0, 'scripts'); sys.argv = ['x']
import os; os.environ['PERF_BIG'] = '0'
import importlib.util
spec = importlib.util.spec_from_file_location('pm', 'scripts/perf_matrix.py'); pm = importlib.util.module_from_spec(spec); spec.loader.exec_module(pm)
pm.run_ci('c5 500 MOs', 12, 10, 1000, 128)
PY
echo "== gather"; OKB_CI_GATHER=1 python - <<'PY' 2>&1 | tee gpurun_out/perf_ci_gather.txt
import sys, os
import importlib.util
spec = importlib.util.spec_from_file_location('pm', 'scripts/perf_matrix.py'); pm = importlib.util.module_from_spec(spec); spec.loader.exec_module(pm)
pm.run_ci('c5 500 MOs', 12, 10, 1000, 128)
PY
