#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/perf_ci_occ.txt
for sm in 0 110000 70000 50000 36000 24000; do echo "== OKB_CI_SMEM=$sm" | tee -a gpurun_out/perf_ci_occ.txt; OKB_CI_SMEM=$sm OKB_CI_GATHER=1 python scripts/perf_ci2.py 2>&1 | tee -a gpurun_out/perf_ci_occ.txt; done
