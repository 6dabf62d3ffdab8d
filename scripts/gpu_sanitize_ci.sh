#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck over the detCI kernels added late in round 2 (split, sequential
# shared-memory and dense orbital-pair kernels) and the narrow-tile variants of the fused kernel
set -u
mkdir -p gpurun_out
T="tests/test_gpu_ci.py::test_fast_sums_dense_and_split tests/test_gpu_ci.py::test_random_terms_bitwise_vs_oracle tests/test_gpu_ci.py::test_h3p_reference_test_reproduced tests/test_gpu_parity.py::test_reference_golden_refdata"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | tail -8
done | tee gpurun_out/sanitize_ci.txt
