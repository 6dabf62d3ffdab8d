#!/bin/bash
# compute-sanitizer over the pipelined dense contraction (okb_td2_kernel: staging warp + cp.async ring, barriers reached from two
# code paths) through its three users: time-dependent detCI sums, dense detCI form, cy_core.mocreator / mooverlapmatrix
set -u
mkdir -p gpurun_out
T="tests/test_gpu_ci.py::test_time_dependent_contractions tests/test_gpu_ci.py::test_fast_sums_dense_and_split tests/test_gpu_overlap.py tests/test_gpu_parity.py::test_cy_core_dropins_random_shells"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | tail -8
done | tee gpurun_out/sanitize_td2.txt
