#!/bin/bash
# compute-sanitizer memcheck + synccheck over the round-2 kernels: remainder-orbital tiles (forced), interleaved AO tiles,
# time-dependent contraction, overlap integrals, z-run calc_ao
set -u
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_remainder_orbital_tiles_forced_on_small_molecules tests/test_gpu_ci.py::test_time_dependent_contractions tests/test_gpu_overlap.py tests/test_gpu_parity.py::test_reference_golden_refdata"
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | tail -8
done | tee gpurun_out/sanitize_r2.txt
