#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the small GPU tests that exercise every kernel family
set -u
mkdir -p gpurun_out
T="tests/test_gpu_out.py tests/test_gpu_ci.py::test_calc_mo_matrix_and_calc_jmo tests/test_gpu_parity.py::test_laplacian_phi_cache_subranges tests/test_gpu_parity.py::test_reference_golden_refdata"
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  OKB_AO_VARIANT="" timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | tail -6
done | tee gpurun_out/sanitize.txt
echo "== memcheck on the warp-specialised AO kernel"
OKB_AO_VARIANT=aows timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_parity.py::test_fixture_molecules_vs_reference_outputs" -x -q -m gpu -k "synth_small_sph or h2o_gaussian_sph" 2>&1 | tail -4 | tee -a gpurun_out/sanitize.txt
