#!/bin/bash
# round 2, first GPU pass: all GPU tests, racecheck / synccheck triage material, one bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== racecheck on minimal, correctly synchronised hand-overs (scripts/racecheck_repro.cu)"
for c in 0 1 2 3; do
  echo "-- case $c"
  timeout 300 compute-sanitizer --tool racecheck --racecheck-report all ./scratch/racecheck_repro $c 2>&1 | tail -25
done | tee gpurun_out/racecheck_repro.txt
T="tests/test_gpu_parity.py::test_laplacian_phi_cache_subranges tests/test_gpu_parity.py::test_reference_golden_refdata"
echo "== racecheck on the fused kernel (full log)"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python -m pytest $T -x -q -m gpu > gpurun_out/racecheck_full.txt 2>&1
tail -5 gpurun_out/racecheck_full.txt
echo "== synccheck"
timeout 1200 compute-sanitizer --tool synccheck python -m pytest $T tests/test_gpu_ci.py::test_calc_mo_matrix_and_calc_jmo -x -q -m gpu > gpurun_out/synccheck.txt 2>&1
tail -5 gpurun_out/synccheck.txt
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
