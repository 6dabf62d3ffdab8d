#!/bin/bash
# parity tests + A/B of the fused kernel variants + short bench + time-dependent contraction timings
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== n_mo A/B"; bash scripts/gpu_ab_nmo.sh 80 82 2>&1 | tee gpurun_out/ab_nmo.txt
echo "== forced MB11 for 82"; OKB_VARIANT=MB11x bash scripts/gpu_ab_nmo.sh 82 2>&1 | tee -a gpurun_out/ab_nmo.txt
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-latency > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['also'])"
echo "== td"; python scripts/perf_td.py 2>&1 | tee gpurun_out/perf_td.txt
