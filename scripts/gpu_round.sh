#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_ab_env.sh "1 2 3" "X=0" 2>&1 | tee gpurun_out/ab_grad.txt
bash scripts/gpu_ab_env.sh "4 5 6" "X=0" 2>&1 | tee gpurun_out/ab_lap.txt
bash scripts/gpu_ab_env.sh "" "X=0" 2>&1 | tee gpurun_out/ab_val.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json
d = json.load(open('gpurun_out/bench_quick.json')); r = d['roofline']
print('value %.3e pts/s  ms/step %.2f  frac %.4f (peak %.2f)  kernel %s  e2e %.3e' % (d['value'], d['ms_per_step'], r['frac'], r['peak'], r['kernel'], d['e2e']['value']))"
