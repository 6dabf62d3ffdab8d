#!/bin/bash
# A/B of kernel configurations selected by environment variables: tests once, then the device bench leg per config
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for cfg in "$@"; do
  echo "== bench [$cfg]"
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python -c "
import json
try:
    d = json.load(open('gpurun_out/bench_ab.json')); r = d['roofline']
    print('value %.3e pts/s  ms/step %.2f  frac %.4f (peak %.2f)  kernel %s' % (d['value'], d['ms_per_step'], r['frac'], r['peak'], r['kernel']))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_ab.err').read()[-2000:])
"
done
