#!/bin/bash
# quick GPU iteration: parity tests, then the device-resident bench leg only
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench (device leg)" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_quick.json'))
    r = d['roofline']
    print('value %.3e pts/s  ms/step %.2f  frac %.4f (peak %.2f)  kernel %s  e2e %.3e' % (d['value'], d['ms_per_step'], r['frac'], r['peak'], r['kernel'], d['e2e']['value']))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
