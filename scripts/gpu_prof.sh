#!/bin/bash
# tests + one ncu --set full capture of the fused kernel (bench device leg, 1 step)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench (device leg)" ; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json
d = json.load(open('gpurun_out/bench_quick.json')); r = d['roofline']
print('value %.3e pts/s  ms/step %.2f  frac %.4f (peak %.2f)  kernel %s  e2e %.3e' % (d['value'], d['ms_per_step'], r['frac'], r['peak'], r['kernel'], d['e2e']['value']))"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 3 -c 1 -f -o gpurun_out/${PROF_NAME:-prof_ws} \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-peaks --no-also > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
