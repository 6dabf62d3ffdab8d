#!/bin/bash
# One gpurun call: smoke -> GPU parity tests -> bench -> ncu launch list + full capture of the fused kernel.
# Everything lands in gpurun_out/.  Each stage has its own timeout so a hung kernel cannot eat the lease.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; tail -c 3000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peaks --no-also > gpurun_out/bench_under_ncu.log 2>&1
tail -n 12 gpurun_out/launches.csv
echo "== ncu full capture of the fused kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 3 -c 1 -f -o gpurun_out/prof_fused \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-peaks --no-also > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
fi
