#!/bin/bash
# ncu --set full captures (one launch each): fused rho+grad kernel, time-dependent contraction, launch list of the bench
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 2 -c 1 -f -o gpurun_out/r02_grad_rem \
    python scripts/prof_rho.py 1 2 3 > gpurun_out/ncu_grad.log 2>&1
tail -2 gpurun_out/ncu_grad.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okb_td_kernel -c 1 -f -o gpurun_out/r02_td \
    python scripts/perf_td.py one > gpurun_out/ncu_td.log 2>&1
tail -2 gpurun_out/ncu_td.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peaks --no-also --no-strong --no-latency > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/r02_launches_bench.csv | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
