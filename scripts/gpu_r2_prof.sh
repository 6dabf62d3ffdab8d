#!/bin/bash
# ncu --set full captures (one launch each) of the fused rho+grad kernel and of the z-run calc_ao kernel
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 2 -c 1 -f -o gpurun_out/r02_grad \
    python scripts/prof_rho.py 1 2 3 > gpurun_out/ncu_grad.log 2>&1
tail -2 gpurun_out/ncu_grad.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okb_ao_zrun -s 2 -c 1 -f -o gpurun_out/r02_zrun \
    python scripts/prof_ao.py > gpurun_out/ncu_zrun.log 2>&1
tail -2 gpurun_out/ncu_zrun.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
