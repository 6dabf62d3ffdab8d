#!/bin/bash
# ncu --set full of the fused kernel on the Config-2 shape with 21 occupied orbitals (generation-bound narrow tile)
set -u
mkdir -p gpurun_out
PROF_CASE=c2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 2 -c 1 -f -o gpurun_out/r02_c2_rho \
    python scripts/prof_rho.py > gpurun_out/ncu_c2_rho.log 2>&1
tail -1 gpurun_out/ncu_c2_rho.log | cut -c1-200
PROF_CASE=c2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:okb_ws_kernel -s 2 -c 1 -f -o gpurun_out/r02_c2_grad \
    python scripts/prof_rho.py 1 2 3 > gpurun_out/ncu_c2_grad.log 2>&1
tail -1 gpurun_out/ncu_c2_grad.log | cut -c1-200
ls -la gpurun_out/r02_c2_*.ncu-rep
