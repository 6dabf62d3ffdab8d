#!/bin/bash
# A/B of the calc_ao kernel variants (OKB_AO_VARIANT substrings given as arguments; "" = default); EXTRA_ENV="A=1 B=2"
mkdir -p gpurun_out
for v in "$@"; do
  echo "== OKB_AO_VARIANT=$v ${EXTRA_ENV:-}"
  env ${EXTRA_ENV:-} OKB_AO_VARIANT="$v" timeout 300 python scripts/perf_ao.py 2>&1 | tail -6
done | tee -a gpurun_out/ab_ao.txt
