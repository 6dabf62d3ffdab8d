#!/bin/bash
# A/B of the narrow MO-tile variants on the Config-2 shape (21 and 40 orbitals)
for nmo in 21 40; do
  for v in "" MB3xBN2 MB3xBN3 MB3xBN8 MB6xBN1 MB6xBN2 MB6xBN4 MB6xBN8; do
    if [ $nmo = 21 ] && [[ $v == MB6* ]]; then continue; fi
    if [ $nmo = 40 ] && [[ $v == MB3* ]]; then continue; fi
    echo "== NMO=$nmo OKB_VARIANT=$v"
    NMO=$nmo OKB_VARIANT=$v timeout 120 python scripts/perf_c2_occ.py 2>&1 | tail -3
  done
done
