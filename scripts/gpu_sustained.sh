#!/bin/bash
# A/B of builds under sustained load in ONE box session; the first build runs again at the end (thermal drift check)
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  MINER_B200_LIB=miner_b200/$lib timeout 200 python scripts/sustained_ab.py ${CALLS:-40} 2>&1 | tail -1
done | tee gpurun_out/sustained_ab.txt
