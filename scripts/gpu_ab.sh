#!/bin/bash
# A/B of instrumented / alternative builds of the table-level kernel in ONE box session (boxes differ in clocks / power cap)
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  for i in 1 2; do MINER_B200_LIB=miner_b200/$lib timeout 200 python scripts/debug_tscore.py speed 2>&1 | tail -1; done
done | tee gpurun_out/ab.txt
