#!/bin/bash
# cycle accounting + gather ablations of tscore_kernel (instrumented library), then the production kernel's speed
mkdir -p gpurun_out
export PYTHONPATH=.
( MINER_B200_LIB=miner_b200/libminer_b200_prof.so timeout 200 python scripts/prof_tscore.py
  echo "== fixed 20 candidates"; MINER_B200_LIB=miner_b200/libminer_b200_prof.so timeout 200 python scripts/prof_tscore.py --fixed | head -3
  for d in 1 2 4 6; do echo "== MINER_TS_DBG=$d"; MINER_TS_DBG=$d MINER_B200_LIB=miner_b200/libminer_b200_prof.so timeout 200 python scripts/prof_tscore.py | head -1; done
  echo "== production build"; timeout 200 python scripts/debug_tscore.py speed ) > gpurun_out/ts_cycles.txt 2>&1
cat gpurun_out/ts_cycles.txt
