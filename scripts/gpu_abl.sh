#!/bin/bash
# A/B of ablation / alternative builds of the X kernel in ONE box session: scripts/gpu_abl.sh lib1.so lib2.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  MINER_B200_LIB=miner_b200/$lib timeout 200 python scripts/debug_tsx.py speed 2>&1 | grep "x kernel"
done | tee gpurun_out/abl.txt
