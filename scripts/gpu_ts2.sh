#!/bin/bash
bash scripts/gpu_ts.sh
MINER_B200_LIB=miner_b200/libminer_b200_prof.so timeout 200 python scripts/prof_tscore.py > gpurun_out/ts_cycles.txt 2>&1
cat gpurun_out/ts_cycles.txt
