#!/bin/bash
# build libminer_b200 with tscore_kernel.cu taken from a git revision (A/B runs on the GPU box): scripts/build_rev.sh <rev> <out.so>
set -e
cp miner_b200/csrc/tc/tscore_kernel.cu /tmp/ts_cur_$$.cu
git show "$1":miner_b200/csrc/tc/tscore_kernel.cu > miner_b200/csrc/tc/tscore_kernel.cu
python -m miner_b200.build --force --out="$2" | tail -1 || true
cp /tmp/ts_cur_$$.cu miner_b200/csrc/tc/tscore_kernel.cu
touch miner_b200/csrc/tc/tscore_kernel.cu
python -m miner_b200.build | tail -1
