#!/bin/bash
# ncu full-set capture of one tscore_x_kernel launch (run via gpurun, 1 GPU); the plain run goes first and must exit 0.
mkdir -p gpurun_out
CMD="python scripts/run_tscore.py ${PROF_IMPR:-65536}"
$CMD > gpurun_out/tsx_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tscore_x_kernel -s 2 -c 1 -o gpurun_out/prof_tsx -f $CMD > gpurun_out/ncu_tsx.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_tsx.log
