#!/bin/bash
# ncu full-set capture of hist_kernel2 + cand_kernel (reference-order tensor family); the plain run goes first and must exit 0.
mkdir -p gpurun_out
CMD="python scripts/run_reforder.py ${PROF_IMPR:-32768}"
$CMD > gpurun_out/ro_plain.log 2>&1 && cat gpurun_out/ro_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:"hist_kernel2|cand_kernel" -s 2 -c 2 -o gpurun_out/prof_ro -f $CMD > gpurun_out/ncu_ro.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_ro.log
