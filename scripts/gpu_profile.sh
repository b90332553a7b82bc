#!/bin/bash
# ncu evidence for the bench command (run via gpurun, 1 GPU): launch list with device times, then a full-set capture of
# the heaviest kernels.  The plain run goes first and must exit 0.
mkdir -p gpurun_out
CMD="python bench.py --impressions ${PROF_IMPR:-65536} --steps 2 --warmup 3 --no-cpu-baseline --no-reference-order --no-extras"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${PROF_KERNELS:-tscore_x_kernel|tscore_kernel|tpack_kernel|rank_metrics_kernel|tc_gemm_kernel|table_logits_kernel}" -s ${PROF_SKIP:-12} -c ${PROF_COUNT:-5} -o gpurun_out/prof -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -8
