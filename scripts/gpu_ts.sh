#!/bin/bash
# quick loop for the table-level kernel on the B200 box: parity on a few shapes + kernel speed, then the table tests
mkdir -p gpurun_out
timeout -k 5 300 python scripts/debug_tscore.py > gpurun_out/dbg_ts.log 2>&1; echo "debug exit $?" >> gpurun_out/dbg_ts.log
tail -15 gpurun_out/dbg_ts.log
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread -k "table" > gpurun_out/t_table.log 2>&1
echo "table tests exit $?" >> gpurun_out/t_table.log
tail -12 gpurun_out/t_table.log
