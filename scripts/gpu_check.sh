#!/bin/bash
# Run on the B200 box via gpurun: parity tests in two groups (CUDA-core kernels first, then the tcgen05 ones under their
# own timeout so a hang cannot hide the other results), then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
TC='tc_gemm or tensor or test_score_impressions_csr or properties_at_scale or table or hist_kernel or cand_kernel or fused or sweep or host_evaluator or fastformer'
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "not ($TC)" > gpurun_out/t1.log 2>&1
echo "t1 exit $?" >> gpurun_out/t1.log
tail -5 gpurun_out/t1.log
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 90 --timeout-method=thread -k "$TC" > gpurun_out/t2.log 2>&1
echo "t2 exit $?" >> gpurun_out/t2.log
tail -15 gpurun_out/t2.log
timeout -k 10 600 python bench.py --impressions ${BENCH_IMPR:-200000} --steps 3 --warmup 3 --cpu-sample 1000 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
