#!/bin/bash
# N-GPU evidence in one box session (gpurun --gpus N): strong-scaling rank-invariance tests, multi-rank global auc, the H x K sweep and
# the bench (weak + strong extra + train step) at N ranks.   scripts/gpu_n8.sh <N>
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_n$N.txt
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 600 -k "strong_scaling or rank_count" > gpurun_out/t_multi_n$N.log 2>&1; echo "multi tests exit $?" >> gpurun_out/t_multi_n$N.log; tail -3 gpurun_out/t_multi_n$N.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 scripts/check_auc_multi.py > gpurun_out/auc_multi_n$N.log 2>&1; echo "auc exit $?" >> gpurun_out/auc_multi_n$N.log; tail -3 gpurun_out/auc_multi_n$N.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29673 scripts/bench_sweep_multi.py > gpurun_out/sweep_n$N.jsonl 2> gpurun_out/sweep_n$N.err; echo "sweep exit $?"; cut -c1-90 gpurun_out/sweep_n$N.jsonl
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"; tail -2 gpurun_out/bench_n$N.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'])
print('strong', d.get('strong_scaling'))
print('train', {k: d['train_step'].get(k) for k in ('value', 'ms_per_step', 'error')})
PY
