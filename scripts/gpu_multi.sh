#!/bin/bash
# 2-GPU validation: strong-scaling rank invariance (torchrun test), multi-rank global auc, bench at N=2 (weak + strong extra)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -k "strong_scaling or rank_count" > gpurun_out/t_multi.log 2>&1; echo "multi tests exit $?" >> gpurun_out/t_multi.log; tail -5 gpurun_out/t_multi.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 scripts/check_auc_multi.py > gpurun_out/auc_multi.log 2>&1; echo "auc exit $?" >> gpurun_out/auc_multi.log; tail -4 gpurun_out/auc_multi.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -2 gpurun_out/bench_n2.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, 'e2e', d['e2e']['value'])
print('strong', d.get('strong_scaling'))
print('train', {k: d['train_step'].get(k) for k in ('value', 'ms_per_step', 'error')})
PY
