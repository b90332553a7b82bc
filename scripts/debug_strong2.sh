#!/bin/bash
# on a fresh 2-GPU box: the strong-scaling bench at 1 rank (three times, the first one is the first CUDA process of the box) and at 2 ranks
C="bench.py --scaling strong --impressions 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-order --no-breakdown --no-extras"
for i in 1 2 3; do python $C 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('one', d['metrics'])"; done
for i in 1 2; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$i $C --gpus 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('two', d['metrics'])"; done
CUDA_VISIBLE_DEVICES=1 python $C 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('one on gpu1', d['metrics'])"
