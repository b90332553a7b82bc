#!/usr/bin/env python
"""BASELINE.json configs[3] at N GPUs (weak scaling: every rank scores its own B impressions per step, metrics partials all-reduced):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sweep_multi.py [B]
One JSON line per (H, K): whole-job impressions/s from the max-over-ranks device time of 3 steps."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib, parallel

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = f'cuda:{local}'
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(dev))
B, N, D, DC = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 100000, 768, 200
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
for H in (50, 100, 200):
    eb = synth.make_eval_batch(B, H, N, 36 + 1000 * rank)
    d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    for K in (8, 16, 32, 64):
        w = synth.make_weights(D, K, DC, 36)
        m = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
        with torch.no_grad():
            m.poly_attn.linear.weight.copy_(w.w_proj); m.poly_attn.context_codes.copy_(w.context_codes); m.target_aware_attn.linear.weight.copy_(w.w_target)

        def step():
            m._table_proj = None                                     # projections recomputed inside every step
            s = m.score_impressions(d['his_ids'], d['his_mask'], d['cand_ids'], d['offsets'])
            part = ops.rank_metrics_raw(s, d['labels'], d['offsets'], 'sigmoid', (5, 10))[0]
            if world > 1:
                dist.all_reduce(part)                                # [sum, count] partials of the six metrics
            return part
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t)
            print(json.dumps({'n_gpus': world, 'H': H, 'K': K, 'impressions_per_s': world * B / ms * 1e3, 'ms_per_step': ms,
                              'impressions_per_gpu_per_step': B}), flush=True)
if world > 1:
    dist.destroy_process_group()
