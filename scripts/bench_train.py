#!/usr/bin/env python
"""Train-step timing of the MINER path (BASELINE.json configs[2]: npratio 4, batch 4096 per GPU, history 50, K=32, Dc=200, D=768).

    python scripts/bench_train.py [--batch 4096] [--steps 10] [--warmup 40]
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_train.py --gpus N

One step = Miner.forward (train variant) + Loss.compute + backward (CUDA kernels of libminer_b200.so) + one flat NCCL
all-reduce of the gradients + SGD update of the three weight matrices.  Prints one JSON line (rank 0).  The table is a
frozen bf16 buffer of synthetic news vectors; inputs are resident on the device (a second line times the same step with the
batch copied from pinned host memory each step).
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=40, help='untimed steps; with several ranks the first ~20 steps (NCCL channel setup, allocator growth) run at half speed')
    ap.add_argument('--table', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--math', default='tensor', choices=['tensor', 'fp32'], help='tensor: projection-sized GEMMs on tcgen05 (bf16 operands)')
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    import torch.distributed as dist
    import miner_b200 as mb
    from miner_b200 import ops, synth, parallel
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    H, K, DC, D, N, NP = 50, 32, 200, 768, 100_000, 4
    table = synth.make_table(N, D, 36, torch.bfloat16 if args.table == 'bf16' else torch.float32).to(dev)
    w = synth.make_weights(D, K, DC, 36)
    model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).train()
    model.train_math = args.math
    with torch.no_grad():
        model.poly_attn.linear.weight.copy_(w.w_proj)
        model.poly_attn.context_codes.copy_(w.context_codes)
        model.target_aware_attn.linear.weight.copy_(w.w_target)
    his, mask, _, cand, _, labels = synth.make_train_batch(args.batch, H, N, NP, 36 + rank)
    host = {'his': his.pin_memory(), 'mask': mask.pin_memory(), 'cand': cand.pin_memory(), 'labels': labels.float().pin_memory()}
    res = {k: v.to(dev) for k, v in host.items()}
    loss_fn = mb.Loss(nn.CrossEntropyLoss(reduction='mean'))
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    B, C = cand.shape
    z = torch.zeros(B, C, 1, dtype=torch.long, device=dev)
    zh = torch.zeros(B, H, 1, dtype=torch.long, device=dev)

    def step(d):
        opt.zero_grad(set_to_none=True)
        I, S = model(d['cand'][..., None], z, d['his'][..., None], zh, d['mask'], z, z, zh, zh)
        loss = loss_fn.compute(I, S, d['labels'])
        loss.backward()
        parallel.allreduce_gradients(list(model.parameters()))
        opt.step()
        return loss

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out

    for _ in range(max(args.warmup, 3)):
        step(res)
    l0 = ops.launch_count()
    ms, loss = timed(lambda: step(res), args.steps)
    launches = (ops.launch_count() - l0) // args.steps
    ms_e2e, _ = timed(lambda: step({k: v.to(dev, non_blocking=True) for k, v in host.items()}), args.steps)
    if rank == 0:
        flops = 3 * (2 * H * D * DC + 2 * H * DC * K + 2 * K * H * D + 2 * K * D * D + 4 * C * K * D)      # fwd + ~2x in the backward
        print(json.dumps({'metric': 'train samples/sec', 'value': B * world / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world,
                          'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
                          'dtype': ('bf16 operands / f32 accumulate (tcgen05) for the five projection-sized GEMMs, f32 elsewhere' if args.math == 'tensor'
                                    else 'f32 kernels (reference operation order)') + ', %s table' % args.table, 'data': 'synthetic',
                          'config': {'workload': 'MINER train step, npratio 4, history 50, K=32, Dc=200, D=768, frozen 100k-news table',
                                     'batch_per_gpu': B, 'parallelism': f'dp{world}, one flat gradient all-reduce'},
                          'e2e': {'value': B * world / (ms_e2e * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_e2e,
                                  'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host.values()), 'd2h_bytes_per_step': 0},
                          'gpu_launches': int(launches), 'loss': float(loss),
                          'achieved_tflops': flops * B / (ms * 1e-3) / 1e12}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
