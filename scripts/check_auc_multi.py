"""torchrun --nproc-per-node N: the multi-rank global auc (all-gather of the positive keys + int64 all-reduce) equals the
single-rank value and the numpy integer formula, bit for bit."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from miner_b200 import ops, synth, parallel
from miner_b200.evaluation import global_auc
rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
B = 200_000
eb = synth.make_eval_batch(B, 50, 100_000, 3)
g = torch.Generator().manual_seed(1)
scores = torch.round(torch.randn(int(eb.offsets[-1]), generator=g) * 64) / 64      # quantised: plenty of ties
s, e = parallel.shard_bounds(eb.offsets, world, 50)[rank]
c0, c1 = int(eb.offsets[s]), int(eb.offsets[e])
auc_multi = global_auc(scores[c0:c1].to(dev), eb.labels[c0:c1].to(dev), None, 'sigmoid')
if rank == 0:
    pos, neg = ops.auc_split(scores.to(dev), eb.labels.to(dev), None, 'sigmoid')
    u2 = ops.auc_count(ops.sort_u32(pos.clone()), neg)
    auc_single = u2 / (2.0 * pos.numel() * neg.numel())
    p = (1.0 / (1.0 + torch.exp(-scores.to(dev)))).cpu().numpy(); y = eb.labels.numpy()
    ps = np.sort(p[y > 0]); n = p[y <= 0]
    lb, ub = np.searchsorted(ps, n, 'left'), np.searchsorted(ps, n, 'right')
    u2_np = int((2 * (len(ps) - ub) + (ub - lb)).sum())
    print(f'world {world}: auc multi {auc_multi!r} single {auc_single!r} numpy {u2_np / (2.0 * len(ps) * len(n))!r}')
    assert auc_multi == auc_single == u2_np / (2.0 * len(ps) * len(n))
    print('OK')
dist.destroy_process_group()
