import sys, torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth
dev='cuda:0'
eb = synth.make_eval_batch(1000000, 50, 100000, 36)
s = (torch.randn(int(eb.offsets[-1])) * 0.35).to(dev); y = eb.labels.to(dev); o = eb.offsets.to(dev)
for _ in range(3): ops.rank_metrics_raw(s, y, o)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.rank_metrics_raw(s, y, o)
e1.record(); torch.cuda.synchronize()
print('rank_metrics 1M impressions: %.3f ms' % (e0.elapsed_time(e1)/5))
