import sys, numpy as np, torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth
from oracle import miner_oracle as O
dev = 'cuda:0'
for B in (20000, 10000, 10016, 3001):
    eb = synth.make_eval_batch(B, 50, 100000, 36, mean_cands=20.0)
    T = int(eb.offsets[-1])
    g = torch.Generator().manual_seed(1)
    s = (torch.randn(T, generator=g) * 0.35)
    p = torch.sigmoid(s).double().numpy()
    ref = O.per_impression_metrics(eb.labels.numpy(), p, eb.offsets.numpy(), ks=(5, 10))
    part, per = ops.rank_metrics_raw(s.to(dev), eb.labels.to(dev), eb.offsets.to(dev), 'sigmoid', (5, 10), per_impression=True)
    per = per.cpu().numpy(); part = part.cpu().numpy().reshape(-1, 2)
    names = ops.metric_names((5, 10))
    lens = (eb.offsets[1:] - eb.offsets[:-1]).numpy()
    for i, n in enumerate(names):
        bad = ~np.isclose(per[:, i], ref[n], rtol=1e-9, atol=0, equal_nan=True)
        agg = part[i, 0] / part[i, 1]
        print(B, n, 'bad rows', int(bad.sum()), 'agg %.12f ref %.12f per-sum %.12f cnt %d/%d' % (agg, np.nanmean(ref[n]), np.nansum(per[:, i]) / np.sum(~np.isnan(per[:, i])), part[i, 1], np.sum(~np.isnan(ref[n]))))
        if bad.any():
            idx = np.nonzero(bad)[0][:8]
            print('   rows', idx, 'group', idx // 32, 'lens', lens[idx], 'got', per[idx, i], 'ref', ref[n][idx])
            grp = idx[0] // 32
            print('   group range', int(eb.offsets[min(grp * 32 + 32, B)] - eb.offsets[grp * 32]))
