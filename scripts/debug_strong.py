import sys, json, subprocess, numpy as np, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, parallel, _lib
common = ['bench.py', '--scaling', 'strong', '--impressions', '20000', '--steps', '1', '--warmup', '3', '--no-cpu-baseline',
          '--no-reference-order', '--no-breakdown', '--no-extras']
for i in range(2):
    r = subprocess.run([sys.executable] + common, capture_output=True, text=True)
    print('bench N=1 run', i, json.loads(r.stdout.strip().splitlines()[-1])['metrics'])
dev = torch.device('cuda:0')
H, K, DC, D, N = 50, 32, 200, 768, 100000
seed = 36
table = synth.make_table(N, D, seed, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, DC, seed)
model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
with torch.no_grad():
    model.poly_attn.linear.weight.copy_(w.w_proj)
    model.poly_attn.context_codes.copy_(w.context_codes)
    model.target_aware_attn.linear.weight.copy_(w.w_target)
eb = synth.make_eval_batch(20000, H, N, seed, mean_cands=20.0)
names = ops.metric_names((5, 10))
sw = model._weights(with_bf16=True)
proj = ops.table_project(table, sw)
res = {}
for ws in (1, 2, 4):
    total, scs = None, []
    for s, e in parallel.shard_bounds(eb.offsets, ws, H):
        c0, c1 = int(eb.offsets[s]), int(eb.offsets[e])
        offs = (eb.offsets[s:e + 1] - c0).to(dev)
        _, sc = ops.score_table(proj, eb.his_ids[s:e].to(dev), eb.his_mask[s:e].to(dev), eb.cand_ids[c0:c1].to(dev), 'weighted', cand_offsets=offs)
        p, _ = ops.rank_metrics_raw(sc, eb.labels[c0:c1].to(dev), offs, 'sigmoid', (5, 10))
        total = p if total is None else total + p
        scs.append(sc)
    res[ws] = torch.cat(scs)
    print('in-process shards', ws, parallel.finalize_metrics(total, names), 'scores equal to ws=1:', bool(torch.equal(res[ws], res[1])),
          'max diff', float((res[ws] - res[1]).abs().max()))
# repeat ws=1 several times: determinism of the scoring kernel at this size
for i in range(3):
    _, sc = ops.score_table(proj, eb.his_ids.to(dev), eb.his_mask.to(dev), eb.cand_ids.to(dev), 'weighted', cand_offsets=eb.offsets.to(dev))
    print('repeat', i, 'equal', bool(torch.equal(sc, res[1])), float((sc - res[1]).abs().max()))
proj2 = ops.table_project(table, sw)
print('table_project deterministic: lg', bool(torch.equal(proj.lg, proj2.lg)), 'tw', bool(torch.equal(proj.tw, proj2.tw)))
