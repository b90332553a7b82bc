"""Thousands of back-to-back device steps on one small batch, every output compared with the first run's on the device (no host
synchronisation inside the loop): how often does a step differ?  Usage: python scripts/stress_repeat.py [reps]"""
import sys, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
dev = torch.device('cuda:0')
H, K, DC, D, N = 50, 32, 200, 768, 100000
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, DC, 36)
model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
with torch.no_grad():
    model.poly_attn.linear.weight.copy_(w.w_proj)
    model.poly_attn.context_codes.copy_(w.context_codes)
    model.target_aware_attn.linear.weight.copy_(w.w_target)
sw = model._weights(with_bf16=True)
proj_ws = torch.empty(max(_lib.load().miner_table_project_workspace_bytes(N, DC), 1), dtype=torch.uint8, device=dev)
for n_impr in (20000, 10000, 2000):
    eb = synth.make_eval_batch(n_impr, H, N, 36, mean_cands=20.0)
    d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    tws = ops.score_table_workspace(n_impr, H, K, dev)
    proj = ops.table_project(table, sw, workspace=proj_ws)
    scores = torch.empty(int(eb.offsets[-1]), dtype=torch.float32, device=dev)
    ref = None
    bad = torch.zeros(4, dtype=torch.int64, device=dev)
    for it in range(reps):
        ops.table_project(table, sw, out=proj, workspace=proj_ws)
        ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'], out_scores=scores, workspace=tws)
        part, _ = ops.rank_metrics_raw(scores, d['labels'], d['offsets'], 'sigmoid', (5, 10))
        if ref is None:
            ref = (proj.lg.clone(), proj.tw.clone(), scores.clone(), part.clone())
            continue
        bad[0] += (proj.lg != ref[0]).any()
        bad[1] += (proj.tw != ref[1]).any()
        bad[2] += (scores != ref[2]).any()
        bad[3] += (part != ref[3]).any()
    print(n_impr, 'runs', reps - 1, 'differing [lg, tw, scores, partials]:', bad.tolist(), flush=True)
