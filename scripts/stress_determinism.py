"""Repeat the device step (table_project -> tpack + scoring -> rank_metrics) with idle gaps in between and compare every output with
the first run bit for bit: a timing-dependent race shows up as a mismatch.  Usage: python scripts/stress_determinism.py [repeats]"""
import sys, time, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
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
bad_total = 0
for n_impr in (20000, 3000, 65536):
    eb = synth.make_eval_batch(n_impr, H, N, 36, mean_cands=20.0)
    d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    tws = ops.score_table_workspace(n_impr, H, K, dev)
    ref = None
    for it in range(reps):
        if it % 4 == 1:
            time.sleep(0.4)                      # let the clocks drop: the next launches start cold
        proj = ops.table_project(table, sw, workspace=proj_ws)
        scores = torch.full((int(eb.offsets[-1]),), float('nan'), dtype=torch.float32, device=dev)   # unwritten scores would stay NaN
        ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'], out_scores=scores, workspace=tws)
        part, _ = ops.rank_metrics_raw(scores, d['labels'], d['offsets'], 'sigmoid', (5, 10))
        cur = {'lg': proj.lg.clone(), 'tw': proj.tw.clone(), 'scores': scores, 'partials': part.clone()}
        torch.cuda.synchronize()
        if ref is None:
            ref = cur
            print(n_impr, 'nan scores in first run:', int(torch.isnan(scores).sum()))
            continue
        for k in cur:
            a, b = cur[k], ref[k]
            same = torch.equal(a, b) or bool(((a == b) | (torch.isnan(a.float()) & torch.isnan(b.float()))).all())
            if not same:
                bad_total += 1
                neq = (a != b)
                idx = neq.reshape(-1).nonzero().reshape(-1)
                print('MISMATCH', n_impr, 'iteration', it, k, 'elements', int(neq.sum()), 'first idx', idx[:5].tolist(), 'last idx', idx[-3:].tolist())
    print(n_impr, 'done')
print('mismatching outputs:', bad_total)
