"""Does any kernel of the device step read memory it did not write?  Before every run the caching allocator's free blocks are filled
with a byte pattern (so every torch.empty of the run returns that pattern), then the whole step runs on freshly allocated buffers;
all outputs must equal the first run's bit for bit.  Usage: python scripts/stress_poison.py"""
import sys, time, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib

import os
ASYNC = os.environ.get('STRESS_ASYNC') == '1'      # no synchronisation between the steps of a run (stream races show up)
if os.environ.get('STRESS_NO_WAIT') == '1':          # what the HostEvaluator did before the copy stream waited for the compute stream
    torch.cuda.Stream.wait_stream = lambda self, other: None
dev = torch.device('cuda:0')
H, K, DC, D, N = 50, 32, 200, 768, 100000
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, DC, 36)
model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
with torch.no_grad():
    model.poly_attn.linear.weight.copy_(w.w_proj)
    model.poly_attn.context_codes.copy_(w.context_codes)
    model.target_aware_attn.linear.weight.copy_(w.w_target)


def poison(byte):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((256 << 20,), byte, dtype=torch.uint8, device=dev) for _ in range(6)]     # 1.5 GB of pattern
    small = [torch.full((1 << 16,), byte, dtype=torch.uint8, device=dev) for _ in range(64)]
    torch.cuda.synchronize()
    del blocks, small                                   # back to the allocator's cache: the next torch.empty calls get these bytes


bad = 0
for n_impr in (20000, 777, 65536):
    eb = synth.make_eval_batch(n_impr, H, N, 36, mean_cands=20.0)
    d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    host = {k: getattr(eb, k).pin_memory() for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    ref = None
    for it, byte in enumerate([0x00, 0xFF, 0x7F, 0x80, 0x3C, 0xFF, 0x00]):
        def step(name, fn):
            out = fn()
            if ASYNC:
                return out
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print('FAULT after', name, 'size', n_impr, 'pattern 0x%02X' % byte, '->', str(e).splitlines()[0], flush=True)
                raise SystemExit(1)
            return out
        step('poison', lambda: poison(byte))
        model.invalidate()
        sw = step('weights', lambda: model._weights(with_bf16=True))
        proj = step('table_project', lambda: ops.table_project(table, sw))
        _, scores = step('score_table', lambda: ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets']))
        part, per = step('rank_metrics', lambda: ops.rank_metrics_raw(scores, d['labels'], d['offsets'], 'sigmoid', (5, 10), per_impression=True))
        ev = mb.HostEvaluator(model, wave=4096, chunk=1024)
        p_e2e, s_e2e = step('evaluate', lambda: ev.evaluate(host, want_scores=True))
        i_hi = step('reference order', lambda: model.score_impressions(d['his_ids'], d['his_mask'], d['cand_ids'], d['offsets'], math=_lib.MATH_TENSOR)) if n_impr <= 20000 else scores
        print('ok', n_impr, 'pattern 0x%02X' % byte, flush=True)
        cur = {'lg': proj.lg, 'tw': proj.tw, 'scores': scores, 'partials': part, 'per': per, 'e2e_scores': s_e2e, 'e2e_partials': p_e2e, 'ref_order': i_hi}
        torch.cuda.synchronize()
        if ref is None:
            ref = {k: v.clone() for k, v in cur.items()}
            continue
        for k in cur:
            a, b = cur[k], ref[k]
            same = bool(((a == b) | (torch.isnan(a.float()) & torch.isnan(b.float()))).all())
            if not same:
                bad += 1
                neq = ~((a == b) | (torch.isnan(a.float()) & torch.isnan(b.float())))
                idx = neq.reshape(-1).nonzero().reshape(-1)
                print('MISMATCH', n_impr, 'pattern 0x%02X' % byte, k, 'elements', int(neq.sum()), 'of', a.numel(), 'first', idx[:5].tolist(), 'last', idx[-3:].tolist())
    print(n_impr, 'done')
print('mismatching outputs:', bad)
