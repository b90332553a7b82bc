"""Where a device step's time goes when the steps run back to back (bench.py's `value` loop), stage by stage:
CUDA events between the stages INSIDE the loop, against each stage timed alone after an idle gap (bench.py's `kernels`).
Usage: python scripts/step_gap.py [impressions] [steps]"""
import sys, time, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib

n_impr = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda:0')
H, K, DC, D, N = 50, 32, 200, 768, 100000
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, DC, 36)
model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
with torch.no_grad():
    model.poly_attn.linear.weight.copy_(w.w_proj)
    model.poly_attn.context_codes.copy_(w.context_codes)
    model.target_aware_attn.linear.weight.copy_(w.w_target)
eb = synth.make_eval_batch(n_impr, H, N, 36, mean_cands=20.0)
d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
sw = model._weights(with_bf16=True)
proj = ops.table_project(table, sw)
proj_ws = torch.empty(max(_lib.load().miner_table_project_workspace_bytes(N, DC), 1), dtype=torch.uint8, device=dev)
tws = ops.score_table_workspace(n_impr, H, K, dev)
scores = torch.empty(int(eb.offsets[-1]), dtype=torch.float32, device=dev)

stages = [
    ('table_project', lambda: ops.table_project(table, sw, out=proj, workspace=proj_ws)),
    ('tpack+tscore', lambda: ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'],
                                             out_scores=scores, workspace=tws)),
    ('rank_metrics', lambda: ops.rank_metrics_raw(scores, d['labels'], d['offsets'], 'sigmoid', (5, 10))),
]


def step():
    for _, fn in stages:
        fn()


for _ in range(3):
    step()
torch.cuda.synchronize()
# (1) the loop as bench.py times it
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    step()
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print('loop: %.3f ms per step (host issued %d steps in %.2f ms)' % (e0.elapsed_time(e1) / steps, steps, t_issue * 1e3))
# (2) the same loop with an event after every stage
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(steps)]
for s in range(steps):
    ev[s][0].record()
    for i, (_, fn) in enumerate(stages):
        fn()
        ev[s][i + 1].record()
torch.cuda.synchronize()
for i, (name, _) in enumerate(stages):
    ts = [ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(steps)]
    print('  in loop  %-14s %s  mean %.3f' % (name, ' '.join('%.3f' % t for t in ts), sum(ts) / steps))
gaps = [ev[s][len(stages)].elapsed_time(ev[s + 1][0]) for s in range(steps - 1)]
print('  between steps: %s' % ' '.join('%.3f' % t for t in gaps))
print('  whole: %.3f ms per step' % (ev[0][0].elapsed_time(ev[-1][-1]) / steps))
# (3) each stage alone after a pause (bench.py's per-kernel figures)
for name, fn in stages:
    fn()
    torch.cuda.synchronize()
    time.sleep(0.2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    print('  alone    %-14s %.3f' % (name, a.elapsed_time(b)))
# (4) the scoring stage back to back without the others
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    stages[1][1]()
b.record()
torch.cuda.synchronize()
print('  tpack+tscore x %d back to back: %.3f ms each' % (steps, a.elapsed_time(b) / steps))
