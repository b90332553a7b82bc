"""Sustained speed of the table-level scoring stage (tpack + scoring kernel) under the board's power cap: N back-to-back calls on 1 M
impressions, time per window of calls, with power / SM clock sampled through NVML while they run.  The library is picked by
MINER_B200_LIB (A/B builds).  Usage: python scripts/sustained_ab.py [calls] [impressions]"""
import sys, time, threading, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth
import pynvml

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_impr = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
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
d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'offsets')}
proj = ops.table_project(table, model._weights(with_bf16=True))
tws = ops.score_table_workspace(n_impr, H, K, dev)
scores = torch.empty(int(eb.offsets[-1]), dtype=torch.float32, device=dev)
fn = lambda: ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'], out_scores=scores,
                             workspace=tws)
fn(); torch.cuda.synchronize()
time.sleep(1.0)                                   # start every variant from an idle board

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def pump():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.01)
t = threading.Thread(target=pump); t.start()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(calls + 1)]
ev[0].record()
for i in range(calls):
    fn()
    ev[i + 1].record()
torch.cuda.synchronize()
stop = True; t.join()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(calls)]
win = max(calls // 4, 1)
means = [sum(ts[i:i + win]) / len(ts[i:i + win]) for i in range(0, calls, win)]
half = samples[len(samples) // 2:]
pw = sorted(s[0] for s in samples); ck = sorted(s[1] for s in samples)
print('first %.2f  windows %s  last-half mean %.3f ms  | power median %.0f max %.0f W, sm clock median %d min %d MHz (%d samples); '
      'second half of the run: mean %.0f W, mean clock %.0f MHz' % (
    ts[0], ' '.join('%.2f' % m for m in means), sum(ts[calls // 2:]) / len(ts[calls // 2:]), pw[len(pw) // 2], pw[-1], ck[len(ck) // 2], ck[0],
    len(samples), sum(s[0] for s in half) / len(half), sum(s[1] for s in half) / len(half)))
