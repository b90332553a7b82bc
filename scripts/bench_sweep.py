#!/usr/bin/env python
"""BASELINE.json configs[3]: eval scoring sweep over history length and number of context codes on one GPU
(scoring + ranking metrics per step, inputs resident, 768-d bf16 table of 100k news, ~20 candidates per impression).
Prints one line per (H, K) with the kernel family `score_impressions` picked for that shape."""
import sys, json
import torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import ops, synth, _lib
dev = 'cuda:0'
B, N, D, DC = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 100000, 768, 200
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
names = {_lib.MATH_TABLE: 'table-level (table_project + tscore_kernel)', _lib.MATH_TENSOR: 'reference order, tcgen05 (hist_kernel2 + cand_kernel) or tc_gemm pipeline',
         _lib.MATH_FP32: 'reference order, fp32'}
rows = []
for H in (50, 100, 200):
    eb = synth.make_eval_batch(B, H, N, 36)
    d = {k: getattr(eb, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    for K in (8, 16, 32, 64):
        w = synth.make_weights(D, K, DC, 36)
        m = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
        with torch.no_grad():
            m.poly_attn.linear.weight.copy_(w.w_proj); m.poly_attn.context_codes.copy_(w.context_codes); m.target_aware_attn.linear.weight.copy_(w.w_target)
        math = ops.default_eval_math(table, H, K)

        def step():
            m._table_proj = None                                     # projections recomputed inside every step
            s = m.score_impressions(d['his_ids'], d['his_mask'], d['cand_ids'], d['offsets'])
            return ops.rank_metrics_raw(s, d['labels'], d['offsets'], 'sigmoid', (5, 10))[0]
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        rows.append({'H': H, 'K': K, 'impressions_per_s': B / ms * 1e3, 'ms_per_step': ms, 'kernels': names[math]})
        print(json.dumps(rows[-1]), flush=True)
