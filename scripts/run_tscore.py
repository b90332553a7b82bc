"""A few launches of the table-level scoring path on synthetic MIND-shaped impressions (for ncu captures)."""
import sys
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth
DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H, N, D, K, Dc = 50, 100000, 768, 32, 200
table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
w = synth.make_weights(D, K, Dc, 5)
eb = synth.make_eval_batch(B, H, N, 7)
sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
args = (eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV))
offs = eb.offsets.to(DEV)
for it in range(4):
    tp = ops.table_project(table, sw)
    ops.score_table(tp, *args, 'weighted', cand_offsets=offs)
torch.cuda.synchronize()
print('ok', B)
