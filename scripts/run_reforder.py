"""A few launches of the reference-order tensor family (hist_kernel2 + cand_kernel) on synthetic impressions (for ncu captures)."""
import sys
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth, _lib
DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
H, N, D, K, Dc = 50, 100000, 768, 32, 200
table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
w = synth.make_weights(D, K, Dc, 5)
eb = synth.make_eval_batch(B, H, N, 7)
sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
args = (eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV))
offs = eb.offsets.to(DEV)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.score(table, *args, sw, 'weighted', cand_offsets=offs, math=_lib.MATH_TENSOR)
    e1.record()
    torch.cuda.synchronize()
    print(f'{e0.elapsed_time(e1):.2f} ms -> {B / e0.elapsed_time(e1) / 1e3:.2f} M impressions/s')
