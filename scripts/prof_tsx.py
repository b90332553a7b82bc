"""Cycle accounting of tscore_x_kernel (needs the -DMINER_TS_PROF build):
    python -m miner_b200.build --force -DMINER_TS_PROF --out=libminer_b200_prof.so
    MINER_B200_LIB=miner_b200/libminer_b200_prof.so python scripts/prof_tsx.py [--fixed]"""
import ctypes as C, sys
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth, _lib
DEV = 'cuda:0'
B, H, N, D, K, Dc = 200000, 50, 100000, 768, 32, 200
table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
w = synth.make_weights(D, K, Dc, 5)
eb = synth.make_eval_batch(B, H, N, 7, fixed_cands=20 if '--fixed' in sys.argv else None)
sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
tp = ops.table_project(table, sw)
lib = _lib.load()
prof = torch.zeros(148 * 5 * 16, dtype=torch.int64, device=DEV)
lib.miner_debug_set_hist_prof.argtypes = [C.c_void_p]
lib.miner_debug_set_hist_prof(prof.data_ptr())
args = (tp, eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV), 'weighted')
offs = eb.offsets.to(DEV)
for _ in range(3):
    ops.score_table(*args, cand_offsets=offs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.score_table(*args, cand_offsets=offs); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'{ms:.2f} ms  {B / ms / 1e3:.2f} M impressions/s', sys.argv[1:])
p = prof.cpu().view(148, 5, 16).double()
names = {0: ('MMA issuer', ['other', 'wait w_ready', 'wait full1 (E,TW)', 'issue S1', 'wait a_ready', 'wait full2 (cand)', 'wait dma_free', 'issue S2', 'wait x_free', 'issue SX']),
         1: ('epilogue', ['other', 'wait ip_full', 'fence+arrive', '-', '-', 'tmem ld+wait', 'gelu/split', '-', 'tmem st+wait']),
         2: ('gather (E, TW ring)', ['other', 'wait empty1', 'issue E,TW']),
         3: ('softmax', ['other', 'wait lg rows + barrier + prefetch issue', 'softmax', 'wait w_free', 'store + arrive']),
         4: ('score warps', ['other', 'wait x_full', 'barrier', 'X -> smem (+barriers)', '-', 'wait dma_full', 'scores', 'A_w -> W', 'm = W X', '-', 'drain D_a', 'barrier after drain'])}
for role, (rn, cn) in names.items():
    tot = p[:, role, 15].mean()
    print(f'{rn}: total {tot:.0f} cycles')
    for i, n in enumerate(cn):
        if n != '-':
            print(f'    {n:40s} {p[:, role, i].mean() / tot * 100:5.1f} %')
