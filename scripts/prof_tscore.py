"""Cycle accounting of tscore_kernel (needs the -DMINER_TS_PROF build):
    python -m miner_b200.build --force -DMINER_TS_PROF --out=libminer_b200_prof.so
    MINER_B200_LIB=miner_b200/libminer_b200_prof.so python scripts/prof_tscore.py [--same-ids] [--few-ids]
Prints, per role, the mean over CTAs of each counter as a share of the kernel's cycles."""
import ctypes as C, sys
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth, _lib
DEV = 'cuda:0'
B, H, N, D, K, Dc = 200000, 50, 100000, 768, 32, 200
table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
w = synth.make_weights(D, K, Dc, 5)
eb = synth.make_eval_batch(B, H, N, 7, fixed_cands=20 if '--fixed' in sys.argv else None)
WHICH = int(sys.argv[sys.argv.index('--epi') + 1]) if '--epi' in sys.argv else 0
sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
tp = ops.table_project(table, sw)
lib = _lib.load()
prof = torch.zeros(148 * 5 * 16, dtype=torch.int64, device=DEV)
lib.miner_debug_set_hist_prof.argtypes = [C.c_void_p]
lib.miner_debug_set_hist_prof(prof.data_ptr())
his, cand = eb.his_ids.to(DEV), eb.cand_ids.to(DEV)
if '--same-ids' in sys.argv:
    his, cand = torch.ones_like(his), torch.ones_like(cand)
if '--few-ids' in sys.argv:
    his, cand = his % 2000, cand % 2000
args = (tp, his, eb.his_mask.to(DEV), cand, 'weighted')
offs = eb.offsets.to(DEV)
for _ in range(3):
    ops.score_table(*args, cand_offsets=offs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.score_table(*args, cand_offsets=offs); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'{ms:.2f} ms  {B / ms / 1e3:.2f} M impressions/s', sys.argv[1:])
p = prof.cpu().view(148, 5, 16).double()
names = {0: ('MMA issuer', ['other', 'wait w_ready', 'wait full1 (E,TW)', 'issue S1', 'wait a_ready', 'wait full2 (cand)', 'wait dma_free', 'issue S2']),
         1: ('epilogue (interest half)', ['other', 'wait ip_full', 'fence+arrive', 'wait dma_full', 'scores', 'tmem ld+wait', 'shuffle/gelu/pack', 'tmem st issue', 'tmem st wait']),
         2: ('gather (E, TW ring)', ['other', 'wait empty1', 'issue E,TW']),
         4: ('epilogue (gelu half)', ['other', 'wait ip_full', 'fence+arrive', 'wait dma_full', 'scores', 'tmem ld+wait', 'shuffle/gelu/pack', 'tmem st issue', 'tmem st wait']),
         3: ('softmax + scores', ['other', 'load lg', 'softmax', 'wait w_free', 'store', 'wait dma_full', 'score stage'])}
for role, (rn, cn) in names.items():
    tot = p[:, role, 15].mean()
    print(f'{rn}: total {tot:.0f} cycles')
    for i, n in enumerate(cn):
        print(f'    {n:24s} {p[:, role, i].mean() / tot * 100:5.1f} %')

