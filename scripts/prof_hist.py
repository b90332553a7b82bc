"""Cycle accounting of hist_kernel2 (needs the -DMINER_HIST_PROF build):
    python -m miner_b200.build --force -DMINER_HIST_PROF --out=libminer_b200_prof.so
    MINER_B200_LIB=miner_b200/libminer_b200_prof.so python scripts/prof_hist.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from miner_b200 import ops, synth, _lib

dev = 'cuda:0'
B, H, N, D, K, Dc = 4096, 50, 100000, 768, 32, 200
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, Dc, 36)
eb = synth.make_eval_batch(B, H, N, 36)
lib = _lib.load()
prof = torch.zeros(148 * 4 * 16, dtype=torch.int64, device=dev)
lib.miner_debug_set_hist_prof.argtypes = [C.c_void_p]
lib.miner_debug_set_hist_prof(prof.data_ptr())
wp16 = w.w_proj.to(torch.bfloat16).to(dev)
his = eb.his_ids.to(dev)
if '--same-ids' in sys.argv:          # ablation: every history row is table row 1 (gathers hit L2 / one DRAM page)
    his = torch.ones_like(his)
if '--few-ids' in sys.argv:           # ablation: ids from a 2000-row window (L2-resident)
    his = his % 2000
args = (table, his, eb.his_mask.to(dev), wp16, w.context_codes.to(dev))
for _ in range(3):
    ops.hist_interests(*args, want_f32=False)
torch.cuda.synchronize()
p = prof.cpu().view(148, 4, 16).double()
tiles = (B / 2) / 148
names = {0: ['other/issue-gap', 'wait E tile', 'issue P1', 'wait proj free (t_ready)', '-', '-', '-', '-', '-', 'wait Wp tile'],
         3: ['other/issue-gap', '-', '-', 'wait t_ready', 'issue LG', 'wait w_ready', 'wait ia_free', 'wait E tile', 'issue P2'],
         1: ['gap', 'wait p1_full', 'E1a tanh->T', 'wait lg_full', 'E1b softmax', 'wait ia_full', 'drain'],
         2: ['setup ids', 'wait empty', 'issue cp.async']}
for role, rn in ((0, 'MMA warp, projection pipeline'), (3, 'MMA warp, logits + weighted-sum pipeline'), (1, 'epilogue thread'), (2, 'gather thread (projection ring)')):
    tot = p[:, role, 15].mean()
    print(f'{rn}: total {tot:.0f} cycles = {tot / tiles:.0f} per tile')
    for i, n in enumerate(names[role]):
        if n == '-':
            continue
        v = p[:, role, i].mean()
        print(f'    {n:18s} {v / tiles:9.0f} cycles/tile  {100 * v / tot:5.1f}%')
