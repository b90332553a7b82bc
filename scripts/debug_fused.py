"""GPU debug driver for the two fused tensor-core kernels: each checked alone against the CPU oracle (test infrastructure)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from miner_b200 import ops, synth, _lib
from oracle import miner_oracle as O

dev = 'cuda:0'
which = sys.argv[1] if len(sys.argv) > 1 else 'all'


def problem(B, H, N, D, K, Dc, seed=3, mean_c=20.0, max_c=300):
    table = synth.make_table(N, D, seed, torch.bfloat16)
    w = synth.make_weights(D, K, Dc, seed)
    eb = synth.make_eval_batch(B, H, N, seed, mean_cands=mean_c, max_cands=max_c)
    return table, w, eb


def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_hist(B, H, N, D, K, Dc):
    table, w, eb = problem(B, H, N, D, K, Dc)
    wp16 = w.w_proj.to(torch.bfloat16)
    E = table.float()[eb.his_ids]
    ref = O.poly_attention(E, eb.his_mask, wp16.float(), w.context_codes)
    t0 = time.time()
    ihi, ilo, out = ops.hist_interests(table.to(dev), eb.his_ids.to(dev), eb.his_mask.to(dev), wp16.to(dev), w.context_codes.to(dev))
    torch.cuda.synchronize()
    e_f32 = nerr(out.cpu(), ref)
    e_split = nerr((ihi.float() + ilo.float()).cpu().view(B, K, D), ref)
    print(f'hist B={B} H={H} D={D} K={K} Dc={Dc}: f32 err {e_f32:.2e}  hi+lo err {e_split:.2e}  ({time.time()-t0:.2f}s)', flush=True)
    return e_f32 < 3e-4 and e_split < 3e-4


def run_cand(B, H, N, D, K, Dc, mean_c=20.0, max_c=300):
    table, w, eb = problem(B, H, N, D, K, Dc, mean_c=mean_c, max_c=max_c)
    E = table.float()[eb.his_ids]
    I = O.poly_attention(E, eb.his_mask, w.w_proj, w.context_codes)          # fp32 interests (CPU)
    ihi = I.to(torch.bfloat16)
    ilo = (I - ihi.float()).to(torch.bfloat16)
    wt16 = w.w_target.to(torch.bfloat16)
    # reference: same bf16-valued Wt, fp32 everything else
    offs = eb.offsets.numpy()
    ref = torch.empty(int(offs[-1]))
    for i in range(B):
        cr = table.float()[eb.cand_ids[offs[i]:offs[i + 1]]][None]
        ref[offs[i]:offs[i + 1]] = O.aggregate_scores(I[i:i + 1], cr, 'weighted', wt16.float())[0]
    t0 = time.time()
    s = ops.cand_score(ihi.view(B * K, D).to(dev), ilo.view(B * K, D).to(dev), wt16.to(dev), table.to(dev), eb.cand_ids.to(dev), K,
                       cand_offsets=eb.offsets.to(dev))
    torch.cuda.synchronize()
    e = nerr(s.cpu(), ref)
    print(f'cand B={B} D={D} K={K} T={int(offs[-1])} maxC={int((eb.offsets[1:]-eb.offsets[:-1]).max())}: score err {e:.2e}  ({time.time()-t0:.2f}s)', flush=True)
    return e < 3e-4


ok = True
if which in ('hist', 'all'):
    for cfg in [(6, 12, 64, 64, 8, 24), (37, 50, 500, 768, 32, 200), (300, 50, 5000, 256, 32, 48), (33, 100, 500, 128, 16, 40), (5, 64, 100, 64, 32, 16)]:
        ok &= run_hist(*cfg)
if which in ('cand', 'all'):
    for cfg in [(6, 12, 64, 64, 8, 24), (37, 50, 500, 768, 32, 200), (300, 50, 5000, 256, 32, 48), (33, 100, 500, 128, 16, 40)]:
        ok &= run_cand(*cfg)
    ok &= run_cand(9, 50, 500, 768, 32, 200, mean_c=150.0, max_c=300)      # groups with several 128-candidate passes
print('ALL OK' if ok else 'FAILURES')
sys.exit(0 if ok else 1)
