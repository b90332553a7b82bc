"""CPU emulation of the table-level kernel's roundings (tw bf16, softmax weights hi/lo, I hi/lo, G single / hi+lo, gelu form)
against the fp32 oracle, at the synthetic weight scale and at 'trained-like' scales where P = I Wt^T has std ~1.
Answers ADVICE r1 (medium): does the tanh-form gelu / single-bf16 G hold the 1e-3 score tolerance off the near-linear regime?
    python scripts/numerics_table_mode.py
"""
import math, sys, torch
sys.path.insert(0, '.')
from miner_b200 import synth
from oracle import miner_oracle as O

torch.manual_seed(0)
def bf(x): return x.to(torch.bfloat16).to(torch.float32)
def hilo(x):
    h = bf(x); return h, bf(x - h)
def gelu_tanh(x): return 0.5 * x * (1 + torch.tanh(x * (0.7978845608 + 0.0356774081 * x * x)))

def run(scale_t, scale_w, B=256, H=50, K=32, Dc=200, D=768, C=20, N=5000):
    table = bf(synth.make_table(N, D) * scale_t)
    w = synth.make_weights(D, K, Dc)
    wp, codes, wt = bf(w.w_proj), w.context_codes, bf(w.w_target * scale_w)
    b = synth.make_eval_batch(B, H, N, fixed_cands=C)
    E = table[b.his_ids]; cand = table[b.cand_ids.view(B, C)]
    ref_i = O.poly_attention(E, b.his_mask, wp, codes)
    ref = O.aggregate_scores(ref_i, cand, 'weighted', wt)
    # table-level emulation
    lg = torch.tanh(table @ wp.T) @ codes.T
    tw = bf(table @ wt.T)
    l = lg[b.his_ids].permute(0, 2, 1).masked_fill(~b.his_mask[:, None, :], 1e-30)
    wgt = torch.softmax(l, dim=2)
    wh, wl = hilo(wgt)
    I = wh @ E + wl @ E
    P = wh @ tw[b.his_ids] + wl @ tw[b.his_ids]
    ih, il = hilo(I)
    m = cand @ ih.permute(0, 2, 1) + cand @ il.permute(0, 2, 1)
    out = {}
    for name, g in (('tanh,G bf16', bf(gelu_tanh(P))), ('erf,G bf16', bf(torch.nn.functional.gelu(P))),
                    ('tanh,G hi+lo', sum(hilo(gelu_tanh(P)))), ('erf,G hi+lo', sum(hilo(torch.nn.functional.gelu(P))))):
        a = cand @ g.permute(0, 2, 1)
        s = (torch.softmax(a, dim=2) * m).sum(2)
        out[name] = ((s - ref).abs().max() / ref.abs().max()).item()
    # tw in fp32 (no bf16 rounding of the projected table) to isolate its share
    P32 = wgt @ (table @ wt.T)[b.his_ids]
    a = cand @ torch.nn.functional.gelu(P32).permute(0, 2, 1)
    out['erf, tw fp32, G fp32'] = (((torch.softmax(a, dim=2) * m).sum(2) - ref).abs().max() / ref.abs().max()).item()
    flips = (torch.argsort(s, dim=1) != torch.argsort(ref, dim=1)).any(1).sum().item()
    Pstd = P.std().item()
    return Pstd, P.abs().max().item(), ref.abs().max().item(), out, flips

for st, sw in ((1, 1), (1, 20), (1, 60), (3, 20), (1, 150)):
    pstd, pmax, rmax, out, flips = run(st, sw)
    print(f'table x{st} Wt x{sw}: P std {pstd:.3f} max {pmax:.2f}  max|score| {rmax:.2f}')
    for k, v in out.items(): print(f'    {k:24s} normwise err {v:.2e}')
