"""X-formulation table-level kernel (scores only) against the oracle and against tscore_kernel (run on the GPU box)."""
import sys
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth
from oracle import miner_oracle as O
DEV = 'cuda:0'

def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

def run(B, H, N, D, K, Dc, mean_c, max_c, st='weighted', bias=False, seed=7):
    table = synth.make_table(N, D, 5, torch.bfloat16)
    w = synth.make_weights(D, K, Dc, 5)
    eb = synth.make_eval_batch(B, H, N, seed, mean_cands=mean_c, max_cands=max_c)
    sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
    tp = ops.table_project(table.to(DEV), sw)
    wp = w.w_proj.to(torch.bfloat16).float(); wt = w.w_target.to(torch.bfloat16).float()
    bm = torch.randn(B, H, generator=torch.Generator().manual_seed(1)) * 0.3 if bias else None
    a = (tp, eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV), st)
    kw = dict(cand_offsets=eb.offsets.to(DEV), bias_mean=None if bm is None else bm.to(DEV))
    _, s_old = ops.score_table(*a, want_interests=True, **kw)     # tscore_kernel (interests requested)
    _, s = ops.score_table(*a, **kw)                              # X kernel where the shape allows
    torch.cuda.synchronize()
    offs = eb.offsets.numpy()
    Iref = O.poly_attention(table.float()[eb.his_ids], eb.his_mask, wp, w.context_codes, None if bm is None else bm[:, :, None])
    ref = torch.empty(int(offs[-1]))
    for i in range(B):
        cr = table.float()[eb.cand_ids[offs[i]:offs[i + 1]]][None]
        ref[offs[i]:offs[i + 1]] = O.aggregate_scores(Iref[i:i + 1], cr, st, wt)[0]
    e, eo, ex = nerr(s.cpu(), ref), nerr(s_old.cpu(), ref), nerr(s.cpu(), s_old.cpu())
    bad = '' if e < 3e-4 else '   <<<<<< BAD'
    print(f'B={B} H={H} D={D} K={K} C~{mean_c} {st} bias={bias}: x-kernel err {e:.2e}  old {eo:.2e}  x vs old {ex:.2e}{bad}', flush=True)
    if bad:
        d = (s.cpu() - ref).abs() / ref.abs().max()
        idx = torch.nonzero(d > 3e-4).flatten()
        imp = torch.bucketize(idx, eb.offsets, right=True) - 1
        print('   bad candidates', idx[:12].tolist(), 'impressions', imp[:12].tolist(), 'of', B, 'nan', int(torch.isnan(s).sum()))

if __name__ == '__main__':
    if 'speed' not in sys.argv:
        run(2, 12, 300, 64, 8, 24, 5.0, 10)
        run(6, 12, 300, 64, 8, 24, 20.0, 300)
        run(37, 50, 900, 768, 32, 200, 20.0, 300)
        run(301, 50, 900, 256, 32, 48, 12.0, 70, bias=True)
        run(33, 56, 900, 128, 16, 40, 20.0, 300, st='max')
        run(33, 33, 900, 128, 16, 40, 20.0, 300, st='mean')
        run(9, 50, 900, 768, 32, 200, 150.0, 300)
        run(1001, 50, 5000, 768, 32, 200, 20.0, 300, seed=11)
        run(700, 50, 5000, 128, 30, 40, 30.0, 300, seed=12)
    B, H, N, D, K, Dc = 200000, 50, 100000, 768, 32, 200
    table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
    w = synth.make_weights(D, K, Dc, 5)
    sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
    for fixed in (20, None):
        eb = synth.make_eval_batch(B, H, N, 7, fixed_cands=fixed)
        args = (eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV))
        offs = eb.offsets.to(DEV)
        tp = ops.table_project(table, sw)
        for wi in (False, True):
            for it in range(3):
                e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                e1.record()
                ops.score_table(tp, *args, 'weighted', cand_offsets=offs, want_interests=wi)
                e2.record()
                torch.cuda.synchronize()
            print(f'cands {fixed}  {"tscore_kernel (interests out)" if wi else "x kernel"}: {e1.elapsed_time(e2):.2f} ms -> {B / e1.elapsed_time(e2) / 1e3:.2f} M impressions/s', flush=True)
