"""Table-level scoring kernel against the oracle on a few shapes (run on the GPU box)."""
import sys, time
import torch
sys.path.insert(0, '.')
from miner_b200 import ops, synth
from oracle import miner_oracle as O
DEV = 'cuda:0'

def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

def run(B, H, N, D, K, Dc, mean_c, max_c, st='weighted', bias=False):
    table = synth.make_table(N, D, 5, torch.bfloat16)
    w = synth.make_weights(D, K, Dc, 5)
    eb = synth.make_eval_batch(B, H, N, 7, mean_cands=mean_c, max_cands=max_c)
    sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
    tp = ops.table_project(table.to(DEV), sw)
    torch.cuda.synchronize()
    # projections against fp32 torch
    wp = w.w_proj.to(torch.bfloat16).float(); wt = w.w_target.to(torch.bfloat16).float()
    lg_ref = torch.tanh(table.float() @ wp.T) @ w.context_codes.T
    tw_ref = table.float() @ wt.T
    print('  lg err', nerr(tp.lg.cpu(), lg_ref), 'tw err', nerr(tp.tw.float().cpu(), tw_ref))
    bm = torch.randn(B, H, generator=torch.Generator().manual_seed(1)) * 0.3 if bias else None
    I, s = ops.score_table(tp, eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV), st, cand_offsets=eb.offsets.to(DEV),
                           bias_mean=None if bm is None else bm.to(DEV), want_interests=True)
    torch.cuda.synchronize()
    offs = eb.offsets.numpy()
    Iref = O.poly_attention(table.float()[eb.his_ids], eb.his_mask, wp, w.context_codes, None if bm is None else bm[:, :, None])
    ref = torch.empty(int(offs[-1]))
    for i in range(B):
        cr = table.float()[eb.cand_ids[offs[i]:offs[i + 1]]][None]
        ref[offs[i]:offs[i + 1]] = O.aggregate_scores(Iref[i:i + 1], cr, st, wt)[0]
    print(f'B={B} H={H} D={D} K={K} Dc={Dc} C~{mean_c} {st} bias={bias}: interests err {nerr(I.cpu(), Iref):.2e}  scores err {nerr(s.cpu(), ref):.2e}', flush=True)

if __name__ == '__main__':
    if 'multi' in sys.argv:
        run(2, 12, 300, 64, 8, 24, 150.0, 300)
        sys.exit(0)
    if 'speed' not in sys.argv:
      run(2, 12, 300, 64, 8, 24, 5.0, 10)
      run(6, 12, 300, 64, 8, 24, 20.0, 300)
      run(37, 50, 900, 768, 32, 200, 20.0, 300)
      run(301, 50, 900, 256, 32, 48, 12.0, 70, bias=True)
      run(33, 64, 900, 128, 16, 40, 20.0, 300, st='max')
      run(33, 33, 900, 128, 16, 40, 20.0, 300, st='mean')
      run(9, 50, 900, 768, 32, 200, 150.0, 300)
    # speed
    B, H, N, D, K, Dc = 200000, 50, 100000, 768, 32, 200
    table = synth.make_table(N, D, 5, torch.bfloat16).to(DEV)
    w = synth.make_weights(D, K, Dc, 5)
    eb = synth.make_eval_batch(B, H, N, 7, fixed_cands=20)
    sw = ops.ScoreWeights(w.w_proj.to(DEV), w.context_codes.to(DEV), w.w_target.to(DEV), True)
    args = (eb.his_ids.to(DEV), eb.his_mask.to(DEV), eb.cand_ids.to(DEV))
    offs = eb.offsets.to(DEV)
    for it in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        tp = ops.table_project(table, sw)
        e1.record()
        ops.score_table(tp, *args, 'weighted', cand_offsets=offs)
        e2.record()
        torch.cuda.synchronize()
        print(f'project {e0.elapsed_time(e1):.2f} ms  score {e1.elapsed_time(e2):.2f} ms  -> {B / e1.elapsed_time(e2) / 1e3:.2f} M impressions/s (kernel), '
              f'{B / e0.elapsed_time(e2) / 1e3:.2f} M/s incl. projections', flush=True)
