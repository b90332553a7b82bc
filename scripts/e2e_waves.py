"""End-to-end time of HostEvaluator.evaluate (pinned host inputs -> metric partials on the host) for several wave schedules, and the
pinned H2D copy rate of the same buffers.  Usage: python scripts/e2e_waves.py [impressions]"""
import sys, torch
sys.path.insert(0, '.')
import miner_b200 as mb
from miner_b200 import synth

n_impr = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device('cuda:0')
H, K, DC, D, N = 50, 32, 200, 768, 100000
table = synth.make_table(N, D, 36, torch.bfloat16).to(dev)
w = synth.make_weights(D, K, DC, 36)
model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
with torch.no_grad():
    model.poly_attn.linear.weight.copy_(w.w_proj)
    model.poly_attn.context_codes.copy_(w.context_codes)
    model.target_aware_attn.linear.weight.copy_(w.w_target)
eb = synth.make_eval_batch(n_impr, H, N, 36, mean_cands=20.0)
host = {k: getattr(eb, k).pin_memory() for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
nbytes = sum(t.numel() * t.element_size() for t in host.values())


def timed(fn, n=4):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


bufs = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
ms = timed(lambda: [bufs[k].copy_(host[k], non_blocking=True) for k in host])
print('H2D of one step (%.0f MB, 5 pinned tensors): %.2f ms = %.1f GB/s' % (nbytes / 1e6, ms, nbytes / ms / 1e6))
ref = None
for name, kw in [('65536 x 16 (round default)', dict(wave=65536)),
                 ('32768 fixed', dict(wave=32768)),
                 ('131072 fixed', dict(wave=131072)),
                 ('16384 then x2 up to 131072', dict(wave=65536, first_wave=16384, wave_growth=2.0, max_wave=131072)),
                 ('16384 then x2 up to 262144', dict(wave=65536, first_wave=16384, wave_growth=2.0, max_wave=262144)),
                 ('16384 then x1.5 up to 262144', dict(wave=65536, first_wave=16384, wave_growth=1.5, max_wave=262144)),
                 ('8192 then x2 up to 524288', dict(wave=65536, first_wave=8192, wave_growth=2.0, max_wave=524288)),
                 ('65536 x 16 again', dict(wave=65536))]:
    ev = mb.HostEvaluator(model, chunk=32768, ks=(5, 10), transform='sigmoid', **kw)
    out = []
    ms = timed(lambda: out.append(ev.evaluate(host)[0].cpu()))
    if ref is None:
        ref = out[-1]
    print('%-32s %2d waves  %.3f ms = %.2f M impressions/s   max |partials - first schedule| %.3g' % (
        name, len(ev._wave_bounds(n_impr)) - 1, ms, n_impr / ms / 1e3, float((out[-1] - ref).abs().max())))
    del ev
    torch.cuda.empty_cache()
