#!/usr/bin/env python
"""bench.py -- the MINER data-parallel scoring path on B200 (BASELINE.json metric: impressions scored / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--impressions B] [--chunk C]

One "step" = one pass of the hot path (gather -> poly attention -> target-aware aggregation -> per-candidate score ->
per-impression AUC/MRR/nDCG/hit partials) over one batch of B synthetic MIND-shaped impressions PER GPU
(BASELINE.json configs[1]: 1 M impressions, history 50, K=32, Dc=200, 768-d news vectors, ~20 candidates each;
table of 100k news, bf16).  Weak scaling: every rank owns its own B impressions (independent units, no data-path
collective); the only exchange is one NCCL all-reduce of the [sum,count] metric partials per step.

Printed JSON (rank 0, one line):
  value        whole-job impressions/s with the step's inputs resident in HBM (device-timed, max over ranks)
  e2e          same metric through the public API with HOST (pinned) inputs: H2D of ids/mask/labels/offsets and D2H of
               the metric partials inside the timed region
  roofline     dominant kernel of the step: algorithmic FLOPs (or bytes) per launch / its CUDA-event duration, against
               MEASURED_PEAKS.json
  cpu_baseline the CPU restatement of the reference path (oracle/, torch CPU fp32 + numpy metrics) timed on this box's
               host cores on a bounded sample -- a reported baseline, not the target
--impl reference times that CPU path alone (the reference is pure Python/PyTorch and cannot travel to the GPU box;
the oracle port is pinned to it by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

H, K, DC, D, N_NEWS, MEAN_C = 50, 32, 200, 768, 100_000, 20.0
KS = (5, 10)
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        d['_source'] = 'measured'
        return d
    d = dict(FALLBACK_PEAKS)
    d['_source'] = 'fallback'
    return d


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active') and not v.lower().startswith('not'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(power)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(n_impr: int, steps: int, warmup: int, seed: int = 36):
    """Times the CPU restatement of the reference path (oracle) on `n_impr` impressions per step, all host threads."""
    from miner_b200 import synth
    from oracle import miner_oracle as O
    table = synth.make_table(N_NEWS, D, seed)                      # fp32, as the reference computes
    w = synth.make_weights(D, K, DC, seed)
    eb = synth.make_eval_batch(n_impr, H, N_NEWS, seed, mean_cands=MEAN_C)
    offs = eb.offsets.numpy()
    labels = eb.labels.numpy()

    def step():
        with torch.no_grad():
            s = O.miner_forward_csr(table, eb.his_ids, eb.his_mask, eb.cand_ids, offs, w.w_proj, w.context_codes, w.w_target,
                                    'weighted', chunk=512)
        probs = np.asarray(O.sigmoid_probs(s))
        targets = [labels[offs[i]:offs[i + 1]].tolist() for i in range(n_impr)]
        preds = [probs[offs[i]:offs[i + 1]].tolist() for i in range(n_impr)]
        return O.compute_scores(targets, preds, ['group_auc', 'mrr', 'ndcg@5', 'ndcg@10'])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_impr / dt, dt, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--impressions', type=int, default=1_000_000, help='impressions per GPU per step')
    ap.add_argument('--chunk', type=int, default=0, help='impressions per kernel wave (0 = library default)')
    ap.add_argument('--cpu-sample', type=int, default=4000, help='impressions in the CPU baseline sample')
    ap.add_argument('--math', default='table', choices=['table', 'tensor', 'fp32'],
                    help='table: projections applied once per table row inside every step + one fused scoring kernel (default); '
                         'tensor / fp32: reference operation order')
    ap.add_argument('--no-reference-order', action='store_true', help='skip the extra reference-order (tensor family) timing')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-breakdown', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cores = len(os.sched_getaffinity(0))

    if args.impl == 'reference':
        if rank != 0:
            return 0
        torch.set_num_threads(cores)
        n = args.cpu_sample
        v, dt, _ = cpu_reference_run(n, max(args.steps, 1), max(args.warmup, 1))
        line = {'impl': 'reference', 'metric': 'impressions scored/sec', 'value': v, 'unit': 'impressions/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'MINER eval scoring, history 50, K=32, Dc=200, D=768, ~20 candidates/impression, 100k-news table',
                           'impressions_per_step': n, 'device': 'host CPU'},
                'cpu_baseline': {'value': v, 'unit': 'impressions/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                 'sample': f'{n} impressions per step (bounded sample of the 1M-impression workload), oracle/miner_oracle.py '
                                           f'= torch CPU fp32 restatement of the reference Miner.forward + numpy ranking metrics'},
                'e2e': {'value': v, 'unit': 'impressions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------- our arm
    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    import miner_b200 as mb
    from miner_b200 import ops, synth, parallel, _lib

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback in the product path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.impressions
    seed = 36
    math = {'table': _lib.MATH_TABLE, 'tensor': _lib.MATH_TENSOR, 'fp32': _lib.MATH_FP32}[args.math]

    # replicated table + weights; this rank's own impressions (weak scaling)
    table = synth.make_table(N_NEWS, D, seed, torch.bfloat16).to(dev)
    w = synth.make_weights(D, K, DC, seed)
    model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
    with torch.no_grad():
        model.poly_attn.linear.weight.copy_(w.w_proj)
        model.poly_attn.context_codes.copy_(w.context_codes)
        model.target_aware_attn.linear.weight.copy_(w.w_target)
    eb = synth.make_eval_batch(B, H, N_NEWS, seed + 1000 * rank, mean_cands=MEAN_C)
    T = int(eb.offsets[-1])
    host = {k: getattr(eb, k).pin_memory() for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    names = ops.metric_names(KS)
    chunk = args.chunk if args.chunk > 0 else 32768
    sw = model._weights(with_bf16=(math != _lib.MATH_FP32))
    scores_buf = torch.empty(T, dtype=torch.float32, device=dev)

    proj = ops.table_project(table, sw) if math == _lib.MATH_TABLE else None          # buffers; recomputed inside every step
    proj_ws = torch.empty(max(_lib.load().miner_table_project_workspace_bytes(table.shape[0], DC), 1), dtype=torch.uint8, device=dev)

    tws = ops.score_table_workspace(B, H, K, dev) if math == _lib.MATH_TABLE else None   # packed tiles: rebuilt by every call

    def score_step(d, m):
        if m == _lib.MATH_TABLE:
            ops.table_project(table, sw, out=proj, workspace=proj_ws)       # part of the step: nothing is carried over between steps
            ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'], out_scores=scores_buf,
                            workspace=tws)
        else:
            ops.score(table, d['his_ids'], d['his_mask'], d['cand_ids'], sw, 'weighted', cand_offsets=d['offsets'], math=m,
                      chunk=chunk, out_scores=scores_buf)

    def device_step(d, m=None):
        score_step(d, math if m is None else m)
        partials, _ = ops.rank_metrics_raw(scores_buf, d['labels'], d['offsets'], 'sigmoid', KS)
        parallel.allreduce_partials(partials)
        return partials

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, out

    # ---- value: inputs resident in HBM
    resident = {k: v.to(dev) for k, v in host.items()}
    for _ in range(max(args.warmup, 3)):
        device_step(resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    ms_dev, partials = timed(lambda: device_step(resident), args.steps)
    launches = ops.launch_count() - l0
    metrics_out = parallel.finalize_metrics(partials, names)
    reference_order = None
    if math == _lib.MATH_TABLE and not args.no_reference_order:
        # the same step in the reference's operation order (tensor family: per-row projections on tcgen05), for comparison
        for _ in range(2):
            device_step(resident, _lib.MATH_TENSOR)
        ms_ro, p_ro = timed(lambda: device_step(resident, _lib.MATH_TENSOR), max(1, min(args.steps, 3)))
        m_ro = parallel.finalize_metrics(p_ro, names)
        reference_order = {'value': B * world / (ms_ro * 1e-3), 'unit': 'impressions/s', 'ms_per_step': ms_ro,
                           'kernels': 'hist_kernel2 + cand_kernel (tcgen05, per-row projections), rank_metrics',
                           'max_metric_abs_diff_vs_table_mode': max(abs(m_ro[k] - metrics_out[k]) for k in names)}

    # ---- e2e: public API with host buffers, H2D + D2H inside the timed region
    # public API: HostEvaluator copies waves of impressions H2D on a copy stream while the previous wave is scored
    evaluator = mb.HostEvaluator(model, wave=2 * chunk, chunk=chunk, ks=KS, transform='sigmoid', math=math)

    def e2e_step():
        p, _ = evaluator.evaluate(host)
        parallel.allreduce_partials(p)
        return p.cpu()

    for _ in range(2):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel breakdown of one step (CUDA events on the launching stream), rank 0 only
    pk = peaks()
    kernels, roofline = None, None
    if rank == 0 and not args.no_breakdown:
        if math == _lib.MATH_TABLE:
            stages = [('table_project: tc_gemm tanh(table Wp^T) + table_logits + tc_gemm table Wt^T (once per step, N rows)', 'proj'),
                      ('tscore_kernel: gather table/tw/cand rows + softmax_H + interests + gelu + matching/attention MMAs + softmax_K + score (tcgen05)', 'score')]
        elif math == _lib.MATH_TENSOR:
            # fused tcgen05 path: two kernels per wave of `chunk` impressions
            stages = [('hist_kernel: gather + tanh(E Wp^T) + logits/softmax + weighted sum (tcgen05)', 1),
                      ('cand_kernel: gelu(I Wt^T) + matching/attention MMAs + softmax_K + score (tcgen05)', 8)]
        else:
            stages = [('sgemm: tanh(table[his] Wp^T) [gather fused]', 1), ('poly_softmax_wsum', 2), ('sgemm: gelu(I Wt^T)', 4),
                      ('target_score [cand gather fused]', 8)]
        nchunks = 1 if math == _lib.MATH_TABLE else (B + chunk - 1) // chunk
        kernels = []
        for name, mask in stages:
            if mask == 'proj':
                fn = lambda: ops.table_project(table, sw, out=proj, workspace=proj_ws)
            elif mask == 'score':
                fn = lambda: ops.score_table(proj, resident['his_ids'], resident['his_mask'], resident['cand_ids'], 'weighted',
                                             cand_offsets=resident['offsets'], out_scores=scores_buf, workspace=tws)
            else:
                fn = lambda: ops.score(table, resident['his_ids'], resident['his_mask'], resident['cand_ids'], sw, 'weighted',
                                       cand_offsets=resident['offsets'], math=math, chunk=chunk, out_scores=scores_buf, stage_mask=mask)
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            kernels.append({'kernel': name, 'ms_per_step': e0.elapsed_time(e1), 'launches_per_step': 3 if mask == 'proj' else nchunks})
        fn = lambda: ops.rank_metrics_raw(scores_buf, resident['labels'], resident['offsets'], 'sigmoid', KS)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        kernels.append({'kernel': 'rank_metrics (+finalize)', 'ms_per_step': e0.elapsed_time(e1), 'launches_per_step': 2})
        tot = sum(k['ms_per_step'] for k in kernels)
        for k in kernels:
            k['share'] = k['ms_per_step'] / tot
        # algorithmic work per impression of each stage
        c_mean = T / B
        if math == _lib.MATH_TABLE:
            # table-level mode: FLOPs actually issued per impression (weighted sums over E and TW, matching + attention dots) and the
            # SURVEY 8(d) algorithmic bytes (gathered table rows + ids + mask + scores + labels); the tw rows the kernel also
            # gathers are extra traffic of this formulation and are reported as `gathered_bytes_per_impression`
            flops = {'proj': (2 * D * DC + 2 * DC * K + 2 * D * D) * (N_NEWS + 1) / B, 'score': 4 * K * H * D + 4 * c_mean * K * D}
            byts = {'proj': (N_NEWS + 1) * D * 2 / B, 'score': synth.algorithmic_bytes_per_impression(H, c_mean, D, 2)}
            tensor_stage = ()
        elif math == _lib.MATH_TENSOR:
            flops = {1: 2 * H * D * DC + 2 * H * DC * K + 2 * K * H * D, 8: 2 * K * D * D + 4 * c_mean * K * D}
            byts = {1: H * D * 2 + H * 8 + H, 8: c_mean * D * 2 + c_mean * 8 + c_mean * 4}
            tensor_stage = (1, 8)
        else:
            flops = {1: 2 * H * D * DC, 2: 2 * H * DC * K + 2 * K * H * D, 4: 2 * K * D * D, 8: 4 * c_mean * K * D}
            byts = {1: H * D * 2 + H * 8, 2: H * D * 2 + H * 8 + H, 4: 0, 8: c_mean * D * 2 + c_mean * 8 + c_mean * 4}
            tensor_stage = ()
        top = max(range(len(stages)), key=lambda i: kernels[i]['ms_per_step'])
        mask = stages[top][1]
        sec_per_launch = kernels[top]['ms_per_step'] * 1e-3 / nchunks
        per_launch_impr = B / nchunks
        tf = flops[mask] * per_launch_impr / sec_per_launch / 1e12
        gbs = byts[mask] * per_launch_impr / sec_per_launch / 1e9
        peak_tf = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath) and math != _lib.MATH_FP32:       # DRAM bytes per impression from the committed ncu capture
            tj = json.load(open(tpath)).get({1: 'hist_kernel', 8: 'cand_kernel', 'score': 'tscore_kernel'}.get(mask, '-'))
            if tj:
                traffic, traffic_src = tj['dram_bytes_per_impression'] * per_launch_impr, tj['source']
        if mask in tensor_stage:
            roofline = {'kernel': stages[top][0], 'bound': 'tensor', 'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                        'frac': tf / peak_tf, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk['_source'] + ' (sustained bf16)',
                        'achieved_hbm_gbs': gbs, 'hbm_frac': gbs / pk['hbm_gbs'],
                        'algorithmic_flops_per_launch': flops[mask] * per_launch_impr, 'ms_per_launch': sec_per_launch * 1e3}
        else:
            roofline = {'kernel': stages[top][0], 'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                        'frac': gbs / pk['hbm_gbs'], 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk['_source'],
                        'achieved_tflops': tf, 'algorithmic_bytes_per_launch': byts[mask] * per_launch_impr, 'ms_per_launch': sec_per_launch * 1e3}
            if math == _lib.MATH_TABLE:
                gathered = (2 * H + c_mean) * D * 2 + H * K * 4 + (H + c_mean) * 8 + H + c_mean * 5     # + tw rows + lg rows
                roofline['gathered_bytes_per_impression'] = gathered
                roofline['gathered_gbs'] = gathered * per_launch_impr / sec_per_launch / 1e9
                roofline['gathered_frac'] = roofline['gathered_gbs'] / pk['hbm_gbs']

    # ---- CPU baseline (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(cores)
        n = args.cpu_sample
        v, dt, _ = cpu_reference_run(n, 3, 1)
        cpu_baseline = {'value': v, 'unit': 'impressions/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': f'{n} impressions x 3 steps of the same workload (oracle/miner_oracle.py, torch CPU fp32 + numpy metrics)'}

    if rank == 0:
        total = B * world
        value = total / (ms_dev * 1e-3)
        e2e_v = total / (ms_e2e * 1e-3)
        bytes_per_impr = synth.algorithmic_bytes_per_impression(H, T / B, D, 2)
        flops_per_impr = synth.algorithmic_flops_per_impression(H, T / B, D, K, DC)
        per_gpu = value / world
        line = {
            'metric': 'impressions scored/sec', 'value': value, 'unit': 'impressions/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32' if math == _lib.MATH_FP32 else 'bf16 operands / f32 accumulate (tcgen05 MMAs, hi+lo splits where fp32 accuracy matters), f32 elsewhere, f64 metrics',
            'data': 'synthetic',
            'config': {'workload': 'MINER eval scoring on 1xB200 per rank: 1M synthetic impressions, history 50, K=32, Dc=200, D=768, '
                                   '~20 candidates each (CSR), 100k-news bf16 table, score_type=weighted, metrics group_auc/mrr/ndcg@5,10/hit@5,10',
                       'math': args.math + (' (both nn.Linear layers applied once per table row INSIDE every step, then one fused scoring kernel)' if math == _lib.MATH_TABLE else ' (reference operation order)'),
                       'impressions_per_gpu_per_step': B, 'candidates_per_gpu_per_step': T, 'chunk_impressions': chunk,
                       'l2': 'inputs per step (>600 MB ids + 154 MB table + workspace) exceed the 126 MB L2; no explicit flush',
                       'parallelism': f'dp{world} (impressions sharded, table+weights replicated)'},
            'e2e': {'value': e2e_v, 'unit': 'impressions/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': 2 * len(names) * 8},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'roofline_path': {'hbm_frac': per_gpu * bytes_per_impr / 1e9 / pk['hbm_gbs'],
                              'tensor_frac_reference_order_flops': per_gpu * flops_per_impr / 1e12 / pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
                              'algorithmic_bytes_per_impression': bytes_per_impr, 'algorithmic_flops_per_impression': flops_per_impr,
                              'peak_source': pk['_source']},
            'kernels': kernels,
            'reference_order': reference_order,
            'cpu_baseline': cpu_baseline,
            'metrics': metrics_out,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
