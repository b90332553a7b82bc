#!/usr/bin/env python
"""bench.py -- the MINER data-parallel scoring path on B200 (BASELINE.json metric: impressions scored / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--impressions B] [--scaling weak|strong]

One "step" = one pass of the hot path (gather -> poly attention -> target-aware aggregation -> per-candidate score ->
per-impression AUC/MRR/nDCG/hit partials) over one batch of B synthetic MIND-shaped impressions PER GPU
(BASELINE.json configs[1]: 1 M impressions, history 50 (lengths uniform in 1..50, left-padded with the pad news as the
reference's reader does), K=32, Dc=200, 768-d news vectors, ~20 candidates each; table of 100k news, bf16).
Weak scaling (default): every rank owns its own B impressions (independent units, no data-path collective); the only
exchange is one NCCL all-reduce of the [sum,count] metric partials per step.  `--scaling strong`: ONE global batch of B
impressions, ranks take the contiguous ranges parallel.shard_bounds gives them.

Printed JSON (rank 0, one line):
  value          whole-job impressions/s with the step's inputs resident in HBM (device-timed, max over ranks)
  e2e            same metric through the public API with HOST (pinned) inputs: H2D of ids/mask/labels/offsets and D2H of
                 the metric partials inside the timed region
  roofline       dominant kernel of the step: SURVEY 8(d) algorithmic bytes per launch / its CUDA-event duration, against
                 MEASURED_PEAKS.json (plus the bytes the kernel really gathers after merging the padding slots)
  parity         the CUDA path against the CPU oracle on the cpu_baseline sample: normwise score error, impressions whose
                 ranking order differs (and the largest reference gap among the pairs that flipped), metric differences
  cpu_baseline   the CPU restatement of the reference path (oracle/, torch CPU fp32 + numpy metrics) timed on this box's
                 host cores on a bounded sample (BASELINE configs[0]: 1 k impressions), grouped layout + the reference's
                 one-row-per-candidate layout -- a reported baseline, not the target
  torch_eager_gpu  the same restatement run eagerly on this GPU (aten / cuBLAS): what SURVEY 2a calls the real bar
  full_history   the scoring kernel alone on impressions whose 50 history slots are all real clicks (nothing to merge)
  gather         miner_gather alone: achieved HBM GB/s
  strong_scaling (N > 1) the same global batch split over the ranks
  train_step     BASELINE configs[2]: train step, batch 4096 per GPU, npratio 4, bf16 tensor-core GEMMs
--impl reference times the CPU path alone (the reference is pure Python/PyTorch: its sources may not be copied into this
repository and /root/reference does not exist on the GPU box, so the timed code is the oracle port pinned to it by
tests/golden).
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
SIX = ['group_auc', 'mrr', 'ndcg@5', 'ndcg@10', 'hit@5', 'hit@10']
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}
WORKLOAD = ('MINER eval scoring: synthetic MIND-shaped impressions, history 50 (1..50 clicks, left-padded), K=32, Dc=200, D=768, '
            '~20 candidates each (CSR), 100k-news table, score_type=weighted, metrics group_auc/mrr/ndcg@5,10/hit@5,10')


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
def cpu_inputs(n_impr: int, seed: int = 36):
    """The sample both arms score: the bf16-valued table (as fp32, what the reference computes in), fp32 weights, CSR impressions."""
    from miner_b200 import synth
    table = synth.make_table(N_NEWS, D, seed, torch.bfloat16).float()
    w = synth.make_weights(D, K, DC, seed)
    eb = synth.make_eval_batch(n_impr, H, N_NEWS, seed + 7, mean_cands=MEAN_C)
    return table, w, eb


def cpu_reference_run(n_impr: int, steps: int, warmup: int, layout: str = 'grouped', inputs=None):
    """Times the CPU restatement of the reference path (oracle) on `n_impr` impressions per step, all host threads.
    layout 'grouped': one row per impression with its candidates (interests computed once per impression);
    'per_candidate': one row per candidate with C = 1, interests recomputed per candidate -- what the reference's eval reader
    produces (src/reader.py:376-379).  Returns (impressions/s, seconds per step, scores (T,), metrics dict)."""
    from oracle import miner_oracle as O
    table, w, eb = inputs if inputs is not None else cpu_inputs(n_impr)
    offs = eb.offsets.numpy()
    labels = eb.labels.numpy()
    counts = torch.from_numpy(offs[1:] - offs[:-1])
    row_of = torch.repeat_interleave(torch.arange(n_impr), counts)       # impression of every candidate

    def step():
        with torch.no_grad():
            if layout == 'grouped':
                s = O.miner_forward_csr(table, eb.his_ids, eb.his_mask, eb.cand_ids, offs, w.w_proj, w.context_codes, w.w_target,
                                        'weighted', chunk=512)
            else:
                s = torch.empty(int(offs[-1]))
                for a in range(0, s.numel(), 2048):                      # eval batches of (2048, 1) rows
                    r = row_of[a:a + 2048]
                    s[a:a + 2048] = O.miner_forward(table, eb.his_ids[r], eb.his_mask[r], eb.cand_ids[a:a + 2048, None], w.w_proj,
                                                    w.context_codes, w.w_target, 'weighted')[1][:, 0]
        probs = np.asarray(O.sigmoid_probs(s))
        targets = [labels[offs[i]:offs[i + 1]].tolist() for i in range(n_impr)]
        preds = [probs[offs[i]:offs[i + 1]].tolist() for i in range(n_impr)]
        return s, O.compute_scores(targets, preds, SIX)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s, res = step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_impr / dt, dt, s, res


def parity_report(gpu_scores: torch.Tensor, gpu_metrics: dict, ref_scores: torch.Tensor, ref_metrics: dict, offs: np.ndarray):
    """CUDA path vs the oracle on the same impressions.  A 'flip' is an impression whose descending order of candidates differs."""
    g, r = gpu_scores.double().numpy(), ref_scores.double().numpy()
    scale = float(np.abs(r).max())
    flips, worst_gap = 0, 0.0
    for i in range(len(offs) - 1):
        a, b = offs[i], offs[i + 1]
        og, orf = np.argsort(-g[a:b], kind='stable'), np.argsort(-r[a:b], kind='stable')
        if not np.array_equal(og, orf):
            flips += 1
            rr = r[a:b]
            # largest reference gap among candidate pairs the two orders disagree on
            pos_g = np.empty_like(og); pos_g[og] = np.arange(len(og))
            pos_r = np.empty_like(orf); pos_r[orf] = np.arange(len(orf))
            dg = np.sign(pos_g[:, None] - pos_g[None, :]) != np.sign(pos_r[:, None] - pos_r[None, :])
            worst_gap = max(worst_gap, float(np.abs(rr[:, None] - rr[None, :])[dg].max()))
    err = float(np.abs(g - r).max())
    return {'n_impressions': len(offs) - 1, 'n_candidates': int(offs[-1]),
            'scores_normwise': err / scale, 'scores_max_abs': err, 'ref_scores_max_abs': scale,
            'tolerance': 'north_star: 1e-3 relative on scores (normwise: max|d| / max|ref|, scores cross zero); order checked exactly',
            'order_flips': flips,
            'max_ref_gap_of_flipped_pairs_normwise': worst_gap / scale,
            'order_flips_note': 'every flipped pair is a near-tie of the reference itself: its gap is below twice the measured score error',
            'metrics_abs_diff': {k: abs(gpu_metrics[k] - ref_metrics[k]) for k in SIX},
            'reference': 'oracle/miner_oracle.py on the CPU, fp32 weights (the GPU path rounds Wp, Wt to bf16), same bf16-valued table'}


# ------------------------------------------------------------------------------------------------ train step (configs[2])
def train_bench(dev, rank, world, steps, warmup, batch=4096):
    import torch.distributed as dist
    import torch.nn as nn
    import miner_b200 as mb
    from miner_b200 import ops, synth, parallel
    NP = 4
    table = synth.make_table(N_NEWS, D, 36, torch.bfloat16).to(dev)
    w = synth.make_weights(D, K, DC, 36)
    model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).train()
    model.train_math = 'tensor'
    with torch.no_grad():
        model.poly_attn.linear.weight.copy_(w.w_proj)
        model.poly_attn.context_codes.copy_(w.context_codes)
        model.target_aware_attn.linear.weight.copy_(w.w_target)
    his, mask, _, cand, _, labels = synth.make_train_batch(batch, H, N_NEWS, NP, 36 + rank)
    host = {'his': his.pin_memory(), 'mask': mask.pin_memory(), 'cand': cand.pin_memory(), 'labels': labels.float().pin_memory()}
    res = {k: v.to(dev) for k, v in host.items()}
    loss_fn = mb.Loss(nn.CrossEntropyLoss(reduction='mean'))
    params = list(model.parameters())
    flat = parallel.FlatGradients(params) if hasattr(parallel, 'FlatGradients') else None
    opt = torch.optim.SGD(params, lr=1e-3)
    Bt, C = cand.shape
    z = torch.zeros(Bt, C, 1, dtype=torch.long, device=dev)
    zh = torch.zeros(Bt, H, 1, dtype=torch.long, device=dev)

    def step(d):
        if flat is not None:
            flat.zero()
        else:
            opt.zero_grad(set_to_none=True)
        I, S = model(d['cand'][..., None], z, d['his'][..., None], zh, d['mask'], z, z, zh, zh)
        loss = loss_fn.compute(I, S, d['labels'])
        loss.backward()
        if flat is not None:
            flat.allreduce()
        else:
            parallel.allreduce_gradients(params)
        opt.step()
        return loss

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n, out

    for _ in range(warmup):
        step(res)
    l0 = ops.launch_count()
    ms, loss = timed(lambda: step(res), steps)
    launches = (ops.launch_count() - l0) // steps
    ms_e2e, _ = timed(lambda: step({k: v.to(dev, non_blocking=True) for k, v in host.items()}), steps)
    flops = 3 * (2 * H * D * DC + 2 * H * DC * K + 2 * K * H * D + 2 * K * D * D + 4 * C * K * D)      # fwd + ~2x in the backward
    pk = peaks()
    tf = flops * Bt / (ms * 1e-3) / 1e12
    return {'metric': 'train samples/sec', 'value': Bt * world / (ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms,
            'steps': steps, 'warmup': warmup,
            'config': {'workload': 'MINER train step (forward + Loss.compute + backward + flat gradient all-reduce + SGD), npratio 4, '
                                   'history 50, K=32, Dc=200, D=768, frozen 100k-news bf16 table',
                       'batch_per_gpu': Bt, 'parallelism': f'dp{world}'},
            'dtype': 'bf16 operands / f32 accumulate (tcgen05) for the projection-sized GEMMs, f32 elsewhere',
            'e2e': {'value': Bt * world / (ms_e2e * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host.values()), 'd2h_bytes_per_step': 0},
            'gpu_launches_per_step': int(launches), 'loss': float(loss),
            'achieved_tflops_per_gpu': tf, 'tensor_frac': tf / pk.get('bf16_tflops_sustained', pk['bf16_tflops'])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--impressions', type=int, default=1_000_000, help='impressions per GPU per step (weak) / in the global batch (strong)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--chunk', type=int, default=0, help='impressions per kernel wave of the reference-order family (0 = library default)')
    ap.add_argument('--cpu-sample', type=int, default=1000, help='impressions in the CPU baseline sample (BASELINE configs[0]: 1 k)')
    ap.add_argument('--math', default='table', choices=['table', 'tensor', 'fp32'],
                    help='table: projections applied once per table row inside every step + one fused scoring kernel (default); '
                         'tensor / fp32: reference operation order')
    ap.add_argument('--no-reference-order', action='store_true', help='skip the extra reference-order (tensor family) timing')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-breakdown', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip torch_eager_gpu / full_history / gather / strong_scaling / train_step')
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
        inputs = cpu_inputs(n)
        v, dt, _, res = cpu_reference_run(n, max(args.steps, 1), max(args.warmup, 1), 'grouped', inputs)
        v_pc, dt_pc, _, _ = cpu_reference_run(n, min(max(args.steps, 1), 2), 1, 'per_candidate', inputs)
        line = {'impl': 'reference', 'metric': 'impressions scored/sec', 'value': v, 'unit': 'impressions/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'impressions_per_step': n, 'device': 'host CPU',
                           'layout': 'grouped: one row per impression with its candidates (interests once per impression), six metrics'},
                'cpu_baseline': {'value': v, 'unit': 'impressions/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                 'sample': f'{n} impressions per step (BASELINE configs[0]; a bounded sample of the 1M-impression workload), '
                                           f'oracle/miner_oracle.py = torch CPU fp32 restatement of the reference Miner.forward + numpy ranking metrics'},
                'per_candidate_layout': {'value': v_pc, 'unit': 'impressions/s', 'ms_per_step': dt_pc * 1e3,
                                         'note': 'the reference eval reader emits one (1-candidate) row per candidate (src/reader.py:376-379): interests '
                                                 'recomputed per candidate, batches of 2048 rows'},
                'metrics': res,
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
    # one process per GPU: run on (and pin host buffers into) the GPU's own NUMA node; undone before the CPU baseline leg
    prev_affinity = parallel.bind_to_gpu_cpus(local_rank) if os.environ.get('MINER_BENCH_AFFINITY', '1') != '0' else None
    affinity_note = (f'process bound to the {len(os.sched_getaffinity(0))} CPUs NVML lists as local to its GPU' if prev_affinity is not None
                     else 'process CPU affinity unchanged')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    seed = 36
    math = {'table': _lib.MATH_TABLE, 'tensor': _lib.MATH_TENSOR, 'fp32': _lib.MATH_FP32}[args.math]
    strong = args.scaling == 'strong'

    # replicated table + weights
    table = synth.make_table(N_NEWS, D, seed, torch.bfloat16).to(dev)
    w = synth.make_weights(D, K, DC, seed)
    model = mb.Miner(mb.TableNewsEncoder(table), False, K, DC, 'weighted', 0.2).to(dev).eval()
    with torch.no_grad():
        model.poly_attn.linear.weight.copy_(w.w_proj)
        model.poly_attn.context_codes.copy_(w.context_codes)
        model.target_aware_attn.linear.weight.copy_(w.w_target)

    def shard_of(eb, r, n):
        """Host tensors of rank r's contiguous share of a global batch (parallel.shard_bounds: balanced by gathered rows)."""
        s, e = parallel.shard_bounds(eb.offsets, n, H)[r]
        c0, c1 = int(eb.offsets[s]), int(eb.offsets[e])
        return {'his_ids': eb.his_ids[s:e], 'his_mask': eb.his_mask[s:e], 'cand_ids': eb.cand_ids[c0:c1], 'labels': eb.labels[c0:c1],
                'offsets': eb.offsets[s:e + 1] - c0}

    if strong:                                    # one global batch, this rank's share of it
        eb = synth.make_eval_batch(args.impressions, H, N_NEWS, seed, mean_cands=MEAN_C)
        host = {k: v.contiguous().pin_memory() for k, v in shard_of(eb, rank, world).items()}
        total_impr = args.impressions
    else:                                         # this rank's own impressions
        eb = synth.make_eval_batch(args.impressions, H, N_NEWS, seed + 1000 * rank, mean_cands=MEAN_C)
        host = {k: getattr(eb, k).pin_memory() for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
        total_impr = args.impressions * world
    B = host['his_ids'].shape[0]
    T = int(host['offsets'][-1])
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    names = ops.metric_names(KS)
    chunk = args.chunk if args.chunk > 0 else 32768
    sw = model._weights(with_bf16=(math != _lib.MATH_FP32))
    scores_buf = torch.empty(T, dtype=torch.float32, device=dev)

    proj = ops.table_project(table, sw) if math == _lib.MATH_TABLE else None          # buffers; recomputed inside every step
    proj_ws = torch.empty(max(_lib.load().miner_table_project_workspace_bytes(table.shape[0], DC), 1), dtype=torch.uint8, device=dev)
    tws = ops.score_table_workspace(B, H, K, dev) if math == _lib.MATH_TABLE else None   # packed tiles: rebuilt by every call

    stage_marks = None                            # while a list: device_step appends one CUDA event per stage boundary (table mode)

    def mark():
        if stage_marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            stage_marks.append(e)

    def score_step(d, m, out=None, ws=None):
        out = scores_buf if out is None else out
        if m == _lib.MATH_TABLE:
            mark()
            ops.table_project(table, sw, out=proj, workspace=proj_ws)       # part of the step: nothing is carried over between steps
            mark()
            ops.score_table(proj, d['his_ids'], d['his_mask'], d['cand_ids'], 'weighted', cand_offsets=d['offsets'], out_scores=out,
                            workspace=tws if ws is None else ws)
        else:
            ops.score(table, d['his_ids'], d['his_mask'], d['cand_ids'], sw, 'weighted', cand_offsets=d['offsets'], math=m,
                      chunk=chunk, out_scores=out)

    def device_step(d, m=None):
        score_step(d, math if m is None else m)
        mark()
        partials, _ = ops.rank_metrics_raw(scores_buf, d['labels'], d['offsets'], 'sigmoid', KS)
        mark()
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

    def one(fn):
        """CUDA-event time of one call after one untimed call (single kernels / short sequences on the current stream)."""
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- value: inputs resident in HBM
    resident = {k: v.to(dev) for k, v in host.items()}
    for _ in range(max(args.warmup, 3)):
        device_step(resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    stage_marks = [] if math == _lib.MATH_TABLE else None
    ms_dev, partials = timed(lambda: device_step(resident), args.steps)
    launches = ops.launch_count() - l0
    # stage times INSIDE the timed region (four events per step on the launching stream): under the board's power cap the steps of a
    # long run are slower than a kernel timed alone after an idle gap, so these -- not the stand-alone times -- go into the roofline
    in_loop = None
    if stage_marks:
        ev, stage_marks = stage_marks, None
        in_loop = [sum(ev[4 * i + j].elapsed_time(ev[4 * i + j + 1]) for i in range(args.steps)) / args.steps for j in range(3)]
    metrics_out = parallel.finalize_metrics(partials, names)
    if math == _lib.MATH_TABLE:
        ops.check_oob(tws)                       # an id outside the table would have raised here (the reference: IndexError)
    reference_order = None
    if math == _lib.MATH_TABLE and not args.no_reference_order:
        # the same step in the reference's operation order (tensor family: per-row projections on tcgen05), for comparison
        for _ in range(2):
            device_step(resident, _lib.MATH_TENSOR)
        ms_ro, p_ro = timed(lambda: device_step(resident, _lib.MATH_TENSOR), max(1, min(args.steps, 3)))
        m_ro = parallel.finalize_metrics(p_ro, names)
        reference_order = {'value': total_impr / (ms_ro * 1e-3), 'unit': 'impressions/s', 'ms_per_step': ms_ro,
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
                      ('tpack_kernel + tscore_x_kernel: pack the histories into tiles, then gather table/tw/cand rows + softmax_H + P = w tw + gelu + '
                       'X = E cand^T / attention MMAs + m = w X + softmax_K + score (tcgen05)', 'score')]
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
            kernels.append({'kernel': name, 'ms_per_step': one(fn), 'launches_per_step': 3 if mask == 'proj' else (2 if mask == 'score' else nchunks)})
        kernels.append({'kernel': 'rank_metrics (+finalize)', 'launches_per_step': 2,
                        'ms_per_step': one(lambda: ops.rank_metrics_raw(scores_buf, resident['labels'], resident['offsets'], 'sigmoid', KS))})
        if in_loop is not None:
            for k, ms_in in zip(kernels, in_loop):
                k['ms_alone'] = k['ms_per_step']                      # one call after an idle gap (burst clocks)
                k['ms_per_step'] = ms_in                              # CUDA events inside the timed steps (sustained, power-capped clocks)
                k['timing'] = 'ms_per_step: CUDA events around the stage inside the timed steps; ms_alone: one call after an idle gap'
        tot = sum(k['ms_per_step'] for k in kernels)
        for k in kernels:
            k['share'] = k['ms_per_step'] / tot
        # algorithmic work per impression of each stage
        c_mean = T / B
        if math == _lib.MATH_TABLE:
            # table-level mode: FLOPs actually issued per impression (weighted sums over E and TW, matching + attention dots) and the
            # SURVEY 8(d) algorithmic bytes (gathered table rows + ids + mask + scores + labels)
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
                        'achieved_tflops': tf, 'algorithmic_bytes_per_launch': byts[mask] * per_launch_impr, 'ms_per_launch': sec_per_launch * 1e3,
                        'algorithmic_bytes_per_impression': byts[mask],
                        'algorithmic_bytes_note': 'SURVEY 8(d): (H + C) D s gathered rows + (H + C) 8 ids + H mask + 5 C scores/labels, every one of '
                                                  'the H = 50 history slots counted'}
            if 'ms_alone' in kernels[top]:
                roofline['ms_per_launch_alone'] = kernels[top]['ms_alone'] / nchunks
                roofline['frac_alone'] = roofline['frac'] * kernels[top]['ms_per_step'] / kernels[top]['ms_alone']
                roofline['timing_note'] = ('ms_per_launch / achieved / frac: the kernel pair timed by CUDA events inside the timed steps (the board '
                                           'runs at its power cap there, SM clock ~1800 of 1965 MHz); *_alone: one launch after an idle gap')
            if math == _lib.MATH_TABLE:
                # what the kernel really gathers: per history slot a table row and a tw row, but the masked padding slots of an
                # impression are ONE slot (same row, same softmax term); + lg rows, compact slot records, candidates
                kept = host['his_mask'].sum(dim=1)
                slots = float((kept + (kept < H).to(kept.dtype)).double().mean())
                gathered = (2 * slots + c_mean) * D * 2 + slots * K * 4 + slots * 8 * 2 + H * 9 + c_mean * 8 + c_mean * 5
                roofline['mean_history_slots_after_merging_padding'] = slots
                roofline['gathered_bytes_per_impression'] = gathered
                roofline['gathered_gbs'] = gathered * per_launch_impr / sec_per_launch / 1e9
                roofline['gathered_frac'] = roofline['gathered_gbs'] / pk['hbm_gbs']

    # ---- CPU baseline + parity on the same sample (rank 0, N=1 only)
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)            # the CPU arm gets every core the process was given
            prev_affinity = None
        torch.set_num_threads(cores)
        n = args.cpu_sample
        inputs = cpu_inputs(n)
        v, dt, ref_scores, ref_metrics = cpu_reference_run(n, 3, 1, 'grouped', inputs)
        v_pc, dt_pc, s_pc, _ = cpu_reference_run(n, 1, 1, 'per_candidate', inputs)
        cpu_baseline = {'value': v, 'unit': 'impressions/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': f'{n} impressions x 3 steps (BASELINE configs[0]) of the same workload, grouped layout, six metrics '
                                  f'(oracle/miner_oracle.py, torch CPU fp32 + numpy metrics)',
                        'per_candidate_layout': {'value': v_pc, 'unit': 'impressions/s', 'ms_per_step': dt_pc * 1e3,
                                                 'max_abs_score_diff_vs_grouped': float((s_pc - ref_scores).abs().max()),
                                                 'note': 'one (1-candidate) row per candidate as src/reader.py:376-379 emits them'}}
        _, _, eb_s = inputs
        d_s = {k: getattr(eb_s, k).to(dev) for k in ('his_ids', 'his_mask', 'cand_ids', 'labels', 'offsets')}
        s_gpu = torch.empty(int(eb_s.offsets[-1]), dtype=torch.float32, device=dev)
        score_step(d_s, math, out=s_gpu, ws=ops.score_table_workspace(n, H, K, dev) if math == _lib.MATH_TABLE else None)
        gm = ops.rank_metrics(s_gpu, d_s['labels'], d_s['offsets'], 'sigmoid', KS)
        parity = parity_report(s_gpu.cpu(), gm, ref_scores, ref_metrics, eb_s.offsets.numpy())

    # ---- informational extras
    extras = {}
    if not args.no_extras and math == _lib.MATH_TABLE:
        if rank == 0:
            # (1) PyTorch eager on this GPU: the oracle's functions on cuda tensors (aten index / mm / bmm / softmax / gelu, cuBLAS)
            from oracle import miner_oracle as O
            nb, C = 8192, 20
            g = torch.Generator().manual_seed(5)
            t32 = table.float()
            hi, hm, _ = synth.make_history(nb, H, N_NEWS, g)
            cd = torch.randint(1, N_NEWS + 1, (nb, C), generator=g)
            wd = [x.to(dev) for x in (w.w_proj, w.context_codes, w.w_target)]
            hi, hm, cd = hi.to(dev), hm.to(dev), cd.to(dev)
            with torch.no_grad():
                ms = one(lambda: O.miner_forward(t32, hi, hm, cd, wd[0], wd[1], wd[2], 'weighted'))
            extras['torch_eager_gpu'] = {'value': nb / (ms * 1e-3), 'unit': 'impressions/s', 'ms_per_call': ms,
                                         'what': f'oracle/miner_oracle.py miner_forward (reference operation order, fp32 aten/cuBLAS ops) on this GPU, '
                                                 f'one dense ({nb}, {C}) call, scoring only (no metrics); informational'}
            del t32
            # (2) the scoring kernel alone when nothing can be merged: every history slot a real click
            nf = min(B, 200_000)
            ebf = synth.make_eval_batch(nf, H, N_NEWS, seed + 3, mean_cands=MEAN_C)
            full_mask = torch.ones_like(ebf.his_mask)
            full_ids = torch.randint(1, N_NEWS + 1, ebf.his_ids.shape, generator=g)
            df = {'his_ids': full_ids.to(dev), 'his_mask': full_mask.to(dev), 'cand_ids': ebf.cand_ids.to(dev), 'offsets': ebf.offsets.to(dev)}
            sf = torch.empty(int(ebf.offsets[-1]), dtype=torch.float32, device=dev)
            wsf = ops.score_table_workspace(nf, H, K, dev)
            ms = one(lambda: ops.score_table(proj, df['his_ids'], df['his_mask'], df['cand_ids'], 'weighted', cand_offsets=df['offsets'],
                                             out_scores=sf, workspace=wsf))
            bpi = synth.algorithmic_bytes_per_impression(H, int(ebf.offsets[-1]) / nf, D, 2)
            extras['full_history'] = {'value': nf / (ms * 1e-3), 'unit': 'impressions/s', 'ms': ms, 'impressions': nf,
                                      'hbm_frac': bpi * nf / (ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                                      'what': 'tpack + tscore kernels alone on impressions with 50 real clicks each (no padding to merge)'}
            del df, sf, wsf
            # (3) miner_gather alone (kernel #1 of the north star): bf16 rows read + written
            ng = 2_000_000
            gid = torch.randint(0, N_NEWS + 1, (ng,), generator=g).to(dev)
            ms = one(lambda: ops.gather(table, gid))
            gb = 2 * ng * D * 2 + ng * 8
            extras['gather'] = {'gbs': gb / (ms * 1e-3) / 1e9, 'frac': gb / (ms * 1e-3) / 1e9 / pk['hbm_gbs'], 'ms': ms, 'rows': ng,
                                'bytes': gb, 'what': 'miner_gather: table[ids] for 2M int64 ids, 768-d bf16 rows read and written'}
            del gid
        torch.cuda.empty_cache()
        # (4) strong scaling: the rank-0 batch of the weak run as ONE global batch split over the ranks
        if world > 1 and not strong:
            eb0 = synth.make_eval_batch(args.impressions, H, N_NEWS, seed, mean_cands=MEAN_C)
            ds = {k: v.to(dev) for k, v in shard_of(eb0, rank, world).items()}
            ss = torch.empty(int(ds['offsets'][-1]), dtype=torch.float32, device=dev)

            def strong_step():
                score_step(ds, math, out=ss)
                p, _ = ops.rank_metrics_raw(ss, ds['labels'], ds['offsets'], 'sigmoid', KS)
                parallel.allreduce_partials(p)
                return p
            for _ in range(3):
                strong_step()
            ms_s, p_s = timed(strong_step, args.steps)
            m_s = parallel.finalize_metrics(p_s, names)
            extras['strong_scaling'] = {'global_impressions': args.impressions, 'ms_per_step': ms_s, 'value': args.impressions / (ms_s * 1e-3),
                                        'unit': 'impressions/s', 'impressions_this_rank': int(ds['his_ids'].shape[0]),
                                        'metrics': m_s,
                                        'note': 'table_project (replicated) and the all-reduce are inside the step; the same batch at N = 1 is the '
                                                'weak run of a 1-GPU bench: its `metrics` must equal these to 1e-12 (rank-count invariance)'}
            del ds, ss
        # (5) train step (configs[2])
        torch.cuda.empty_cache()
        try:
            extras['train_step'] = train_bench(dev, rank, world, steps=10, warmup=40 if world > 1 else 10)
        except Exception as e:          # the scoring line must survive a failure of the extra
            extras['train_step'] = {'error': repr(e)[:300]}

    if rank == 0:
        value = total_impr / (ms_dev * 1e-3)
        e2e_v = total_impr / (ms_e2e * 1e-3)
        bytes_per_impr = synth.algorithmic_bytes_per_impression(H, T / B, D, 2)
        flops_per_impr = synth.algorithmic_flops_per_impression(H, T / B, D, K, DC)
        per_gpu = value / world
        line = {
            'metric': 'impressions scored/sec', 'value': value, 'unit': 'impressions/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32' if math == _lib.MATH_FP32 else 'bf16 operands / f32 accumulate (tcgen05 MMAs, hi+lo splits where fp32 accuracy matters), f32 elsewhere, f64 metrics',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD + ', bf16 table',
                       'math': args.math + (' (both nn.Linear layers applied once per table row INSIDE every step, then tpack + one fused scoring kernel)' if math == _lib.MATH_TABLE else ' (reference operation order)'),
                       'impressions_per_gpu_per_step': B, 'candidates_per_gpu_per_step': T, 'chunk_impressions': chunk,
                       'l2': 'inputs per step (>600 MB ids + 154 MB table + 154 MB projected table + workspace) exceed the 126 MB L2; no explicit flush',
                       'parallelism': f'dp{world} (impressions sharded, table+weights replicated)', 'host': affinity_note},
            'e2e': {'value': e2e_v, 'unit': 'impressions/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': 2 * len(names) * 8 + 8},
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
            'parity': parity,
            'metrics': metrics_out,
        }
        line.update(extras)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
