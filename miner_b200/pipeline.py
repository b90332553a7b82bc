"""Host-to-metrics evaluation pipeline: the call a user makes to score a block of impressions that lives in (pinned)
host memory and get the ranking metrics back.

Waves of impressions are copied host->device on a copy stream, into one of two persistent sets of device buffers, while the
previous wave is being scored (table-level tpack + scoring kernel, or hist_kernel -> cand_kernel in reference order, then
rank_metrics) on the compute stream; the per-wave [sum, count] metric partials add up (exactly what ranks all-reduce,
parallel.allreduce_partials).  The first wave is small (its copy is the only one nothing hides), later waves double up to
4 x `wave`.  Everything on the device runs in the sm_100a kernels of
libminer_b200.so; this file is stream / buffer plumbing only.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops
from . import _lib as L


def wave_bounds(B: int, first_wave: int, wave_growth: float, max_wave: int):
    """Impression boundaries ``[0, b1, ..., B]`` of the waves: the first has ``first_wave`` impressions, each later one ``wave_growth``
    times the previous up to ``max_wave``; every boundary but the last is a multiple of 4 (tile boundaries of the table-level kernel)."""
    first_wave = max(4, int(first_wave) // 4 * 4)
    max_wave = max(first_wave, int(max_wave) // 4 * 4)
    bounds, w = [0], first_wave
    while bounds[-1] < B:
        bounds.append(min(B, bounds[-1] + w))
        w = min(max_wave, max(4, int(w * max(1.0, float(wave_growth))) // 4 * 4))
    return bounds


class HostEvaluator:
    def __init__(self, model, wave: int = 65536, chunk: int = 32768, ks: Sequence[int] = (5, 10), transform: str = 'sigmoid',
                 math: Optional[int] = None, check_bounds: bool = True, first_wave: Optional[int] = None, wave_growth: float = 2.0,
                 max_wave: Optional[int] = None):
        from .model import TableNewsEncoder
        if not isinstance(model.news_encoder, TableNewsEncoder):
            raise L.MinerError('HostEvaluator needs a Miner with a TableNewsEncoder')
        if model.use_category_bias:
            raise NotImplementedError('grouped scoring with category bias: use Miner.forward with the row layout you mean (model.py:176)')
        # waves start on tile boundaries of the table-level kernel (4 impressions at most share a tile): an impression then keeps its
        # tile partner -- and every bit of its scores -- whatever the wave size
        self.model, self.wave, self.chunk, self.ks, self.transform = model, max(4, int(wave) // 4 * 4), int(chunk), tuple(ks), transform
        self.table = model.news_encoder.table
        self.dev = self.table.device
        self._math_arg = math
        self.math = math
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.check_bounds = check_bounds
        # wave schedule: the first wave's copy is the only one nothing hides, so it may be smaller (`first_wave`); later waves grow by
        # `wave_growth` up to `max_wave` (fewer launch ramps / tails) as long as a wave's copy still fits under the previous wave's scoring
        # (defaults: wave / 4, doubling up to 4 x wave -- measured 32.7 against 33.5 ms per 1 M impressions with 16 equal waves of 65 536;
        # pinned H2D runs at 55 GB/s = 11.5 ms of copies under 30 ms of scoring, profiles/r02_e2e_waves.txt)
        self.first_wave = max(4, int(first_wave if first_wave is not None else self.wave // 4) // 4 * 4)
        self.wave_growth = max(1.0, float(wave_growth))
        self.max_wave = max(self.first_wave, int(max_wave if max_wave is not None else 4 * self.wave) // 4 * 4)

    def _wave_bounds(self, B: int):
        return wave_bounds(B, self.first_wave, self.wave_growth, self.max_wave)

    def _proj_ws(self) -> torch.Tensor:
        n = L.load().miner_table_project_workspace_bytes(self.table.shape[0], self.model.poly_attn.context_codes.shape[1])
        if getattr(self, '_pws', None) is None or self._pws.numel() < n:
            self._pws = torch.empty(max(n, 1), dtype=torch.uint8, device=self.dev)
        return self._pws

    def _buffers(self, host: Dict[str, torch.Tensor], bounds, cb):
        """Two persistent sets of device buffers for the waves (kept across calls while they are large enough): no allocator traffic and
        no cross-stream recycling inside the timed path.  A set is rewritten by the copy stream only after the kernels that read it."""
        n_imp = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
        n_cand = max(max(b - a for a, b in zip(cb[:-1], cb[1:])), 1)
        H = host['his_ids'].shape[1]
        key = (H, host['his_ids'].dtype, host['his_mask'].dtype, host['cand_ids'].dtype, host['labels'].dtype)
        cur = getattr(self, '_bufs', None)
        if cur is None or cur['key'] != key or cur['n_imp'] < n_imp or cur['n_cand'] < n_cand:
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.dev)
            sets = [{'his_ids': mk((n_imp, H), key[1]), 'his_mask': mk((n_imp, H), key[2]), 'cand_ids': mk((n_cand,), key[3]),
                     'labels': mk((n_cand,), key[4]), 'offsets': mk((n_imp + 1,), torch.int64), 'scores': mk((n_cand,), torch.float32),
                     'free': None} for _ in range(2)]
            cur = self._bufs = {'key': key, 'n_imp': n_imp, 'n_cand': n_cand, 'sets': sets}
        return cur['sets']

    def _stage(self, host: Dict[str, torch.Tensor], buf, a: int, b: int, c0: int, c1: int):
        """Issue the H2D copies of impressions [a, b) into one buffer set on the copy stream; returns views of it + the event to wait on."""
        with torch.cuda.stream(self.copy_stream):
            if buf['free'] is not None:
                self.copy_stream.wait_event(buf['free'])          # the kernels of the wave that used this set last are done
            d = {}
            for k, lo, hi in (('his_ids', a, b), ('his_mask', a, b), ('cand_ids', c0, c1), ('labels', c0, c1), ('offsets', a, b + 1)):
                d[k] = buf[k][:hi - lo]
                d[k].copy_(host[k][lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        d['scores'] = buf['scores'][:c1 - c0]
        d['c0'] = c0
        return d, ev

    @torch.no_grad()
    def evaluate(self, host: Dict[str, torch.Tensor], want_scores: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """host: pinned CPU tensors his_ids (B,H), his_mask (B,H), cand_ids (T,), labels (T,), offsets (B+1,).
        Returns (partials (2*M,) float64 on the device: [sum, count] per metric of ops.metric_names(ks), scores (T,) or None)."""
        B = host['his_ids'].shape[0]
        Hh, Kk = host['his_ids'].shape[1], self.model.poly_attn.context_codes.shape[0]
        compute = torch.cuda.current_stream(self.dev)
        # the wave buffers are allocated on (and may be recycled from blocks last used on) the compute stream: the copy stream must not
        # write into them ahead of the work already queued there
        self.copy_stream.wait_stream(compute)
        if self._math_arg is None:
            self.math = ops.default_eval_math(self.table, host['his_ids'].shape[1], self.model.poly_attn.context_codes.shape[0])
        w = self.model._weights(with_bf16=(self.math != L.MATH_FP32))
        if self.math == L.MATH_TABLE:
            # the table-level projections belong to the call: they are recomputed here, not carried over between calls
            self._proj = ops.table_project(self.table, w, weighted=self.model.score_type == 'weighted', out=getattr(self, '_proj', None),
                                           workspace=self._proj_ws())
        total = None
        scores_all = torch.empty(int(host['offsets'][-1]), dtype=torch.float32, device=self.dev) if want_scores else None
        bounds = self._wave_bounds(B)
        n_waves = len(bounds) - 1
        cb = [int(host['offsets'][b]) for b in bounds]
        sets = self._buffers(host, bounds, cb) if n_waves > 0 else None
        stage = lambda i: self._stage(host, sets[i & 1], bounds[i], bounds[i + 1], cb[i], cb[i + 1])
        nxt = stage(0) if n_waves > 0 else None
        for i in range(n_waves):
            d, ev = nxt
            nxt = stage(i + 1) if i + 1 < n_waves else None
            compute.wait_event(ev)
            offs = d['offsets'].sub_(d['c0'])
            if self.math == L.MATH_TABLE:
                nb = d['his_ids'].shape[0]
                if getattr(self, '_tws', None) is None or self._tws.numel() < L.load().miner_score_table_workspace_bytes(nb, Hh, Kk):
                    self._tws = ops.score_table_workspace(max(nb, self.max_wave), Hh, Kk, self.dev)
                _, s = ops.score_table(self._proj, d['his_ids'], d['his_mask'], d['cand_ids'], self.model.score_type, cand_offsets=offs,
                                       out_scores=d['scores'], workspace=self._tws)
            else:
                _, s = ops.score(self.table, d['his_ids'], d['his_mask'], d['cand_ids'], w, self.model.score_type, cand_offsets=offs,
                                 math=self.math, chunk=self.chunk, out_scores=d['scores'])
            p, _ = ops.rank_metrics_raw(s, d['labels'], offs, self.transform, self.ks)
            total = p if total is None else total + p
            if want_scores:
                scores_all[d['c0']:d['c0'] + s.numel()] = s
            free = torch.cuda.Event()
            free.record(compute)
            sets[i & 1]['free'] = free
        if total is None:
            total = torch.zeros(2 * (2 + 2 * len(self.ks)), dtype=torch.float64, device=self.dev)
        if self.check_bounds and getattr(self, '_tws', None) is not None:
            # ids outside the table raise what the reference's indexing raises (one 8-byte read at the end of the call)
            try:
                ops.check_oob(self._tws)
            finally:
                self._tws[:8].zero_()
        return total, scores_all
