"""Host-to-metrics evaluation pipeline: the call a user makes to score a block of impressions that lives in (pinned)
host memory and get the ranking metrics back.

Waves of `wave` impressions are copied host->device on a copy stream while the previous wave is being scored
(table-level tscore_kernel, or hist_kernel -> cand_kernel in reference order, then rank_metrics) on the compute stream; the per-wave [sum, count] metric partials add up
(exactly what ranks all-reduce, parallel.allreduce_partials).  Everything on the device runs in the sm_100a kernels of
libminer_b200.so; this file is stream / buffer plumbing only.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops
from . import _lib as L


class HostEvaluator:
    def __init__(self, model, wave: int = 65536, chunk: int = 32768, ks: Sequence[int] = (5, 10), transform: str = 'sigmoid',
                 math: Optional[int] = None, check_bounds: bool = True):
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

    def _proj_ws(self) -> torch.Tensor:
        n = L.load().miner_table_project_workspace_bytes(self.table.shape[0], self.model.poly_attn.context_codes.shape[1])
        if getattr(self, '_pws', None) is None or self._pws.numel() < n:
            self._pws = torch.empty(max(n, 1), dtype=torch.uint8, device=self.dev)
        return self._pws

    def _stage(self, host: Dict[str, torch.Tensor], a: int, b: int):
        """Issue the H2D copies of impressions [a, b) on the copy stream; returns device tensors + the event to wait on."""
        offs = host['offsets']
        c0, c1 = int(offs[a]), int(offs[b])
        with torch.cuda.stream(self.copy_stream):
            d = {'his_ids': host['his_ids'][a:b].to(self.dev, non_blocking=True),
                 'his_mask': host['his_mask'][a:b].to(self.dev, non_blocking=True),
                 'cand_ids': host['cand_ids'][c0:c1].to(self.dev, non_blocking=True),
                 'labels': host['labels'][c0:c1].to(self.dev, non_blocking=True),
                 'offsets': offs[a:b + 1].to(self.dev, non_blocking=True)}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        d['c0'] = c0
        return d, ev

    @torch.no_grad()
    def evaluate(self, host: Dict[str, torch.Tensor], want_scores: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """host: pinned CPU tensors his_ids (B,H), his_mask (B,H), cand_ids (T,), labels (T,), offsets (B+1,).
        Returns (partials (2*M,) float64 on the device: [sum, count] per metric of ops.metric_names(ks), scores (T,) or None)."""
        B = host['his_ids'].shape[0]
        Hh, Kk = host['his_ids'].shape[1], self.model.poly_attn.context_codes.shape[0]
        if self._math_arg is None:
            self.math = ops.default_eval_math(self.table, host['his_ids'].shape[1], self.model.poly_attn.context_codes.shape[0])
        w = self.model._weights(with_bf16=(self.math != L.MATH_FP32))
        if self.math == L.MATH_TABLE:
            # the table-level projections belong to the call: they are recomputed here, not carried over between calls
            self._proj = ops.table_project(self.table, w, weighted=self.model.score_type == 'weighted', out=getattr(self, '_proj', None),
                                           workspace=self._proj_ws())
        compute = torch.cuda.current_stream(self.dev)
        total = None
        scores_all = torch.empty(int(host['offsets'][-1]), dtype=torch.float32, device=self.dev) if want_scores else None
        bounds = list(range(0, B, self.wave)) + [B]
        nxt = self._stage(host, bounds[0], bounds[1]) if B > 0 else None
        for i in range(len(bounds) - 1):
            d, ev = nxt
            nxt = self._stage(host, bounds[i + 1], bounds[i + 2]) if i + 2 < len(bounds) else None
            compute.wait_event(ev)
            offs = d['offsets'] - d['c0']
            if self.math == L.MATH_TABLE:
                nb = d['his_ids'].shape[0]
                if getattr(self, '_tws', None) is None or self._tws.numel() < L.load().miner_score_table_workspace_bytes(nb, Hh, Kk):
                    self._tws = ops.score_table_workspace(max(nb, self.wave), Hh, Kk, self.dev)
                _, s = ops.score_table(self._proj, d['his_ids'], d['his_mask'], d['cand_ids'], self.model.score_type, cand_offsets=offs,
                                       workspace=self._tws)
            else:
                _, s = ops.score(self.table, d['his_ids'], d['his_mask'], d['cand_ids'], w, self.model.score_type, cand_offsets=offs,
                                 math=self.math, chunk=self.chunk)
            p, _ = ops.rank_metrics_raw(s, d['labels'], offs, self.transform, self.ks)
            total = p if total is None else total + p
            if want_scores:
                scores_all[d['c0']:d['c0'] + s.numel()] = s
            for t in d.values():                      # the copy stream must not recycle these buffers before the kernels are done
                if isinstance(t, torch.Tensor):
                    t.record_stream(compute)
        if total is None:
            total = torch.zeros(2 * (2 + 2 * len(self.ks)), dtype=torch.float64, device=self.dev)
        if self.check_bounds and getattr(self, '_tws', None) is not None:
            # ids outside the table raise what the reference's indexing raises (one 8-byte read at the end of the call)
            try:
                ops.check_oob(self._tws)
            finally:
                self._tws[:8].zero_()
        return total, scores_all
