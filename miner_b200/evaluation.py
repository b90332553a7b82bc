"""Mirror of the reference's ``src/evaluation.py`` on the segmented ranking-metric kernel.

Same evaluator classes and call sequence as the reference trainer uses (src/trainer.py:265-295):

    evaluator = SlowEvaluator(dataset)            # or FastEvaluator(dataset)
    evaluator.eval_batch(logits, impression_ids)  # once per batch
    scores = evaluator.compute_scores(metrics, save_result, path)

but logits stay on the GPU (the reference does ``sigmoid(logits).tolist()`` per batch, a device->host sync each
time, then per-impression Python loops over numpy/sklearn -- evaluation.py:57-80,165-170).  At ``compute_scores``
the candidates are grouped by impression id into CSR form and one kernel launch produces every per-impression
metric and their NaN-skipping sums (np.nanmean).  Tie rule among equal scores: see oracle/miner_oracle.py.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import ops

_KNOWN = ('auc', 'group_auc', 'mrr')


def _parse_metrics(metrics: Sequence[str]):
    ks: List[int] = []
    for m in metrics:
        if m.startswith('ndcg@') or m.startswith('hit@'):
            k = int(m.split('@')[1])
            if k not in ks:
                ks.append(k)
    return ks


def global_auc(scores: Tensor, labels: Tensor, offsets: Optional[Tensor] = None, transform: str = 'sigmoid', group=None,
               _kernels=None) -> float:
    """``roc_auc_score`` over ALL flattened candidates (reference evaluation.py:53-55) = the tie-aware Mann-Whitney statistic,
    exact: probabilities -> order-preserving integer keys (``miner_auc_split``), radix sort of the positive keys
    (``miner_sort_u32``), every negative binary-searches them (``miner_auc_count``).  With several ranks (``torch.distributed``
    initialised) the positive keys of all ranks are all-gathered, every rank counts its own negatives against the global
    positives and ``[2U, N]`` are all-reduced as int64: the value is the single-process one, bit for bit.

    ``scores`` are the raw logits (``transform`` as in ``rank_metrics``: 'sigmoid' SlowEvaluator, 'softmax' FastEvaluator, 'none').
    ``_kernels`` (tests): ``(split, sort, count)`` stand-ins for the three device calls."""
    import torch.distributed as dist
    split, sort, count = _kernels if _kernels is not None else (ops.auc_split, ops.sort_u32, ops.auc_count)
    pos, neg = split(scores, labels, offsets, transform)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    n_neg = neg.numel()
    if multi:
        world = dist.get_world_size(group)
        sizes = torch.zeros(world, dtype=torch.int64, device=pos.device)
        sizes[dist.get_rank(group)] = pos.numel()
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
        sizes = sizes.tolist()
        cap = max(max(sizes), 1)
        mine = torch.zeros(cap, dtype=pos.dtype, device=pos.device)
        mine[:pos.numel()] = pos
        gathered = [torch.empty(cap, dtype=pos.dtype, device=pos.device) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        pos = torch.cat([g[:n] for g, n in zip(gathered, sizes)])
    n_pos = pos.numel()
    u2 = count(sort(pos.contiguous()), neg) if (n_pos and n_neg) else 0
    if multi:
        t = torch.tensor([u2, n_neg], dtype=torch.int64, device=pos.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        u2, n_neg = (int(v) for v in t.tolist())
    if n_pos == 0 or n_neg == 0:
        return float('nan')
    return u2 / (2.0 * n_pos * n_neg)


class BaseEvaluator:
    transform = 'none'

    def __init__(self, dataset):
        self.dataset = dataset
        self._logits: List[Tensor] = []
        self._ids: List[Tensor] = []
        self._convert_targets()

    def _convert_targets(self):
        raise NotImplementedError

    def _csr(self, device):
        raise NotImplementedError

    def compute_scores(self, metrics: List[str], save_result: bool = False, path: Optional[str] = None) -> Dict[str, float]:
        ks = _parse_metrics(metrics)
        ops.check_oob_all()            # an id outside the news table raises IndexError, as the reference's table indexing does
        scores_flat, labels_flat, offsets = self._csr()
        partials, per = ops.rank_metrics_raw(scores_flat, labels_flat, offsets, self.transform, ks, per_impression=save_result)
        p = partials.cpu().view(-1, 2)
        names = ops.metric_names(ks)
        table = {n: (float(p[i, 0] / p[i, 1]) if p[i, 1] > 0 else float('nan')) for i, n in enumerate(names)}
        out: Dict[str, float] = {}
        for m in metrics:
            if m == 'auc':
                out['auc'] = global_auc(scores_flat, labels_flat, offsets, self.transform)
            elif m in table:
                out[m] = table[m]
        if save_result and path is not None:               # evaluation.py:60-61,66-67,73-74,81-82
            per_cpu = per.cpu()
            fname = {'group_auc': 'group_auc.txt', 'mrr': 'mrr.txt'}
            for i, n in enumerate(names):
                if n not in metrics:
                    continue
                f = fname.get(n, n.replace('@', '') + '.txt')
                with open(os.path.join(path, f), mode='w', encoding='utf-8') as fh:
                    for v in per_cpu[:, i].tolist():
                        fh.write(str(v))
                        fh.write('\n')
        return out


class SlowEvaluator(BaseEvaluator):
    """One sample per candidate, grouped by impression id (reference evaluation.py:113-175)."""
    transform = 'sigmoid'

    def _convert_targets(self):
        ids, labels = [], []
        for sample in self.dataset.samples:                                   # evaluation.py:118-121
            imp = sample.impression
            ids.extend([int(imp.impression_id)] * len(imp.label))
            labels.extend(int(v) for v in imp.label)
        self._target_ids = torch.tensor(ids, dtype=torch.int64)
        self._target_labels = torch.tensor(labels, dtype=torch.int8)

    def eval_batch(self, logits: Tensor, impression_ids: Tensor):
        r"""logits ``(batch_size, 1)``, impression_ids ``(batch_size,)`` -- both stay on the device."""
        lg = logits.detach().reshape(logits.shape[0], -1).float()
        self._logits.append(lg.reshape(-1))
        self._ids.append(impression_ids.detach().reshape(-1).repeat_interleave(lg.shape[1]))

    def _csr(self):
        logits = torch.cat(self._logits)
        ids = torch.cat(self._ids).to(torch.int64)
        dev = logits.device
        # predictions: stable sort by impression id keeps arrival order inside a group (evaluation.py:135-149)
        order = torch.sort(ids, stable=True)[1]
        uniq, counts = torch.unique_consecutive(ids[order], return_counts=True)
        offsets = torch.zeros(uniq.numel() + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(counts, 0)
        # targets: same grouping of the dataset's labels (evaluation.py:118-131)
        tid = self._target_ids.to(dev)
        torder = torch.sort(tid, stable=True)[1]
        tuniq, tcounts = torch.unique_consecutive(tid[torder], return_counts=True)
        assert torch.equal(tuniq, uniq) and torch.equal(tcounts, counts), 'predictions and targets do not cover the same impressions'
        return logits[order], self._target_labels.to(dev)[torder], offsets

    def save_predictions(self, path: str):
        import pickle as pk
        probs = torch.sigmoid(torch.cat(self._logits)).tolist()
        pk.dump({'pred': probs, 'impression_id': torch.cat(self._ids).tolist()}, open(os.path.join(path, 'preds.pkl'), 'wb'))


class FastEvaluator(BaseEvaluator):
    """One sample per impression with npratio+1 candidates; softmax over them (reference evaluation.py:87-110)."""
    transform = 'softmax'

    def _convert_targets(self):
        self._targets = [list(sample.impression.label) for sample in self.dataset.samples]

    def eval_batch(self, logits: Tensor, impression_ids: Tensor = None):
        self._logits.append(logits.detach().float())

    def _csr(self):
        logits = torch.cat(self._logits, dim=0)
        dev = logits.device
        counts = torch.tensor([len(t) for t in self._targets], dtype=torch.int64)
        assert logits.shape[0] == len(self._targets) and bool((counts == logits.shape[1]).all())
        offsets = torch.zeros(len(self._targets) + 1, dtype=torch.int64)
        offsets[1:] = torch.cumsum(counts, 0)
        labels = torch.tensor([v for t in self._targets for v in t], dtype=torch.int8)
        return logits.reshape(-1), labels.to(dev), offsets.to(dev)
