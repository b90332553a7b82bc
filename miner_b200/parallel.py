"""Data-parallel sharding of impressions.

Impressions are independent units: an impression's scores depend only on its own ids plus the replicated table and
weights (reference model.py:61-138 mixes nothing across rows), so the path shards with NO data-path collective.
Ranks take contiguous impression ranges balanced by cumulative candidate count; the only exchange is one
all-reduce(SUM) of the [sum, count] metric partials at the end of the evaluation (exact for every np.nanmean
metric of evaluation.py:59-80).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(offsets: torch.Tensor, world_size: int, his_len: int = 0, align: int = 4) -> List[Tuple[int, int]]:
    """Split impressions ``[0, B)`` into ``world_size`` contiguous ranges of near-equal cost.

    cost(impression) = his_len + n_candidates (rows gathered), taken from the CSR ``offsets`` (B+1,).
    Returns ``[(start, end), ...]``; ranges are contiguous, cover every impression once and may be empty.
    Interior boundaries are multiples of ``align`` (4 = the most impressions the table-level kernel puts in one tile): an
    impression then shares its tile with the same neighbours whatever the number of ranks, so every score -- and with it every
    metric -- is bit-identical at 1, 2, 4 and 8 ranks.
    """
    offs = offsets.detach().to('cpu', torch.int64)
    B = offs.numel() - 1
    if B <= 0:
        return [(0, 0)] * world_size
    cost = (offs[1:] - offs[:-1]) + int(his_len)
    csum = torch.cumsum(cost, 0)
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        idx = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.int64), right=False))
        idx = (idx + align // 2) // align * align if align > 1 else idx
        bounds.append(min(max(idx, bounds[-1]), B))
    bounds.append(B)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def local_shard(offsets: torch.Tensor, rank: int, world_size: int, his_len: int = 0):
    """``(start, end, local_offsets)`` for this rank; ``local_offsets`` is rebased to start at 0."""
    s, e = shard_bounds(offsets, world_size, his_len)[rank]
    loc = offsets[s:e + 1] - offsets[s]
    return s, e, loc


def allreduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the ``[sum, count]`` metric partials over ranks (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def allreduce_gradients(params, group=None) -> None:
    """Train variant: average the gradients of ``params`` over ranks with ONE flat all-reduce (the path has ~0.75 M weights:
    a single bucket, latency-bound).  With equal local batches the result is the gradient of the loss over the GLOBAL batch,
    which is what the reference's single-process ``CrossEntropyLoss(reduction='mean')`` / ``.mean()`` compute (loss.py:39-41)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


class FlatGradients:
    """The gradients of ``params`` as views of ONE flat fp32 buffer, so that the data-parallel exchange of a train step is a
    single in-place all-reduce (no concatenate / split copies around a 3 MB collective).  ``zero()`` replaces
    ``optimizer.zero_grad()`` (which would drop the views); autograd accumulates into the views in place."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, group=None) -> None:
        """Average over ranks: with equal local batches this is the gradient of the loss over the GLOBAL batch (loss.py:39-41)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if dist.get_backend(group) == 'nccl':
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:                                        # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat /= dist.get_world_size(group)


def finalize_metrics(partials: torch.Tensor, names: Sequence[str]) -> Dict[str, float]:
    p = partials.detach().to('cpu', torch.float64).view(-1, 2)
    return {n: (float(p[i, 0] / p[i, 1]) if p[i, 1] > 0 else float('nan')) for i, n in enumerate(names)}


def bind_to_gpu_cpus(gpu_index: int):
    """Restrict this process to the CPUs NVML reports as local to ``gpu_index`` (its NUMA node), intersected with the CPUs the process
    may already use.  One process per GPU streams pinned host buffers to its GPU: allocated after this call they are first touched --
    and therefore placed -- on the GPU's own memory node instead of wherever the launcher started the process.  Returns the previous
    affinity set (pass it to ``os.sched_setaffinity(0, ...)`` to undo) or ``None`` when nothing was changed."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return allowed
    except Exception:
        return None
