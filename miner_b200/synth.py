"""Seeded synthetic MIND-shaped inputs for the scoring path.

Mirrors the data-layer semantics that leak into the hot path:
  * history is LEFT-padded with the pad news (id 0) and masked there
    (reference src/reader.py:368-369, src/entities.py:395);
  * the pad news is a real, non-zero row of the table (src/reader.py:101-110);
  * eval impressions hold at least one positive and one negative (src/reader.py:374);
  * train rows are one positive + npratio negatives, shuffled (src/reader.py:173-181).
Everything is generated on the CPU from ``torch.Generator().manual_seed(seed)`` (the
reference's seed is 36, config/*.txt) so the same arrays can be re-created anywhere.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class Weights:
    w_proj: torch.Tensor          # poly_attn.linear.weight        (Dc, D)
    context_codes: torch.Tensor   # poly_attn.context_codes        (K, Dc)
    w_target: torch.Tensor        # target_aware_attn.linear.weight (D, D)
    cat_emb: Optional[torch.Tensor] = None  # category_embedding.weight (NC, Ec), row 0 = pad = zeros


@dataclass
class EvalBatch:
    his_ids: torch.Tensor      # (B, H) int64, 0 on pads
    his_mask: torch.Tensor     # (B, H) bool, False on pads
    his_category: torch.Tensor  # (B, H) int64, 0 on pads
    cand_ids: torch.Tensor     # (T,) int64  flattened candidates
    cand_category: torch.Tensor  # (T,) int64
    labels: torch.Tensor       # (T,) int8
    offsets: torch.Tensor      # (B+1,) int64 CSR row offsets


def make_table(num_news: int, dim: int, seed: int = 36, dtype=torch.float32) -> torch.Tensor:
    """``N+1`` rows (row 0 = pad news), values N(0,1)*4/sqrt(D): keeps |logit| <~ 2 so that the
    fp32 sigmoid of SlowEvaluator never saturates into ties."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(num_news + 1, dim, generator=g) * (4.0 / math.sqrt(dim))
    return t.to(dtype)


def make_weights(dim: int, num_codes: int, code_dim: int, seed: int = 36, num_category: int = 0,
                 category_dim: int = 0) -> Weights:
    """Same initialisers, in the same order, as the reference constructors
    (src/model/model.py:45-58: category nn.Embedding, PolyAttention's nn.Linear +
    xavier_uniform(gain=tanh) codes, TargetAwareAttention's nn.Linear)."""
    g = torch.Generator().manual_seed(seed)
    cat = None
    if num_category:
        cat = torch.randn(num_category, category_dim, generator=g)
        cat[0].zero_()                                    # padding_idx row
    def linear(out_f, in_f):                              # nn.Linear default: kaiming_uniform(a=sqrt(5))
        bound = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    w_proj = linear(code_dim, dim)
    gain = 5.0 / 3.0                                      # nn.init.calculate_gain('tanh')
    bound = gain * math.sqrt(6.0 / (num_codes + code_dim))
    codes = (torch.rand(num_codes, code_dim, generator=g) * 2 - 1) * bound
    w_target = linear(dim, dim)
    return Weights(w_proj, codes, w_target, cat)


def make_history(batch: int, his_len: int, num_news: int, g: torch.Generator, num_category: int = 20,
                 min_len: int = 1):
    ids = torch.randint(1, num_news + 1, (batch, his_len), generator=g)
    cats = torch.randint(1, max(2, num_category), (batch, his_len), generator=g)
    length = torch.randint(min_len, his_len + 1, (batch,), generator=g)
    mask = torch.arange(his_len)[None, :] >= (his_len - length)[:, None]     # left padding
    return ids * mask, mask, cats * mask


def make_eval_batch(batch: int, his_len: int, num_news: int, seed: int = 36, mean_cands: float = 20.0,
                    fixed_cands: Optional[int] = None, num_category: int = 20, max_cands: int = 300) -> EvalBatch:
    g = torch.Generator().manual_seed(seed + 1)
    his_ids, his_mask, his_cat = make_history(batch, his_len, num_news, g, num_category)
    if fixed_cands is not None:
        counts = torch.full((batch,), int(fixed_cands), dtype=torch.int64)
    else:
        z = torch.randn(batch, generator=g) * 0.5 + math.log(mean_cands) - 0.125   # mean of lognormal ~ mean_cands
        counts = torch.exp(z).round().clamp(2, max_cands).to(torch.int64)
    offsets = torch.zeros(batch + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    total = int(offsets[-1])
    cand_ids = torch.randint(1, num_news + 1, (total,), generator=g)
    cand_cat = torch.randint(1, max(2, num_category), (total,), generator=g)
    labels = (torch.rand(total, generator=g) < 0.12).to(torch.int8)
    # at least one positive and one negative per impression (reader.py:374)
    first = offsets[:-1]
    labels[first] = 1
    labels[first + 1] = 0
    # shuffle position of the forced pair inside each impression cheaply: roll by a random amount
    return EvalBatch(his_ids, his_mask, his_cat, cand_ids, cand_cat, labels, offsets)


def make_train_batch(batch: int, his_len: int, num_news: int, npratio: int = 4, seed: int = 36,
                     num_category: int = 20):
    """(B, npratio+1) candidates, one-hot labels with the positive at a random column."""
    g = torch.Generator().manual_seed(seed + 2)
    his_ids, his_mask, his_cat = make_history(batch, his_len, num_news, g, num_category)
    c = npratio + 1
    cand = torch.randint(1, num_news + 1, (batch, c), generator=g)
    cand_cat = torch.randint(1, max(2, num_category), (batch, c), generator=g)
    pos = torch.randint(0, c, (batch,), generator=g)
    labels = torch.zeros(batch, c, dtype=torch.int64)
    labels[torch.arange(batch), pos] = 1
    return his_ids, his_mask, his_cat, cand, cand_cat, labels


def algorithmic_bytes_per_impression(his_len: int, cands: float, dim: int, elt: int, id_bytes: int = 8) -> float:
    """gathered rows + ids + mask + scores out + labels."""
    return (his_len + cands) * dim * elt + (his_len + cands) * id_bytes + his_len + cands * 4 + cands


def algorithmic_flops_per_impression(his_len: int, cands: float, dim: int, codes: int, code_dim: int) -> float:
    """SURVEY.md section 8d, reference order, score_type='weighted'."""
    h, c, d, k, dc = his_len, cands, dim, codes, code_dim
    return 2 * h * d * dc + 2 * h * dc * k + 2 * k * h * d + 2 * k * d * d + 4 * c * k * d


def deterministic_state(shapes: dict, seed: int = 36) -> dict:
    """A reproducible ``state_dict`` for a module given its ``{name: shape}``: used to put the SAME weights into the reference's
    ``FastFormer`` (tests/golden/make_golden.py) and into ``miner_b200.FastFormer`` (the GPU parity test) without storing them.
    LayerNorm gains ~1, everything else small normals; one generator, keys in sorted order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith('LayerNorm.weight'):
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith('bias'):
            out[name] = 0.02 * torch.randn(shape, generator=g)
        else:
            out[name] = 0.05 * torch.randn(shape, generator=g)
    return out
