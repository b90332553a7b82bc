"""CPU oracle for the MINER scoring path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, op for op and in the same order, what the reference
(MrRobot2211/miner, mounted at /root/reference while the repo was authored)
computes on the data-parallel scoring path.  It is a floating-point path, so the
restatement uses torch CPU fp32 ops (the same aten kernels the reference itself
dispatches to) for the model part and numpy float64 for the ranking metrics (the
reference's own choice, src/evaluation.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or as
the timed CPU baseline -- never on the product path.  ``miner_b200`` never imports
it (tests/test_boundary.py asserts that).

Pinning: the reference ships no tests, golden vectors or known-answer files
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports the unmodified reference classes
in the authoring container and writes ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against them.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MASK_FILL = 1e-30  # src/model/model.py:180 -- masked logits become 1e-30, NOT -inf


# --------------------------------------------------------------------------- #
# Step 1: embedding gather (NewsEncoder interface restated as a table lookup)
# --------------------------------------------------------------------------- #
def gather(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """``news_encoder(title_encoding=ids[:, None], ...) -> (N, embed_dim)``.

    Restates the call contract of NewsEncoder.forward (src/model/news_encoder.py:60-106)
    as used by Miner.forward (src/model/model.py:96-97,109-110) with the RoBERTa body
    replaced by ``table[id]``; ids ride in ``title[..., 0]``.
    """
    return table[ids.reshape(-1).long()].reshape(*ids.shape, table.shape[1])


# --------------------------------------------------------------------------- #
# pairwise cosine similarity  (src/utils.py:9-29)
# --------------------------------------------------------------------------- #
def pairwise_cosine_similarity(x: torch.Tensor, y: torch.Tensor, zero_diagonal: bool = False) -> torch.Tensor:
    x_norm = torch.linalg.norm(x, dim=2, keepdim=True)                       # utils.py:21
    y_norm = torch.linalg.norm(y, dim=2, keepdim=True)                       # utils.py:22
    distance = torch.matmul(torch.div(x, x_norm), torch.div(y, y_norm).permute(0, 2, 1))  # utils.py:23
    if zero_diagonal:                                                        # utils.py:24-27
        assert x.shape[1] == y.shape[1]
        eye = torch.eye(x.shape[1]).repeat(x.shape[0], 1, 1).bool()
        distance = distance.masked_fill(eye, 0)
    return distance


def category_bias(cat_emb: torch.Tensor, his_category: torch.Tensor, category: torch.Tensor) -> torch.Tensor:
    """Category-similarity bias ``(B, H, C)``  (src/model/model.py:113-120, eval mode: dropout is identity).

    ``cat_emb`` is ``category_embedding.weight`` (NC, Ec) whose padding row is zero, so a pad
    category yields 0/0 = NaN exactly as in the reference (SURVEY.md section 7 "NaN behaviour").
    """
    his = cat_emb[his_category.long()]
    cand = cat_emb[category.long()]
    return pairwise_cosine_similarity(his, cand)


# --------------------------------------------------------------------------- #
# Step 2: poly attention  (src/model/model.py:159-185)
# --------------------------------------------------------------------------- #
def poly_attention(embeddings: torch.Tensor, attn_mask: torch.Tensor, w_proj: torch.Tensor,
                   context_codes: torch.Tensor, bias: Optional[torch.Tensor] = None,
                   return_weights: bool = False):
    proj = torch.tanh(F.linear(embeddings, w_proj))                         # model.py:171
    if bias is None:
        weights = torch.matmul(proj, context_codes.T)                       # model.py:174
    else:
        b = bias.mean(dim=2).unsqueeze(dim=2)                               # model.py:176
        weights = torch.matmul(proj, context_codes.T) + b                   # model.py:177
    weights = weights.permute(0, 2, 1)                                      # model.py:178
    weights = weights.masked_fill(~attn_mask.unsqueeze(dim=1), MASK_FILL)   # model.py:180
    weights = F.softmax(weights, dim=2)                                     # model.py:181
    poly_repr = torch.matmul(weights, embeddings)                           # model.py:182
    if return_weights:
        return poly_repr, weights
    return poly_repr


# --------------------------------------------------------------------------- #
# Step 3+4: target-aware attention and the per-candidate score
# --------------------------------------------------------------------------- #
def target_aware_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                           w_target: torch.Tensor) -> torch.Tensor:
    """src/model/model.py:200-216 (exact erf gelu)."""
    proj = F.gelu(F.linear(query, w_target))                                # model.py:212
    weights = F.softmax(torch.matmul(key, proj.permute(0, 2, 1)), dim=2)    # model.py:213
    return torch.mul(weights, value).sum(dim=2)                             # model.py:214


def aggregate_scores(interests: torch.Tensor, candidate_repr: torch.Tensor, score_type: str,
                     w_target: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src/model/model.py:127-136."""
    matching = torch.matmul(candidate_repr, interests.permute(0, 2, 1))     # model.py:127
    if score_type == 'max':
        return matching.max(dim=2)[0]                                       # model.py:129
    if score_type == 'mean':
        return matching.mean(dim=2)                                         # model.py:131
    if score_type == 'weighted':
        return target_aware_attention(interests, candidate_repr, matching, w_target)  # model.py:133
    raise ValueError('Invalid method of aggregating matching score')        # model.py:136


def miner_forward(table: torch.Tensor, his_ids: torch.Tensor, his_mask: torch.Tensor, cand_ids: torch.Tensor,
                  w_proj: torch.Tensor, context_codes: torch.Tensor, w_target: Optional[torch.Tensor],
                  score_type: str = 'weighted', cat_emb: Optional[torch.Tensor] = None,
                  his_category: Optional[torch.Tensor] = None, category: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Miner.forward (src/model/model.py:61-138) with the encoder restated as ``table[id]``.

    Returns ``(multi_user_interest (B,K,D), matching_scores (B,C))`` (model.py:138).
    """
    table = table.float()
    candidate_repr = gather(table, cand_ids)                                # model.py:96-98
    history_repr = gather(table, his_ids)                                   # model.py:109-111
    bias = None
    if cat_emb is not None:                                                 # model.py:113-120
        bias = category_bias(cat_emb, his_category, category)
    interests = poly_attention(history_repr, his_mask.bool(), w_proj, context_codes, bias)  # model.py:122-124
    scores = aggregate_scores(interests, candidate_repr, score_type, w_target)
    return interests, scores


def miner_forward_csr(table, his_ids, his_mask, cand_ids_flat, cand_offsets, w_proj, context_codes, w_target,
                      score_type='weighted', chunk: int = 512):
    """Grouped eval layout: impression ``i`` owns candidates ``cand_offsets[i]:cand_offsets[i+1]``.

    Without category bias the score of a candidate depends only on its own impression
    (SURVEY.md section 8e), so this equals the reference's per-candidate C=1 rows
    (src/reader.py:376-379) up to fp32 summation order.  Interests are computed once per
    impression and candidates of equal count are batched together.
    """
    table = table.float()
    B = his_ids.shape[0]
    offs = np.asarray(cand_offsets, dtype=np.int64)
    counts = offs[1:] - offs[:-1]
    scores = torch.empty(int(offs[-1]), dtype=torch.float32)
    for c in np.unique(counts):
        if c == 0:
            continue
        rows = np.nonzero(counts == c)[0]
        for s in range(0, len(rows), chunk):
            r = rows[s:s + chunk]
            idx = torch.from_numpy(offs[r][:, None] + np.arange(c)[None, :])
            cid = cand_ids_flat[idx]
            rt = torch.from_numpy(r)
            _, sc = miner_forward(table, his_ids[rt], his_mask[rt], cid, w_proj, context_codes, w_target, score_type)
            scores[idx.reshape(-1)] = sc.reshape(-1)
    return scores


# --------------------------------------------------------------------------- #
# Losses (train variant + eval loss)
# --------------------------------------------------------------------------- #
def loss_compute(poly_attn: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Loss.compute (src/loss.py:27-44) with nn.CrossEntropyLoss(reduction='mean') (src/trainer.py:303)."""
    disagreement = pairwise_cosine_similarity(poly_attn, poly_attn, zero_diagonal=True).mean()  # loss.py:39
    targets = labels.argmax(dim=1)                                                              # loss.py:40
    rank_loss = F.cross_entropy(logits, targets, reduction='mean')                              # loss.py:41
    return disagreement + rank_loss                                                             # loss.py:42


def loss_compute_eval(poly_attn: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor) -> float:
    """Loss.compute_eval_loss (src/loss.py:68-85)."""
    disagreement = pairwise_cosine_similarity(poly_attn, poly_attn, zero_diagonal=True).mean()  # loss.py:81
    rank_loss = -(F.logsigmoid(logits) * labels).sum()                                          # loss.py:82
    return (disagreement + rank_loss).item()                                                    # loss.py:83-85


# --------------------------------------------------------------------------- #
# Ranking metrics  (src/evaluation.py)
# --------------------------------------------------------------------------- #
def _desc_order_numpy_like(y_score: np.ndarray) -> np.ndarray:
    """``np.argsort(y_score)[::-1]`` (evaluation.py:188,208) with a *defined* tie rule.

    numpy's default sort is not stable, so the reference's order among equal scores is
    implementation-defined.  The oracle (and the CUDA kernel) fix it as "stable ascending
    sort, reversed" = among ties the LATER index ranks first, which is what numpy produces
    for the short arrays (n <= 16, insertion sort) that dominate MIND impressions.
    """
    return np.argsort(y_score, kind='stable')[::-1]


def mrr_score(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """compute_mrr_score (evaluation.py:177-192)."""
    rank = _desc_order_numpy_like(np.asarray(y_score))
    y = np.take(np.asarray(y_true), rank)
    rr = y / (np.arange(len(y)) + 1)
    with np.errstate(invalid='ignore', divide='ignore'):
        return float(np.sum(rr) / np.sum(y))


def dcg_score(y_true: np.ndarray, y_score: np.ndarray, k: int) -> float:
    """compute_dcg_score (evaluation.py:195-213)."""
    y_true = np.asarray(y_true)
    k = min(np.shape(y_true)[-1], k)
    order = _desc_order_numpy_like(np.asarray(y_score))
    y = np.take(y_true, order[:k])
    gains = 2 ** y - 1
    discounts = np.log2(np.arange(len(y)) + 2)
    return float(np.sum(gains / discounts))


def ndcg_score(y_true: np.ndarray, y_score: np.ndarray, k: int) -> float:
    """compute_ndcg_score (evaluation.py:216-231)."""
    best = dcg_score(y_true, y_true, k)
    actual = dcg_score(y_true, y_score, k)
    with np.errstate(invalid='ignore', divide='ignore'):
        return float(np.float64(actual) / np.float64(best))


def hit_score(y_true: Sequence, y_score: Sequence, k: int) -> int:
    """is_hit (evaluation.py:245-249): python's stable ``sorted(..., reverse=True)`` keeps the
    EARLIER index first among ties."""
    ordered = sorted(zip(y_score, y_true), key=lambda x: x[0], reverse=True)
    return int(sum(label for _, label in ordered[:k]) > 0)


def auc_score(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """sklearn.metrics.roc_auc_score as called at evaluation.py:54,57, restated as the tie-aware
    Mann-Whitney statistic ``(#[s_p > s_n] + 0.5 #[s_p == s_n]) / (P N)`` (scikit-learn 1.4.1.post1
    is the reference's pin, environment.yml:308; not vendored).  One-class input: sklearn raises;
    the reference never feeds one (src/reader.py:374) -- the oracle returns NaN.
    """
    y = np.asarray(y_true).astype(bool)
    s = np.asarray(y_score, dtype=np.float64)
    pos, neg = s[y], s[~y]
    if len(pos) == 0 or len(neg) == 0:
        return float('nan')
    order = np.argsort(s, kind='stable')
    ranks = np.empty(len(s), dtype=np.float64)
    ss = s[order]
    i = 0
    while i < len(ss):                       # average ranks over ties
        j = i
        while j + 1 < len(ss) and ss[j + 1] == ss[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    u = ranks[y].sum() - len(pos) * (len(pos) + 1) / 2.0
    return float(u / (len(pos) * len(neg)))


def sigmoid_probs(logits: torch.Tensor) -> List[float]:
    """SlowEvaluator.eval_batch (evaluation.py:165-168): fp32 sigmoid, then python floats."""
    return torch.sigmoid(logits.float()).reshape(-1).tolist()


def group_by_impression(values: Sequence, impression_ids: Sequence[int]) -> List[list]:
    """SlowEvaluator._convert_targets/_convert_pred (evaluation.py:118-131,135-149): concatenate per
    impression id in arrival order, then sort groups by id."""
    groups: Dict[int, list] = {}
    for v, i in zip(values, impression_ids):
        groups.setdefault(int(i), []).append(v)
    return [g for _, g in sorted(groups.items())]


def compute_scores(targets: List[list], probs: List[list], metrics: Sequence[str]) -> Dict[str, float]:
    """BaseEvaluator.compute_scores (evaluation.py:36-84) over already grouped lists."""
    assert len(targets) == len(probs)
    out: Dict[str, float] = {}
    for metric in metrics:
        if metric == 'auc':                                                  # evaluation.py:53-55
            flat_t = [t for g in targets for t in g]
            flat_p = [p for g in probs for p in g]
            out['auc'] = auc_score(np.array(flat_t), np.array(flat_p))
        elif metric == 'group_auc':                                          # evaluation.py:56-59
            out['group_auc'] = float(np.nanmean([auc_score(np.array(t), np.array(p)) for t, p in zip(targets, probs)]))
        elif metric == 'mrr':                                                # evaluation.py:62-65
            out['mrr'] = float(np.nanmean([mrr_score(np.array(t), np.array(p)) for t, p in zip(targets, probs)]))
        elif metric.startswith('ndcg'):                                      # evaluation.py:68-72
            k = int(metric.split('@')[1])
            out[f'ndcg@{k}'] = float(np.nanmean([ndcg_score(np.array(t), np.array(p), k) for t, p in zip(targets, probs)]))
        elif metric.startswith('hit'):                                       # evaluation.py:76-80
            k = int(metric.split('@')[1])
            out[f'hit@{k}'] = float(np.nanmean([hit_score(t, p, k) for t, p in zip(targets, probs)]))
    return out


def per_impression_metrics(labels_flat: np.ndarray, probs_flat: np.ndarray, offsets: np.ndarray,
                           ks: Sequence[int] = (5, 10)) -> Dict[str, np.ndarray]:
    """Per-impression values of every metric, CSR layout (used to check the CUDA kernel row by row)."""
    n = len(offsets) - 1
    out = {'group_auc': np.empty(n), 'mrr': np.empty(n)}
    for k in ks:
        out[f'ndcg@{k}'] = np.empty(n)
        out[f'hit@{k}'] = np.empty(n)
    for i in range(n):
        a, b = int(offsets[i]), int(offsets[i + 1])
        t, p = np.asarray(labels_flat[a:b]), np.asarray(probs_flat[a:b], dtype=np.float64)
        out['group_auc'][i] = auc_score(t, p)
        out['mrr'][i] = mrr_score(t, p)
        for k in ks:
            out[f'ndcg@{k}'][i] = ndcg_score(t, p, k)
            out[f'hit@{k}'][i] = hit_score(list(t), list(p), k)
    return out


def fast_eval_probs(logits: torch.Tensor) -> List[list]:
    """FastEvaluator.eval_batch (evaluation.py:98-110): softmax over the npratio+1 columns."""
    return F.softmax(logits.float(), dim=1).tolist()
