"""Fastformer user-encoder variant (BASELINE configs[4]) on the same gather + score kernels, and
the news-table builder that makes the embedding table real.

``FastFormer`` mirrors the reference's ``src/model/model.py:223-341``: same constructor ``(news_encoder, score_type,
dropout)``, same ``forward`` keywords, same return value (the reference returns the ``(batch_size, num_candidates)`` score
matrix only, ``model.py:341``), and the same parameter names (``fast_attn.encoders.<i>.attention.self.query.weight`` ...
``fast_attn.poolers.0.att_fc2.bias``), so a reference checkpoint's ``state_dict`` loads.  The hidden size is 256 as in the
reference (``model.py:245-266`` hard-codes the config); a news encoder of another width is rejected.

What runs where:
  * candidate / history news vectors: with a ``TableNewsEncoder`` the bit-exact ``miner_gather`` kernel (otherwise the encoder
    is called exactly as the reference calls it, ``model.py:312-329``);
  * the additive-attention encoder body (two layers of 256-d dense blocks) stays PyTorch -- it is a small dense MLP stack,
    not the path this repository accelerates -- written here from the Fastformer equations, independent of HF ``transformers``;
  * the click score ``candidate . user`` (``model.py:335``): the dot-score kernel (``miner_target_score_fwd`` with one
    interest vector).

``build_news_table`` replaces the reference's per-batch encoding of every title (``model.py:96-111``,
``news_encoder.py:60-106``) by ONE pass of the news encoder over the news set, written into an ``(N + 1, D)`` table whose row
``id`` is the vector of news ``id`` (row 0 = the pad news, ``reader.py:101-110``).
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from . import _lib as L
from .model import TableNewsEncoder

HIDDEN, HEADS, LAYERS, MAX_POS, INIT_STD, LN_EPS = 256, 16, 2, 256, 0.02, 1e-12   # model.py:245-266
HIDDEN_DROPOUT = 0.2


def _init(module: nn.Module) -> None:
    # model.py:497-508: normal(0, 0.02) weights, zero biases, unit LayerNorm
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=INIT_STD)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


class _AdditiveSelfAttention(nn.Module):
    """Fastformer additive attention (model.py:374-455): a global query pooled with per-head additive scores, mixed into the
    keys, a global key pooled the same way, multiplied into the queries, transformed, plus the query residual."""

    def __init__(self):
        super().__init__()
        self.head_dim = HIDDEN // HEADS
        self.query = nn.Linear(HIDDEN, HIDDEN)
        self.query_att = nn.Linear(HIDDEN, HEADS)
        self.key = nn.Linear(HIDDEN, HIDDEN)
        self.key_att = nn.Linear(HIDDEN, HEADS)
        self.transform = nn.Linear(HIDDEN, HIDDEN)

    def forward(self, x: Tensor, additive_mask: Tensor) -> Tensor:
        B, S, _ = x.shape
        scale = self.head_dim ** 0.5
        q = self.query(x)                                                     # (B, S, 256)
        k = self.key(x)
        a = F.softmax(self.query_att(q).transpose(1, 2) / scale + additive_mask, dim=-1)          # (B, heads, S)
        qh = q.view(B, S, HEADS, self.head_dim).permute(0, 2, 1, 3)           # (B, heads, S, hd)
        gq = torch.matmul(a.unsqueeze(2), qh).transpose(1, 2).reshape(B, 1, HIDDEN)               # global query
        p = k * gq                                                            # (B, S, 256)
        b = F.softmax((self.key_att(p) / scale).transpose(1, 2) + additive_mask, dim=-1)          # (B, heads, S)
        ph = p.view(B, S, HEADS, self.head_dim).permute(0, 2, 1, 3)
        gk = torch.matmul(b.unsqueeze(2), ph)                                 # (B, heads, 1, hd) global key
        v = (gk * qh).transpose(1, 2).reshape(B, S, HIDDEN)
        return self.transform(v) + q


class _ResidualDense(nn.Module):
    """``LayerNorm(dropout(dense(h)) + residual)`` with the parameter names of HF's BertSelfOutput / BertOutput."""

    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HIDDEN, HIDDEN)
        self.LayerNorm = nn.LayerNorm(HIDDEN, eps=LN_EPS)
        self.dropout = nn.Dropout(HIDDEN_DROPOUT)

    def forward(self, h: Tensor, residual: Tensor) -> Tensor:
        return self.LayerNorm(self.dropout(self.dense(h)) + residual)


class _GeluDense(nn.Module):
    """``gelu(dense(h))`` (HF BertIntermediate; intermediate size 256)."""

    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HIDDEN, HIDDEN)

    def forward(self, h: Tensor) -> Tensor:
        return F.gelu(self.dense(h))


class _Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.self = _AdditiveSelfAttention()
        self.output = _ResidualDense()

    def forward(self, x: Tensor, additive_mask: Tensor) -> Tensor:
        return self.output(self.self(x, additive_mask), x)


class _Layer(nn.Module):
    def __init__(self):
        super().__init__()
        self.attention = _Attention()
        self.intermediate = _GeluDense()
        self.output = _ResidualDense()

    def forward(self, x: Tensor, additive_mask: Tensor) -> Tensor:
        a = self.attention(x, additive_mask)
        return self.output(self.intermediate(a), a)


class _AttentionPooling(nn.Module):
    """model.py:345-371: ``alpha = exp(fc2(tanh(fc1 x))) * mask``, normalised with +1e-8, weighted sum over the sequence."""

    def __init__(self):
        super().__init__()
        self.att_fc1 = nn.Linear(HIDDEN, HIDDEN)
        self.att_fc2 = nn.Linear(HIDDEN, 1)

    def forward(self, x: Tensor, attn_mask: Optional[Tensor] = None) -> Tensor:
        alpha = torch.exp(self.att_fc2(torch.tanh(self.att_fc1(x))))
        if attn_mask is not None:
            alpha = alpha * attn_mask.unsqueeze(2)
        alpha = alpha / (alpha.sum(dim=1, keepdim=True) + 1e-8)
        return torch.bmm(x.transpose(1, 2), alpha).squeeze(2)


class FastformerEncoder(nn.Module):
    """model.py:482-545: position embeddings + LayerNorm, two Fastformer layers, attention pooling to one user vector."""

    def __init__(self):
        super().__init__()
        self.encoders = nn.ModuleList([_Layer() for _ in range(LAYERS)])
        self.position_embeddings = nn.Embedding(MAX_POS, HIDDEN)
        self.LayerNorm = nn.LayerNorm(HIDDEN, eps=LN_EPS)
        self.dropout = nn.Dropout(HIDDEN_DROPOUT)
        self.poolers = nn.ModuleList([_AttentionPooling()])
        self.apply(_init)

    def forward(self, input_embs: Tensor, attention_mask: Tensor, pooler_index: int = 0) -> Tensor:
        B, S, _ = input_embs.shape
        mask_f = attention_mask.to(self.LayerNorm.weight.dtype)
        additive = ((1.0 - mask_f) * -10000.0).unsqueeze(1)                   # (B, 1, S), model.py:521
        h = input_embs + self.position_embeddings.weight[:S].unsqueeze(0)
        h = self.dropout(self.LayerNorm(h))
        for layer in self.encoders:
            h = layer(h, additive)
        return self.poolers[pooler_index](h, attention_mask)


class FastFormer(nn.Module):
    r"""Fastformer user encoder + dot-product click score (reference model.py:223-341) on the gather / score kernels."""

    def __init__(self, news_encoder: nn.Module, score_type: str, dropout: float):
        super().__init__()
        self.news_encoder = news_encoder
        self.news_embed_dim = self.news_encoder.embed_dim
        if self.news_embed_dim != HIDDEN:
            raise ValueError(f'FastFormer: the user encoder is {HIDDEN}-d (reference model.py:251); the news encoder is {self.news_embed_dim}-d')
        self.fast_attn = FastformerEncoder()
        self.score_type = score_type                    # stored and unused, as in the reference (model.py:270,337-339)
        self.dropout = nn.Dropout(dropout)

    def forward(self, title: Tensor, title_mask: Tensor, his_title: Tensor, his_title_mask: Tensor,
                his_mask: Tensor, sapo: Union[Tensor, None] = None, sapo_mask: Union[Tensor, None] = None,
                his_sapo: Union[Tensor, None] = None, his_sapo_mask: Union[Tensor, None] = None,
                category: Union[Tensor, None] = None, his_category: Union[Tensor, None] = None) -> Tensor:
        r"""Returns ``matching_scores`` ``(batch_size, num_candidates)`` (reference model.py:341)."""
        batch_size, num_candidates = title.shape[0], title.shape[1]
        his_length = his_title.shape[1]
        if isinstance(self.news_encoder, TableNewsEncoder):
            table = self.news_encoder.table
            candidate_repr = ops.gather(table, title.reshape(batch_size, num_candidates, -1)[..., 0])      # model.py:312-314
            history_repr = ops.gather(table, his_title.reshape(batch_size, his_length, -1)[..., 0])        # model.py:326-328
        else:
            candidate_repr = self.news_encoder(title_encoding=title.view(batch_size * num_candidates, -1),
                                               title_attn_mask=title_mask.view(batch_size * num_candidates, -1),
                                               sapo_encoding=sapo.view(batch_size * num_candidates, -1),
                                               sapo_attn_mask=sapo_mask.view(batch_size * num_candidates, -1)).view(batch_size, num_candidates, -1)
            history_repr = self.news_encoder(title_encoding=his_title.view(batch_size * his_length, -1),
                                             title_attn_mask=his_title_mask.view(batch_size * his_length, -1),
                                             sapo_encoding=his_sapo.view(batch_size * his_length, -1),
                                             sapo_attn_mask=his_sapo_mask.view(batch_size * his_length, -1)).view(batch_size, his_length, -1)
        user = self.fast_attn(input_embs=history_repr.float(), attention_mask=his_mask)                    # (B, 256), model.py:331
        # click score = candidate . user (model.py:335): the dot-score kernel with ONE interest vector per row
        if torch.is_grad_enabled() and user.requires_grad:
            return torch.matmul(candidate_repr.float(), user.unsqueeze(-1)).squeeze(-1)    # training the encoder body: autograd needs the torch op
        return ops.target_score(user.detach().unsqueeze(1), candidate_repr, None, 'mean')


@torch.no_grad()
def build_news_table(news_encoder: nn.Module, batches: Iterable[Tuple[Tensor, Tensor, Tensor, Optional[Tensor], Optional[Tensor]]],
                     num_news: int, dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None) -> Tensor:
    """One pass of ``news_encoder`` over the news set -> ``(num_news + 1, embed_dim)`` table, row ``id`` = vector of news ``id``.

    ``batches`` yields ``(news_ids (n,), title (n, L), title_mask (n, L), sapo (n, Ls) or None, sapo_mask or None)``; the encoder is
    called with the reference's keywords (news_encoder.py:60-61).  This is the step that replaces the reference's re-encoding of
    every history and candidate title in every batch (model.py:96-111): afterwards ``TableNewsEncoder(table)`` is the news encoder
    of ``Miner`` / ``FastFormer`` and the per-batch work is a row gather.  Rows no batch covers stay zero; a row written twice keeps
    the last write.  The vectors are rounded to ``dtype`` once, here (bf16 by default: what the tensor-core kernels read).
    """
    was_training = news_encoder.training
    news_encoder.eval()
    table = None
    try:
        for ids, title, title_mask, sapo, sapo_mask in batches:
            vec = news_encoder(title_encoding=title, title_attn_mask=title_mask, sapo_encoding=sapo, sapo_attn_mask=sapo_mask)
            if table is None:
                dev = device if device is not None else vec.device
                table = torch.zeros(num_news + 1, vec.shape[1], dtype=dtype, device=dev)
            ids = ids.to(table.device).long()
            if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) > num_news):
                raise IndexError('build_news_table: news id outside [0, num_news]')
            table[ids] = vec.to(table.device, dtype)
    finally:
        news_encoder.train(was_training)
    if table is None:
        raise ValueError('build_news_table: no batches')
    return table
