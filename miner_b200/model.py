"""Drop-in mirror of the reference's ``src/model/model.py`` API for the scoring path.

Same class names, constructor signatures, forward keywords, parameter names and return tuple as the
reference (MrRobot2211/miner):

    Miner(news_encoder, use_category_bias, num_context_codes, context_code_dim, score_type, dropout,
          num_category=None, category_embed_dim=None, category_pad_token_id=None, category_embed=None)   model.py:18-21
    Miner.forward(title, title_mask, his_title, his_title_mask, his_mask, sapo=None, sapo_mask=None,
                  his_sapo=None, his_sapo_mask=None, category=None, his_category=None)
                  -> (multi_user_interest (B,K,D), matching_scores (B,C))                                 model.py:61-64,138
    PolyAttention(in_embed_dim, num_context_codes, context_code_dim).forward(embeddings, attn_mask, bias=None)  model.py:145,159
    TargetAwareAttention(embed_dim).forward(query, key, value)                                            model.py:190,200
    encoder contract: ``embed_dim`` + ``forward(title_encoding, title_attn_mask, sapo_encoding, sapo_attn_mask)``
                                                                                                          news_encoder.py:60-61,108-110

State-dict keys match the reference (``poly_attn.linear.weight``, ``poly_attn.context_codes``,
``target_aware_attn.linear.weight``, ``category_embedding.weight``), so reference checkpoints load.

All arithmetic runs in the sm_100a kernels of libminer_b200.so through ``miner_b200.ops``; there is no
PyTorch fallback -- CPU tensors raise.  The train variant has real backward kernels for the
table-based forward (``miner_train_fwd`` / ``miner_train_bwd``: every ``score_type``, with or without the category bias, optional
sparse gradient of the table rows); the op-level modules on dense tensors (``PolyAttention`` / ``TargetAwareAttention`` behind a
generic encoder) still run forward under autograd and raise NotImplementedError on ``.backward()``.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from . import _lib as L


class _NoBackward(torch.autograd.Function):
    """Marks kernel outputs as differentiable so that a missing backward fails loudly instead of silently."""

    @staticmethod
    def forward(ctx, out: Tensor, *params: Tensor):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError('miner_b200: the backward kernels belong to Miner.forward (any encoder, any score_type, with or without '
                                  'the category bias); the op-level modules called on their own have none')


class _MinerTrainFn(torch.autograd.Function):
    """Train variant of the table-based forward: ``miner_train_fwd`` / ``miner_train_bwd``."""

    @staticmethod
    def forward(ctx, w_proj: Tensor, codes: Tensor, w_target: Optional[Tensor], bias_mean: Optional[Tensor], table: Tensor,
                his_ids: Tensor, his_mask: Tensor, cand_ids: Tensor, math: int, score_type: str):
        interests, scores, saved = ops.train_forward(table, his_ids, his_mask, cand_ids, w_proj, codes, w_target, math, score_type, bias_mean)
        # the big intermediates go through save_for_backward (in-place changes are detected, and saving the OUTPUT `interests` this
        # way does not tie output -> grad_fn -> ctx -> output into a cycle that only the cyclic GC would free: ~3 GB per step)
        ctx.save_for_backward(saved.t, saved.w, saved.z, interests)
        saved.t = saved.w = saved.z = saved.interests = None
        ctx.meta = saved
        return interests, scores

    @staticmethod
    def backward(ctx, d_interests, d_scores):
        saved = ctx.meta
        saved.t, saved.w, saved.z, saved.interests = ctx.saved_tensors
        want_table = ctx.needs_input_grad[4]
        gwp, gc, gwt, dbias, gtab = ops.train_backward(saved, d_scores, d_interests, want_table_grad=want_table)
        saved.t = saved.w = saved.z = saved.interests = saved.ws = None          # release the step's buffers now, not at GC time
        ctx.meta = None
        if gtab is not None and gtab.dtype != saved.table.dtype:
            gtab = gtab.to(saved.table.dtype)
        return gwp, gc, gwt, dbias, gtab, None, None, None, None, None


def _attach(out: Tensor, *params: Tensor) -> Tensor:
    live = [p for p in params if isinstance(p, Tensor) and p.requires_grad]
    if torch.is_grad_enabled() and live:
        return _NoBackward.apply(out, *live)
    return out


class TableNewsEncoder(nn.Module):
    """News encoder restated as an embedding table: ``forward`` returns ``table[title_encoding[:, 0]]``.

    Keeps the NewsEncoder call contract (reference news_encoder.py:60-61,108-110) so it drops into ``Miner``.
    ``table`` is (N, D) float32 or bfloat16 on the GPU; row 0 is the pad news.
    """

    def __init__(self, table: Tensor, trainable: bool = False):
        super().__init__()
        if trainable:
            # the table as a parameter: the train variant returns the gradient of its rows (sparse scatter-add, miner_train_bwd)
            self.table = nn.Parameter(table.detach().contiguous().clone())
        else:
            self.register_buffer('table', table.detach().contiguous())

    @property
    def embed_dim(self) -> int:
        return self.table.shape[1]

    def forward(self, title_encoding: Tensor, title_attn_mask: Tensor = None, sapo_encoding: Union[Tensor, None] = None,
                sapo_attn_mask: Union[Tensor, None] = None) -> Tensor:
        return ops.gather(self.table, title_encoding[:, 0])


class _Linear(nn.Module):
    """``nn.Linear(bias=False)`` parameter holder: keeps the state-dict key ``<name>.weight`` and the default init."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        ref = nn.Linear(in_features=in_features, out_features=out_features, bias=False)
        self.weight = nn.Parameter(ref.weight.detach().clone())
        self.in_features, self.out_features = in_features, out_features


class PolyAttention(nn.Module):
    r"""Poly attention: ``K`` additive attentions over the click history (reference model.py:141-185)."""

    def __init__(self, in_embed_dim: int, num_context_codes: int, context_code_dim: int):
        super().__init__()
        self.linear = _Linear(in_embed_dim, context_code_dim)
        self.context_codes = nn.Parameter(nn.init.xavier_uniform_(torch.empty(num_context_codes, context_code_dim),
                                                                  gain=nn.init.calculate_gain('tanh')))

    def forward(self, embeddings: Tensor, attn_mask: Tensor, bias: Tensor = None):
        r"""
        Args:
            embeddings: ``(batch_size, his_length, embed_dim)``
            attn_mask: ``(batch_size, his_length)`` bool
            bias: ``(batch_size, his_length, num_candidates)`` or None
        Returns: ``(batch_size, num_context_codes, embed_dim)``
        """
        bias_mean = None if bias is None else bias.mean(dim=2)          # model.py:176 (plumbing on a caller tensor)
        out = ops.poly_attention(embeddings, attn_mask, self.linear.weight, self.context_codes, bias_mean)
        return _attach(out, embeddings, self.linear.weight, self.context_codes)


class TargetAwareAttention(nn.Module):
    """Target-aware attention network (reference model.py:188-216)."""

    def __init__(self, embed_dim: int):
        super().__init__()
        self.linear = _Linear(embed_dim, embed_dim)

    def forward(self, query: Tensor, key: Tensor, value: Tensor):
        r"""
        Args:
            query: ``(batch_size, num_context_codes, embed_dim)``
            key: ``(batch_size, num_candidates, embed_dim)``
            value: ``(batch_size, num_candidates, num_context_codes)``
        Returns: ``(batch_size, num_candidates)``
        """
        B, Cn, _ = key.shape
        out = ops.target_score(query, key, self.linear.weight, 'weighted', matching=value.reshape(B * Cn, -1))
        return _attach(out, query, key, value, self.linear.weight)


class Miner(nn.Module):
    r"""MINER scoring path (reference model.py:13-138) on the B200 kernels.

    With a :class:`TableNewsEncoder` the whole forward is the fused table-based path (gather -> poly attention ->
    target-aware aggregation -> scores; the history tile is never materialised).  With any other encoder module the
    encoder is called exactly as the reference does and the op-level kernels take its dense outputs.
    """

    def __init__(self, news_encoder: nn.Module, use_category_bias: bool, num_context_codes: int,
                 context_code_dim: int, score_type: str, dropout: float, num_category: Union[int, None] = None,
                 category_embed_dim: Union[int, None] = None, category_pad_token_id: Union[int, None] = None,
                 category_embed: Union[Tensor, None] = None):
        super().__init__()
        self.news_encoder = news_encoder
        self.news_embed_dim = self.news_encoder.embed_dim
        self.use_category_bias = use_category_bias
        if self.use_category_bias:
            self.category_dropout = nn.Dropout(dropout)
            if category_embed is not None:
                self.category_embedding = nn.Embedding.from_pretrained(category_embed, freeze=False,
                                                                       padding_idx=category_pad_token_id)
                self.category_embed_dim = category_embed.shape[1]
            else:
                assert num_category is not None
                self.category_embedding = nn.Embedding(num_embeddings=num_category, embedding_dim=category_embed_dim,
                                                       padding_idx=category_pad_token_id)
                self.category_embed_dim = category_embed_dim
        self.poly_attn = PolyAttention(in_embed_dim=self.news_embed_dim, num_context_codes=num_context_codes,
                                       context_code_dim=context_code_dim)
        self.score_type = score_type
        if self.score_type == 'weighted':
            self.target_aware_attn = TargetAwareAttention(self.news_embed_dim)
        self.dropout = nn.Dropout(dropout)           # constructed but never applied, as in the reference (model.py:59)
        self._prepared = None                        # (version key, ops.ScoreWeights)
        self._table_proj = None                      # (version key, ops.TableProjections)
        self.table_level = False                     # forward(): opt into the table-level mode (score_impressions uses it by default)
        self.train_math = 'fp32'                     # train variant: 'fp32' (reference arithmetic) or 'tensor' (bf16 tcgen05 GEMMs, as autocast)

    # -- parameter staging -------------------------------------------------------------------------------------
    def _weights(self, with_bf16: bool) -> ops.ScoreWeights:
        wt = self.target_aware_attn.linear.weight if self.score_type == 'weighted' else None
        params = (self.poly_attn.linear.weight, self.poly_attn.context_codes, wt)
        key = tuple((p.data_ptr(), p._version) if p is not None else None for p in params) + (with_bf16,)
        if self._prepared is None or self._prepared[0] != key:
            self._prepared = (key, ops.ScoreWeights(params[0], params[1], wt, with_bf16))
        return self._prepared[1]

    def table_projections(self) -> 'ops.TableProjections':
        """Table-level projections ``lg = tanh(table Wp^T) codes^T`` and ``tw = table Wt^T`` (model.py:171,174,212 applied once per
        table row), recomputed whenever a parameter or the table changes."""
        if not isinstance(self.news_encoder, TableNewsEncoder):
            raise L.MinerError('table-level mode needs a TableNewsEncoder')
        table = self.news_encoder.table
        w = self._weights(with_bf16=True)
        key = (self._prepared[0], table.data_ptr(), table._version)
        if self._table_proj is None or self._table_proj[0] != key:
            self._table_proj = (key, ops.table_project(table, w, weighted=self.score_type == 'weighted'))
        return self._table_proj[1]

    def _table_ws(self, B: int, H: int) -> Tensor:
        """Workspace of the table-level kernel, kept across calls (its out-of-range id counters accumulate; ``check_bounds``)."""
        K = self.poly_attn.context_codes.shape[0]
        need = L.load().miner_score_table_workspace_bytes(B, H, K)
        ws = getattr(self, '_tws', None)
        if ws is None or ws.numel() < need or ws.device != self.news_encoder.table.device:
            if ws is not None:
                ops.check_oob(ws)
            self._tws = ws = ops.score_table_workspace(B, H, K, self.news_encoder.table.device)
        return ws

    def check_bounds(self) -> None:
        """Raise IndexError if any table-level call so far saw a news id outside the table (the reference's indexing would have)."""
        if getattr(self, '_tws', None) is not None:
            ops.check_oob(self._tws)

    def invalidate(self) -> None:
        """Drop the staged weight copies and table projections.  They are keyed on ``(data_ptr, _version)`` of the parameters and
        the table; an update through ``p.data.*`` does not bump ``_version`` -- call this after one."""
        self._prepared = None
        self._table_proj = None

    def _bias_mean(self, his_category: Tensor, category: Tensor) -> Tensor:
        """``category_bias.mean(dim=2)`` (model.py:113-120,176).  Inference: the ``miner_category_bias`` kernel.  Under autograd the
        (B,H,Ec) x (B,Ec,C) cosine with its train-mode dropout is the reference's own handful of torch ops on tiny tensors, so the
        category embedding gets its gradient through torch autograd; the scoring kernels take ``bias_mean`` and return its gradient."""
        if torch.is_grad_enabled() and (self.category_embedding.weight.requires_grad or self.training):
            he = self.category_dropout(self.category_embedding(his_category))              # model.py:113-114
            ce = self.category_dropout(self.category_embedding(category))                  # model.py:115-116
            he = torch.div(he, torch.linalg.norm(he, dim=2, keepdim=True))                 # utils.py:21-23
            ce = torch.div(ce, torch.linalg.norm(ce, dim=2, keepdim=True))
            return torch.matmul(he, ce.permute(0, 2, 1)).mean(dim=2)                       # model.py:176
        mean, _ = ops.category_bias(self.category_embedding.weight, his_category, category)
        return mean

    def forward(self, title: Tensor, title_mask: Tensor, his_title: Tensor, his_title_mask: Tensor,
                his_mask: Tensor, sapo: Union[Tensor, None] = None, sapo_mask: Union[Tensor, None] = None,
                his_sapo: Union[Tensor, None] = None, his_sapo_mask: Union[Tensor, None] = None,
                category: Union[Tensor, None] = None, his_category: Union[Tensor, None] = None):
        r"""
        Returns:
            tuple
                - multi_user_interest: ``(batch_size, num_context_codes, embed_dim)``
                - matching_scores: ``(batch_size, num_candidates)``
        """
        if self.score_type not in L.SCORE_TYPES:
            raise ValueError('Invalid method of aggregating matching score')        # model.py:136
        batch_size, num_candidates = title.shape[0], title.shape[1]
        his_length = his_title.shape[1]
        bias_mean = self._bias_mean(his_category, category) if self.use_category_bias else None

        if isinstance(self.news_encoder, TableNewsEncoder):
            table = self.news_encoder.table
            math = ops.default_math(table, table.shape[1])
            w = self._weights(with_bf16=(math == L.MATH_TENSOR))
            cand_ids = title.reshape(batch_size, num_candidates, -1)[..., 0]
            his_ids = his_title.reshape(batch_size, his_length, -1)[..., 0]
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                # train variant (reference trainer.py:246-261): forward that keeps its intermediates + real backward kernels
                wt = self.target_aware_attn.linear.weight if self.score_type == 'weighted' else None
                return _MinerTrainFn.apply(self.poly_attn.linear.weight, self.poly_attn.context_codes, wt, bias_mean, table, his_ids,
                                           his_mask, cand_ids,
                                           L.MATH_TENSOR if (self.train_math == 'tensor' and table.dtype == torch.bfloat16
                                                             and table.shape[1] % 64 == 0) else L.MATH_FP32, self.score_type)
            if self.table_level and table.dtype == torch.bfloat16 and ops.score_table_supported(his_length, self.poly_attn.context_codes.shape[0],
                                                                                               table.shape[1]):
                interests, scores = ops.score_table(self.table_projections(), his_ids, his_mask, cand_ids, self.score_type,
                                                    bias_mean=bias_mean, want_interests=True,
                                                    workspace=self._table_ws(batch_size, his_length))
                params = [p for p in self.parameters()]
                return _attach(interests, *params), _attach(scores, *params)
            interests, scores = ops.score(table, his_ids, his_mask, cand_ids, w, self.score_type, bias_mean=bias_mean,
                                          math=math, want_interests=True)
        else:
            # encoder called exactly as the reference does (model.py:86-111)
            title = title.view(batch_size * num_candidates, -1)
            title_mask = title_mask.view(batch_size * num_candidates, -1)
            sapo = sapo.view(batch_size * num_candidates, -1)
            sapo_mask = sapo_mask.view(batch_size * num_candidates, -1)
            candidate_repr = self.news_encoder(title_encoding=title, title_attn_mask=title_mask, sapo_encoding=sapo,
                                               sapo_attn_mask=sapo_mask).view(batch_size, num_candidates, -1)
            his_title = his_title.view(batch_size * his_length, -1)
            his_title_mask = his_title_mask.view(batch_size * his_length, -1)
            his_sapo = his_sapo.view(batch_size * his_length, -1)
            his_sapo_mask = his_sapo_mask.view(batch_size * his_length, -1)
            history_repr = self.news_encoder(title_encoding=his_title, title_attn_mask=his_title_mask,
                                             sapo_encoding=his_sapo, sapo_attn_mask=his_sapo_mask).view(batch_size, his_length, -1)
            if torch.is_grad_enabled() and (history_repr.requires_grad or any(p.requires_grad for p in self.parameters())):
                # train variant behind ANY differentiable encoder: its dense outputs are the "table" (rows = the B*H history vectors
                # followed by the B*C candidate vectors), the ids are just row numbers, and the table-row gradient of the backward
                # kernels is d loss / d encoder output -- autograd carries it on into the encoder (reference trainer.py:146-169)
                D = history_repr.shape[-1]
                dev = history_repr.device
                virt = torch.cat([history_repr.reshape(-1, D), candidate_repr.reshape(-1, D)], dim=0).float()
                hid = torch.arange(batch_size * his_length, device=dev).view(batch_size, his_length)
                cid = (batch_size * his_length + torch.arange(batch_size * num_candidates, device=dev)).view(batch_size, num_candidates)
                wt = self.target_aware_attn.linear.weight if self.score_type == 'weighted' else None
                return _MinerTrainFn.apply(self.poly_attn.linear.weight, self.poly_attn.context_codes, wt, bias_mean, virt, hid, his_mask, cid,
                                           L.MATH_FP32, self.score_type)
            interests = ops.poly_attention(history_repr, his_mask, self.poly_attn.linear.weight, self.poly_attn.context_codes,
                                           bias_mean)
            wt = self.target_aware_attn.linear.weight if self.score_type == 'weighted' else None
            scores = ops.target_score(interests, candidate_repr, wt, self.score_type)
        params = [p for p in self.parameters()]
        return _attach(interests, *params), _attach(scores, *params)

    # -- grouped (CSR) evaluation entry: extension beyond the reference API ---------------------------------------
    @torch.no_grad()
    def score_impressions(self, his_ids: Tensor, his_mask: Tensor, cand_ids: Tensor, cand_offsets: Tensor,
                          chunk: int = 16384, math: Optional[int] = None) -> Tensor:
        """Scores of every candidate of every impression, CSR layout (``cand_offsets`` (B+1,) into flat ``cand_ids``).

        ``math``: ``MATH_TABLE`` (default when the shape allows: bf16 table, H <= 256, K <= 64, D % 64 == 0) applies the two linear
        layers once per table row and scores in one kernel; ``MATH_TENSOR`` / ``MATH_FP32`` keep the reference operation order.

        Equals the reference's per-candidate eval rows (src/reader.py:376-379: one sample per candidate, interests
        recomputed per candidate) when category bias is off, but computes the interests once per impression.
        """
        if self.use_category_bias:
            raise NotImplementedError('grouped scoring with category bias: the bias depends on the row layout (model.py:176); '
                                      'use forward() with the layout you mean')
        if not isinstance(self.news_encoder, TableNewsEncoder):
            raise L.MinerError('score_impressions needs a TableNewsEncoder')
        table = self.news_encoder.table
        if math is None:
            math = ops.default_eval_math(table, his_ids.shape[1], self.poly_attn.context_codes.shape[0])
        if math == L.MATH_TABLE:
            _, scores = ops.score_table(self.table_projections(), his_ids, his_mask, cand_ids, self.score_type, cand_offsets=cand_offsets,
                                        workspace=self._table_ws(his_ids.shape[0], his_ids.shape[1]))
            return scores
        w = self._weights(with_bf16=(math == L.MATH_TENSOR))
        _, scores = ops.score(table, his_ids, his_mask, cand_ids, w, self.score_type, cand_offsets=cand_offsets, math=math,
                              chunk=chunk)
        return scores
