"""Mirror of the reference's ``src/loss.py`` for the MINER path (``Loss.compute`` / ``Loss.compute_eval_loss``)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops


class _LossFn(torch.autograd.Function):
    """Loss.compute with its backward kernel (``miner_loss_fwd`` / ``miner_loss_bwd``)."""

    @staticmethod
    def forward(ctx, poly_attn: Tensor, logits: Tensor, labels: Tensor):
        out = ops.loss_forward(poly_attn, logits, labels, eval_mode=False)
        ctx.save_for_backward(poly_attn, logits, labels)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        poly_attn, logits, labels = ctx.saved_tensors
        d_i, d_l = ops.loss_backward(poly_attn, logits, labels, grad_out)
        return d_i, d_l, None


class Loss:
    def __init__(self, criterion=None):
        # the reference passes nn.CrossEntropyLoss(reduction='mean') (src/trainer.py:303); the fused kernel implements
        # exactly that criterion, anything else is rejected rather than silently mis-computed
        if criterion is not None:
            import torch.nn as nn
            if not isinstance(criterion, nn.CrossEntropyLoss) or criterion.reduction != 'mean' or criterion.weight is not None \
                    or criterion.label_smoothing != 0.0:
                raise NotImplementedError('miner_b200.Loss implements nn.CrossEntropyLoss(reduction="mean") only')
        self._criterion = criterion

    def compute(self, poly_attn: Tensor, logits: Tensor, labels: Tensor) -> Tensor:
        r"""Disagreement (mean pairwise cosine of the K interests, zero diagonal) + cross-entropy (reference loss.py:27-44).

        poly_attn ``(B, K, D)``, logits ``(B, npratio+1)``, labels one-hot ``(B, npratio+1)``.  Returns a 0-dim tensor.
        """
        labels = labels.to(torch.float32)
        if torch.is_grad_enabled() and (poly_attn.requires_grad or logits.requires_grad):
            return _LossFn.apply(poly_attn, logits, labels)
        return ops.loss_forward(poly_attn, logits, labels, eval_mode=False)[0]

    @staticmethod
    def compute_eval_loss(poly_attn: Tensor, logits: Tensor, labels: Tensor) -> float:
        r"""Reference loss.py:68-85: disagreement + ``-(logsigmoid(logits) * labels).sum()``; returns a python float."""
        B = logits.shape[0]
        out = ops.loss_forward(poly_attn, logits.reshape(B, -1), labels.reshape(B, -1).to(torch.float32), eval_mode=True)
        return float(out[0].item())
