"""Mirror of the hot-path helper of the reference's ``src/utils.py``."""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops


def category_cosine_bias(category_embedding: Tensor, his_category: Tensor, category: Tensor) -> Tensor:
    """``pairwise_cosine_similarity(E[his_category], E[category])`` (reference utils.py:9-29 as called at model.py:120),
    computed by the category-bias kernel.  Returns ``(B, H, C)``."""
    _, full = ops.category_bias(category_embedding, his_category, category, want_full=True)
    return full
