"""miner_b200 -- B200-native (sm_100a) implementation of the MINER data-parallel scoring path.

Public API mirrors the reference's ``src/model`` / ``src/evaluation`` / ``src/loss`` for this path:
``Miner``, ``PolyAttention``, ``TargetAwareAttention``, ``TableNewsEncoder`` (NewsEncoder contract),
``SlowEvaluator``, ``FastEvaluator``, ``Loss``, plus ``FastFormer`` (the Fastformer user-encoder variant on the same gather / score
kernels) and ``build_news_table``.  See include/miner_b200.h for the C ABI underneath.
"""
from .model import Miner, PolyAttention, TargetAwareAttention, TableNewsEncoder  # noqa: F401
from .evaluation import SlowEvaluator, FastEvaluator  # noqa: F401
from .loss import Loss  # noqa: F401
from .pipeline import HostEvaluator  # noqa: F401
from .fastformer import FastFormer, FastformerEncoder, build_news_table  # noqa: F401

__all__ = ['Miner', 'PolyAttention', 'TargetAwareAttention', 'TableNewsEncoder', 'SlowEvaluator', 'FastEvaluator', 'Loss', 'HostEvaluator',
           'FastFormer', 'FastformerEncoder', 'build_news_table']
