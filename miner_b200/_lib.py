"""ctypes binding of libminer_b200.so (the C ABI declared in include/miner_b200.h).

There is NO fallback: if the shared library cannot be loaded (or built with nvcc), importing any
compute entry point raises.  PyTorch is only used by callers for device memory and streams; nothing
here takes a torch type.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MINER_B200_LIB') or os.path.join(_PKG, 'libminer_b200.so')   # override: instrumented builds

# constants mirrored from include/miner_b200.h
OK, ERR_INVALID_ARG, ERR_SCORE_TYPE, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4, 5
F32, BF16 = 0, 1
I32, I64 = 0, 1
SCORE_MAX, SCORE_MEAN, SCORE_WEIGHTED = 0, 1, 2
MATH_FP32, MATH_TENSOR, MATH_TABLE = 0, 1, 2
EPI_NONE, EPI_TANH, EPI_GELU = 0, 1, 2
ABI_VERSION = 2

SCORE_TYPES = {'max': SCORE_MAX, 'mean': SCORE_MEAN, 'weighted': SCORE_WEIGHTED}

EXPORTS = [
    'miner_abi_version', 'miner_last_error', 'miner_device_info', 'miner_launch_count', 'miner_gather', 'miner_category_bias',
    'miner_poly_attn_workspace_bytes', 'miner_poly_attn_fwd', 'miner_target_score_workspace_bytes',
    'miner_target_score_fwd', 'miner_score_workspace_bytes', 'miner_score_fwd', 'miner_cast_f32_to_bf16',
    'miner_tc_gemm', 'miner_tc_gemm_tn', 'miner_rank_metrics_workspace_bytes', 'miner_rank_metrics', 'miner_loss_workspace_bytes',
    'miner_loss_fwd', 'miner_hist_interests_workspace_bytes', 'miner_hist_interests_fwd', 'miner_cand_score_fwd',
    'miner_table_project_workspace_bytes', 'miner_table_project', 'miner_score_table_supported', 'miner_score_table_fwd',
    'miner_score_table_workspace_bytes', 'miner_score_table_tile_geometry',
    'miner_auc_split', 'miner_sort_u32_workspace_bytes', 'miner_sort_u32', 'miner_auc_count',
    'miner_train_workspace_bytes', 'miner_train_fwd', 'miner_loss_bwd', 'miner_train_bwd', 'miner_train_table_grad_workspace_bytes',
]


class ScoreParams(C.Structure):
    """struct miner_score_params"""
    _fields_ = [
        ('table', C.c_void_p), ('n_rows', C.c_int64), ('table_dtype', C.c_int),
        ('his_ids', C.c_void_p), ('his_mask', C.c_void_p),
        ('cand_ids', C.c_void_p), ('cand_offsets', C.c_void_p), ('id_dtype', C.c_int),
        ('bias_mean', C.c_void_p),
        ('w_proj', C.c_void_p), ('codes', C.c_void_p), ('w_target', C.c_void_p),
        ('w_proj_bf16', C.c_void_p), ('w_target_bf16', C.c_void_p),
        ('B', C.c_int64), ('H', C.c_int64), ('C', C.c_int64), ('K', C.c_int64), ('Dc', C.c_int64), ('D', C.c_int64),
        ('T', C.c_int64),
        ('score_type', C.c_int), ('math', C.c_int),
        ('out_scores', C.c_void_p), ('out_interests', C.c_void_p), ('stage_mask', C.c_int),
    ]


class MinerError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL) -> None:
    vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
    lib.miner_abi_version.restype = i32
    lib.miner_last_error.restype = C.c_char_p
    lib.miner_device_info.argtypes = [C.POINTER(i32)] * 3
    lib.miner_launch_count.restype = C.c_uint64
    lib.miner_gather.argtypes = [vp, i64, i64, i32, vp, i64, i32, vp, vp, vp]
    lib.miner_category_bias.argtypes = [vp, i64, i64, vp, vp, i32, i64, i64, i64, vp, vp, vp]
    lib.miner_poly_attn_workspace_bytes.argtypes = [i64] * 5
    lib.miner_poly_attn_workspace_bytes.restype = sz
    lib.miner_poly_attn_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, vp, vp, vp, sz, vp]
    lib.miner_target_score_workspace_bytes.argtypes = [i64, i64, i64, i32]
    lib.miner_target_score_workspace_bytes.restype = sz
    lib.miner_target_score_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i64, i64, i64, i64, vp, vp, sz, vp]
    lib.miner_score_workspace_bytes.argtypes = [C.POINTER(ScoreParams), i64]
    lib.miner_score_workspace_bytes.restype = sz
    lib.miner_score_fwd.argtypes = [C.POINTER(ScoreParams), i64, vp, sz, vp]
    lib.miner_cast_f32_to_bf16.argtypes = [vp, vp, i64, vp]
    lib.miner_tc_gemm.argtypes = [vp, vp, i32, i64, vp, vp, vp, i64, i64, i64, i32, vp]
    lib.miner_tc_gemm_tn.argtypes = [vp, vp, vp, i64, i64, i64, i32, vp]
    lib.miner_rank_metrics_workspace_bytes.argtypes = [i64, i32]
    lib.miner_rank_metrics_workspace_bytes.restype = sz
    lib.miner_rank_metrics.argtypes = [vp, vp, vp, i64, i32, C.POINTER(i32), i32, vp, vp, vp, sz, vp]
    lib.miner_loss_workspace_bytes.argtypes = [i64, i64]
    lib.miner_loss_workspace_bytes.restype = sz
    lib.miner_loss_fwd.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, vp, vp, sz, vp]
    lib.miner_hist_interests_workspace_bytes.argtypes = [i64]
    lib.miner_hist_interests_workspace_bytes.restype = sz
    lib.miner_hist_interests_fwd.argtypes = [vp, i64, vp, i32, vp, vp, vp, vp, i64, i64, i64, i64, i64, vp, vp, vp, vp, sz, vp]
    lib.miner_cand_score_fwd.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, i64, i64, i64, i64, vp, vp]
    lib.miner_table_project_workspace_bytes.argtypes = [i64, i64]
    lib.miner_table_project_workspace_bytes.restype = sz
    lib.miner_table_project.argtypes = [vp, i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, sz, vp]
    lib.miner_score_table_supported.argtypes = [i64, i64, i64]
    lib.miner_train_workspace_bytes.argtypes = [i64, i64, i64, i64, i64, i32]
    lib.miner_train_workspace_bytes.restype = sz
    lib.miner_train_fwd.argtypes = [vp, i64, i32, vp, vp, vp, i32, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, vp, vp, vp, vp, i32, vp, vp,
                                    i32, vp, vp, sz, vp]
    lib.miner_loss_bwd.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp]
    lib.miner_train_bwd.argtypes = [vp, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, vp, vp,
                                    i32, vp, vp, i32, vp, vp, vp, sz, vp, sz, vp]
    lib.miner_train_table_grad_workspace_bytes.argtypes = [i64, i64, i64, i64]
    lib.miner_train_table_grad_workspace_bytes.restype = sz
    lib.miner_score_table_fwd.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i32, vp, i64, i64, i64, i64, i64, i32, vp, vp, vp, sz, vp]
    lib.miner_score_table_workspace_bytes.argtypes = [i64, i64, i64]
    lib.miner_score_table_workspace_bytes.restype = sz
    lib.miner_score_table_tile_geometry.argtypes = [i64, i64, C.POINTER(i32), C.POINTER(i32)]
    lib.miner_auc_split.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp, vp, vp]
    lib.miner_sort_u32_workspace_bytes.argtypes = [i64]
    lib.miner_sort_u32_workspace_bytes.restype = sz
    lib.miner_sort_u32.argtypes = [vp, i64, vp, sz, vp]
    lib.miner_auc_count.argtypes = [vp, i64, vp, i64, vp, vp]


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building once with nvcc if absent) the CUDA library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise MinerError(f'{LIB_PATH} is missing; run `python -m miner_b200.build`')
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    missing = [n for n in EXPORTS if not hasattr(lib, n)]
    if missing:
        raise MinerError(f'libminer_b200.so does not export {missing}; rebuild with `python -m miner_b200.build --force`')
    _declare(lib)
    if lib.miner_abi_version() != ABI_VERSION:
        raise MinerError('libminer_b200.so ABI version mismatch; rebuild with `python -m miner_b200.build --force`')
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Translate a C-ABI return code into the exception the reference would raise."""
    if rc == OK:
        return
    msg = load().miner_last_error().decode('utf-8', 'replace')
    if rc == ERR_SCORE_TYPE:
        raise ValueError('Invalid method of aggregating matching score')     # reference model.py:136
    raise MinerError(f'miner_b200 error {rc}: {msg}')
