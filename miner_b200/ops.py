"""Functional layer: torch tensors in, C-ABI calls out (include/miner_b200.h), torch tensors back.

PyTorch supplies device memory and the current stream; all arithmetic happens in the hand-written
sm_100a kernels of libminer_b200.so.  Every function requires CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise L.MinerError('miner_b200 kernels need CUDA tensors (there is no CPU fallback)')
        if dev is not None and t.device != dev:
            raise L.MinerError('all tensors must live on the same CUDA device')
        dev = t.device
    if dev is None:
        raise L.MinerError('no tensor given')
    return dev


def _ids(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    if t.dtype == torch.int64:
        return t.contiguous(), L.I64
    if t.dtype == torch.int32:
        return t.contiguous(), L.I32
    raise L.MinerError(f'ids must be int32 or int64, got {t.dtype}')


def _table_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise L.MinerError(f'embedding table must be float32 or bfloat16, got {t.dtype}')


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _mask_u8(mask: torch.Tensor) -> torch.Tensor:
    m = mask.contiguous()
    if m.dtype == torch.bool:
        return m.view(torch.uint8)
    return (m != 0).view(torch.uint8)


def launch_count() -> int:
    """Kernels launched by libminer_b200.so in this process so far."""
    return int(L.load().miner_launch_count())


def device_info() -> Dict[str, int]:
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    L.check(L.load().miner_device_info(C.byref(sm), C.byref(major), C.byref(minor)))
    return {'sm_count': sm.value, 'cc_major': major.value, 'cc_minor': minor.value}


# ------------------------------------------------------------------------------------------------ (a1)
def gather(table: torch.Tensor, ids: torch.Tensor, check_bounds: bool = False) -> torch.Tensor:
    """``table[ids]`` -- bit-exact row gather (NewsEncoder contract, reference model.py:96-97,109-110)."""
    dev = _need_cuda(table, ids)
    lib = L.load()
    table = table.detach().contiguous()
    idc, it = _ids(ids)
    out = torch.empty(*ids.shape, table.shape[1], dtype=table.dtype, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev) if check_bounds else None
    with torch.cuda.device(dev):
        L.check(lib.miner_gather(_ptr(table), table.shape[0], table.shape[1], _table_dtype(table), _ptr(idc), idc.numel(), it,
                                 _ptr(out), _ptr(flag), _stream()))
    if check_bounds and int(flag.item()):
        raise IndexError('index out of range in embedding table')      # what torch indexing raises on the CPU
    return out


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(src)
    s = _f32(src)
    out = torch.empty(s.shape, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().miner_cast_f32_to_bf16(_ptr(s), _ptr(out), s.numel(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ (a2)
def category_bias(cat_emb: torch.Tensor, his_cat: torch.Tensor, cand_cat: torch.Tensor, want_full: bool = False):
    """Returns ``(bias_mean (B,H), bias_full (B,H,C) or None)`` (reference model.py:113-120,176)."""
    dev = _need_cuda(cat_emb, his_cat, cand_cat)
    B, H = his_cat.shape
    Cn = cand_cat.shape[1]
    hc, it = _ids(his_cat)
    cc, it2 = _ids(cand_cat)
    if it != it2:
        cc, it2 = _ids(cand_cat.to(his_cat.dtype))
    emb = _f32(cat_emb)
    mean = torch.empty(B, H, dtype=torch.float32, device=dev)
    full = torch.empty(B, H, Cn, dtype=torch.float32, device=dev) if want_full else None
    with torch.cuda.device(dev):
        L.check(L.load().miner_category_bias(_ptr(emb), emb.shape[0], emb.shape[1], _ptr(hc), _ptr(cc), it, B, H, Cn,
                                             _ptr(full), _ptr(mean), _stream()))
    return mean, full


# ------------------------------------------------------------------------------------------------ (a3)
def poly_attention(emb: torch.Tensor, mask: torch.Tensor, w_proj: torch.Tensor, codes: torch.Tensor,
                   bias_mean: Optional[torch.Tensor] = None, return_weights: bool = False):
    """PolyAttention.forward on a dense (B,H,D) history tensor (reference model.py:159-185)."""
    dev = _need_cuda(emb, mask, w_proj, codes, bias_mean)
    lib = L.load()
    B, H, D = emb.shape
    K, Dc = codes.shape
    e, wp, cd = _f32(emb), _f32(w_proj), _f32(codes)
    m = _mask_u8(mask)
    bm = _f32(bias_mean) if bias_mean is not None else None
    out = torch.empty(B, K, D, dtype=torch.float32, device=dev)
    wts = torch.empty(B, K, H, dtype=torch.float32, device=dev) if return_weights else None
    ws_bytes = lib.miner_poly_attn_workspace_bytes(B, H, K, Dc, D)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_poly_attn_fwd(_ptr(e), _ptr(m), _ptr(bm), _ptr(wp), _ptr(cd), B, H, K, Dc, D, _ptr(out), _ptr(wts),
                                        _ptr(ws), ws_bytes, _stream()))
    return (out, wts) if return_weights else out


# ------------------------------------------------------------------------------------------------ (a4, a5)
def target_score(interests: torch.Tensor, cand: torch.Tensor, w_target: Optional[torch.Tensor], score_type: str,
                 cand_offsets: Optional[torch.Tensor] = None, matching: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Matching scores + aggregation (reference model.py:127-136, 200-216).

    ``cand`` is (B,C,D) dense or, with ``cand_offsets`` (B+1,) int64, a flat (T,D) CSR block.
    """
    dev = _need_cuda(interests, cand, w_target, cand_offsets, matching)
    lib = L.load()
    st = L.SCORE_TYPES.get(score_type, -1)
    if st < 0:
        raise ValueError('Invalid method of aggregating matching score')
    B, K, D = interests.shape
    I, cd = _f32(interests), _f32(cand)
    wt = _f32(w_target) if w_target is not None else None
    if cand_offsets is None:
        Cn = cand.shape[1]
        out = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        offs = None
    else:
        Cn = 0
        out = torch.empty(cand.shape[0], dtype=torch.float32, device=dev)
        offs = cand_offsets.to(torch.int64).contiguous()
    mt = _f32(matching) if matching is not None else None
    ws_bytes = lib.miner_target_score_workspace_bytes(B, K, D, st)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_target_score_fwd(_ptr(I), _ptr(cd), _ptr(mt), _ptr(offs), _ptr(wt), st, B, Cn, K, D, _ptr(out), _ptr(ws),
                                           ws_bytes, _stream()))
    return out


# ------------------------------------------------------------------------------------------------ (a1..a6) fused
class ScoreWeights:
    """Device copies of the path's parameters in the layouts the kernels read (fp32 + bf16 for the tensor family)."""

    def __init__(self, w_proj: torch.Tensor, codes: torch.Tensor, w_target: Optional[torch.Tensor], with_bf16: bool):
        self.w_proj, self.codes = _f32(w_proj), _f32(codes)
        self.w_target = _f32(w_target) if w_target is not None else None
        self.w_proj_bf16 = cast_bf16(self.w_proj) if with_bf16 else None
        self.w_target_bf16 = cast_bf16(self.w_target) if (with_bf16 and self.w_target is not None) else None


def default_math(table: torch.Tensor, D: int) -> int:
    """Tensor-core family whenever it applies: bf16 table and D a multiple of 64."""
    return L.MATH_TENSOR if (table.dtype == torch.bfloat16 and D % 64 == 0 and D >= 64) else L.MATH_FP32


def default_eval_math(table: torch.Tensor, H: int, K: int) -> int:
    """Grouped (CSR) evaluation: the table-level mode whenever its kernel covers the shape, else :func:`default_math`."""
    N, D = table.shape[0], table.shape[1]
    # the table-level kernel addresses rows with 32-bit byte offsets: tables of 4 GB and more run the reference-order family
    fits = N * D * 2 <= 0xffffffff and N < (1 << 30) and N * K <= 0xffffffff
    if table.dtype == torch.bfloat16 and fits and score_table_supported(H, K, D):
        return L.MATH_TABLE
    return default_math(table, D)


def score(table: torch.Tensor, his_ids: torch.Tensor, his_mask: torch.Tensor, cand_ids: torch.Tensor, weights: ScoreWeights,
          score_type: str = 'weighted', cand_offsets: Optional[torch.Tensor] = None, bias_mean: Optional[torch.Tensor] = None,
          math: Optional[int] = None, want_interests: bool = False, chunk: int = 16384,
          out_scores: Optional[torch.Tensor] = None, stage_mask: int = 0, workspace: Optional[torch.Tensor] = None):
    """Miner.forward for a block of impressions straight from the embedding table (reference model.py:61-138).

    Dense layout: ``cand_ids`` (B,C).  CSR layout: ``cand_ids`` (T,) + ``cand_offsets`` (B+1,) int64.
    Returns ``(interests (B,K,D) or None, scores)``.
    """
    dev = _need_cuda(table, his_ids, his_mask, cand_ids, cand_offsets, bias_mean)
    lib = L.load()
    st = L.SCORE_TYPES.get(score_type, -1)
    if st < 0:
        raise ValueError('Invalid method of aggregating matching score')
    table = table.detach().contiguous()
    B, H = his_ids.shape
    D = table.shape[1]
    K, Dc = weights.codes.shape
    hid, it = _ids(his_ids)
    cid, it2 = _ids(cand_ids)
    if it != it2:
        cid, it2 = _ids(cand_ids.to(his_ids.dtype))
    m = _mask_u8(his_mask)
    if math is None:
        math = default_math(table, D)
    if math == L.MATH_TENSOR and weights.w_proj_bf16 is None:
        raise L.MinerError('tensor-core family requested but the weights were prepared without bf16 copies')
    p = L.ScoreParams()
    p.table, p.n_rows, p.table_dtype = _ptr(table), table.shape[0], _table_dtype(table)
    p.his_ids, p.his_mask = _ptr(hid), _ptr(m)
    p.cand_ids, p.id_dtype = _ptr(cid), it
    if cand_offsets is None:
        Cn = cand_ids.shape[1]
        T = B * Cn
        offs = None
        out_shape = (B, Cn)
    else:
        Cn = 0
        T = cid.numel()
        offs = cand_offsets.to(torch.int64).contiguous()
        out_shape = (T,)
    p.cand_offsets = _ptr(offs)
    bm = _f32(bias_mean) if bias_mean is not None else None
    p.bias_mean = _ptr(bm)
    p.w_proj, p.codes, p.w_target = _ptr(weights.w_proj), _ptr(weights.codes), _ptr(weights.w_target)
    p.w_proj_bf16, p.w_target_bf16 = _ptr(weights.w_proj_bf16), _ptr(weights.w_target_bf16)
    p.B, p.H, p.C, p.K, p.Dc, p.D, p.T = B, H, Cn, K, Dc, D, T
    p.score_type, p.math = st, math
    scores = out_scores if out_scores is not None else torch.empty(out_shape, dtype=torch.float32, device=dev)
    interests = torch.empty(B, K, D, dtype=torch.float32, device=dev) if want_interests else None
    p.out_scores, p.out_interests = _ptr(scores), _ptr(interests)
    p.stage_mask = stage_mask
    chunk = max(1, min(int(chunk), max(B, 1)))
    ws_bytes = lib.miner_score_workspace_bytes(C.byref(p), chunk)
    ws = workspace if (workspace is not None and workspace.numel() >= ws_bytes) else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_score_fwd(C.byref(p), chunk, _ptr(ws), ws.numel(), _stream()))
    return interests, scores


# ------------------------------------------------------------------------------------------------ table-level mode
class TableProjections:
    """``lg (N,K) fp32 = tanh(table Wp^T) codes^T`` and ``tw (N,D) bf16 = table Wt^T``: the two nn.Linear layers of the path
    (reference model.py:171,174 and :212) applied once per table row instead of once per gathered row."""

    def __init__(self, table: torch.Tensor, lg: torch.Tensor, tw: Optional[torch.Tensor]):
        self.table, self.lg, self.tw = table, lg, tw


def table_project(table: torch.Tensor, weights: 'ScoreWeights', weighted: bool = True,
                  out: Optional[TableProjections] = None, workspace: Optional[torch.Tensor] = None) -> TableProjections:
    """Compute the table-level projections (``miner_table_project``).  ``out`` / ``workspace`` let a caller reuse buffers."""
    dev = _need_cuda(table)
    lib = L.load()
    if table.dtype != torch.bfloat16:
        raise L.MinerError('table-level mode needs a bfloat16 table')
    if weights.w_proj_bf16 is None:
        raise L.MinerError('table-level mode needs weights prepared with bf16 copies')
    table = table.detach().contiguous()
    N, D = table.shape
    K, Dc = weights.codes.shape
    want_tw = weighted and weights.w_target_bf16 is not None
    if out is None:
        lg = torch.empty(N, K, dtype=torch.float32, device=dev)
        tw = torch.empty(N, D, dtype=torch.bfloat16, device=dev) if want_tw else None
        out = TableProjections(table, lg, tw)
    ws_bytes = lib.miner_table_project_workspace_bytes(N, Dc)
    ws = workspace if (workspace is not None and workspace.numel() >= ws_bytes) else torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_table_project(_ptr(table), N, D, _ptr(weights.w_proj_bf16), _ptr(weights.codes),
                                        _ptr(weights.w_target_bf16) if want_tw else None, K, Dc, _ptr(out.lg),
                                        _ptr(out.tw) if want_tw else None, _ptr(ws), ws.numel(), _stream()))
    out.table = table
    return out


def score_table_supported(H: int, K: int, D: int) -> bool:
    return bool(L.load().miner_score_table_supported(H, K, D))


def score_table_geometry(H: int, K: int) -> Tuple[int, int]:
    """``(impressions per tile, 128-slot halves)`` of the table-level kernel for a shape (``miner_score_table_tile_geometry``)."""
    ipt, nh = C.c_int(), C.c_int()
    if not L.load().miner_score_table_tile_geometry(H, K, C.byref(ipt), C.byref(nh)):
        raise L.MinerError(f'table-level scoring does not cover H={H} K={K}')
    return ipt.value, nh.value


_WORKSPACES = weakref.WeakSet()      # live score_table workspaces: their out-of-range counters are checked by check_oob_all()


def score_table_workspace(B: int, H: int, K: int, device) -> torch.Tensor:
    """Workspace of :func:`score_table` (packed history tiles + the two out-of-range id counters, zeroed here)."""
    n = int(L.load().miner_score_table_workspace_bytes(B, H, K))
    ws = torch.empty(max(n, 256), dtype=torch.uint8, device=device)
    ws[:8].zero_()
    _WORKSPACES.add(ws)
    return ws


def oob_counts(workspace: torch.Tensor) -> torch.Tensor:
    """``(2,) int32`` view of a :func:`score_table` workspace: history ids / candidate ids outside the table seen so far."""
    return workspace[:8].view(torch.int32)


def check_oob(workspace: torch.Tensor) -> None:
    """Raise what the reference's table indexing raises (IndexError) if any call that used ``workspace`` saw an id outside the table."""
    n = oob_counts(workspace).tolist()
    if n[0] or n[1]:
        raise IndexError(f'index out of range in embedding table ({n[0]} history ids, {n[1]} candidate ids)')


def check_oob_all() -> None:
    """:func:`check_oob` over every live workspace (what the evaluators call once per ``compute_scores`` / ``evaluate``)."""
    for ws in list(_WORKSPACES):
        check_oob(ws)


def score_table(proj: TableProjections, his_ids: torch.Tensor, his_mask: torch.Tensor, cand_ids: torch.Tensor,
                score_type: str = 'weighted', cand_offsets: Optional[torch.Tensor] = None, bias_mean: Optional[torch.Tensor] = None,
                want_interests: bool = False, out_scores: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                check_bounds: bool = False):
    """Miner.forward (reference model.py:61-138) for a block of impressions from the table-level projections: one launch packs the
    histories into tiles (masked slots of the same news row merged), one fused kernel scores them.

    Dense layout: ``cand_ids`` (B,C).  CSR layout: ``cand_ids`` (T,) + ``cand_offsets`` (B+1,) int64.
    ``workspace`` (from :func:`score_table_workspace`) can be reused across calls; its out-of-range counters accumulate
    (:func:`check_oob`); ``check_bounds=True`` checks them here (one device sync).
    Returns ``(interests (B,K,D) or None, scores)``.
    """
    dev = _need_cuda(proj.table, his_ids, his_mask, cand_ids, cand_offsets, bias_mean)
    lib = L.load()
    st = L.SCORE_TYPES.get(score_type, -1)
    if st < 0:
        raise ValueError('Invalid method of aggregating matching score')
    B, H = his_ids.shape
    N, D = proj.table.shape
    K = proj.lg.shape[1]
    hid, it = _ids(his_ids)
    cid, it2 = _ids(cand_ids)
    if it != it2:
        cid, it2 = _ids(cand_ids.to(his_ids.dtype))
    m = _mask_u8(his_mask)
    if cand_offsets is None:
        Cn = cand_ids.shape[1]
        offs = None
        out_shape = (B, Cn)
    else:
        Cn = 0
        offs = cand_offsets.to(torch.int64).contiguous()
        out_shape = (cid.numel(),)
    bm = _f32(bias_mean) if bias_mean is not None else None
    scores = out_scores if out_scores is not None else torch.empty(out_shape, dtype=torch.float32, device=dev)
    interests = torch.empty(B, K, D, dtype=torch.float32, device=dev) if want_interests else None
    ws_bytes = int(lib.miner_score_table_workspace_bytes(B, H, K))
    ws = workspace if (workspace is not None and workspace.numel() >= ws_bytes) else score_table_workspace(B, H, K, dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_score_table_fwd(_ptr(proj.table), _ptr(proj.tw), _ptr(proj.lg), N, _ptr(hid), _ptr(m), _ptr(cid), _ptr(offs), it,
                                          _ptr(bm), B, H, Cn, K, D, st, _ptr(scores), _ptr(interests), _ptr(ws), ws.numel(), _stream()))
    if check_bounds:
        check_oob(ws)
    return interests, scores


def hist_interests(table: torch.Tensor, his_ids: torch.Tensor, his_mask: torch.Tensor, w_proj_bf16: torch.Tensor, codes: torch.Tensor,
                   bias_mean: Optional[torch.Tensor] = None, want_f32: bool = True):
    """History kernel of the fused tensor-core path on its own: ``(i_hi, i_lo, interests_f32 or None)``."""
    dev = _need_cuda(table, his_ids, his_mask, w_proj_bf16, codes, bias_mean)
    lib = L.load()
    assert table.dtype == torch.bfloat16 and w_proj_bf16.dtype == torch.bfloat16
    table, wp = table.contiguous(), w_proj_bf16.contiguous()
    B, H = his_ids.shape
    D = table.shape[1]
    K, Dc = codes.shape
    hid, it = _ids(his_ids)
    m, cd = _mask_u8(his_mask), _f32(codes)
    bm = _f32(bias_mean) if bias_mean is not None else None
    i_hi = torch.empty(B * K, D, dtype=torch.bfloat16, device=dev)
    i_lo = torch.empty(B * K, D, dtype=torch.bfloat16, device=dev)
    out = torch.empty(B, K, D, dtype=torch.float32, device=dev) if want_f32 else None
    ws_bytes = lib.miner_hist_interests_workspace_bytes(Dc)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_hist_interests_fwd(_ptr(table), table.shape[0], _ptr(hid), it, _ptr(m), _ptr(bm), _ptr(wp), _ptr(cd), B, H, K, Dc, D,
                                             _ptr(i_hi), _ptr(i_lo), _ptr(out), _ptr(ws), ws_bytes, _stream()))
    return i_hi, i_lo, out


def cand_score(i_hi: torch.Tensor, i_lo: torch.Tensor, w_target_bf16: torch.Tensor, table: torch.Tensor, cand_ids: torch.Tensor, K: int,
               cand_offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Candidate kernel of the fused tensor-core path on its own (score_type 'weighted')."""
    dev = _need_cuda(i_hi, i_lo, w_target_bf16, table, cand_ids, cand_offsets)
    lib = L.load()
    assert i_hi.dtype == torch.bfloat16 and i_lo.dtype == torch.bfloat16 and w_target_bf16.dtype == torch.bfloat16 and table.dtype == torch.bfloat16
    D = table.shape[1]
    B = i_hi.shape[0] // K
    cid, it = _ids(cand_ids)
    if cand_offsets is None:
        Cn = cand_ids.shape[1]
        offs = None
        out = torch.empty(B, Cn, dtype=torch.float32, device=dev)
    else:
        Cn = 0
        offs = cand_offsets.to(torch.int64).contiguous()
        out = torch.empty(cid.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_cand_score_fwd(_ptr(i_hi.contiguous()), _ptr(i_lo.contiguous()), _ptr(w_target_bf16.contiguous()), _ptr(table.contiguous()),
                                         table.shape[0], _ptr(cid), it, _ptr(offs), B, Cn, K, D, _ptr(out), _stream()))
    return out


def tc_gemm(a_bf16: torch.Tensor, b_bf16: torch.Tensor, epilogue: int = 0, a_ids: Optional[torch.Tensor] = None,
            want_bf16: bool = False):
    """The tcgen05 projection GEMM on its own: epi(A B^T) with optional row gather of A (tests / profiling)."""
    dev = _need_cuda(a_bf16, b_bf16, a_ids)
    assert a_bf16.dtype == torch.bfloat16 and b_bf16.dtype == torch.bfloat16
    a, b = a_bf16.contiguous(), b_bf16.contiguous()
    N, K = b.shape
    if a_ids is None:
        M, it, idp, rows = a.shape[0], L.I64, None, 0
        idc = None
    else:
        idc, it = _ids(a_ids)
        M, idp, rows = idc.numel(), _ptr(idc), a.shape[0]
    c = torch.empty(M, N, dtype=torch.float32, device=dev)
    cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    with torch.cuda.device(dev):
        L.check(L.load().miner_tc_gemm(_ptr(a), idp, it, rows, _ptr(b), _ptr(c), _ptr(cb), M, N, K, epilogue, _stream()))
    return (c, cb) if want_bf16 else c


def tc_gemm_tn(a_bf16: torch.Tensor, b_bf16: torch.Tensor, k_splits: int = 1) -> torch.Tensor:
    """The tcgen05 weight-gradient GEMM on its own: ``A^T B`` over the rows of ``A (R,M)`` and ``B (R,N)`` (both row-major bf16, read
    MN-major); returns the ``(k_splits, M, N)`` fp32 partials (tests / profiling)."""
    dev = _need_cuda(a_bf16, b_bf16)
    assert a_bf16.dtype == torch.bfloat16 and b_bf16.dtype == torch.bfloat16 and a_bf16.shape[0] == b_bf16.shape[0]
    a, b = a_bf16.contiguous(), b_bf16.contiguous()
    R, M = a.shape
    N = b.shape[1]
    c = torch.empty(k_splits, M, N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().miner_tc_gemm_tn(_ptr(a), _ptr(b), _ptr(c), R, M, N, k_splits, _stream()))
    return c


# ------------------------------------------------------------------------------------------------ (a7..a12)
TRANSFORMS = {'none': 0, 'sigmoid': 1, 'softmax': 2}


def metric_names(ks: Sequence[int]):
    return ['group_auc', 'mrr'] + [f'ndcg@{k}' for k in ks] + [f'hit@{k}' for k in ks]


def rank_metrics_raw(scores: torch.Tensor, labels: torch.Tensor, offsets: torch.Tensor, transform: str = 'sigmoid',
                     ks: Sequence[int] = (5, 10), per_impression: bool = False):
    """Returns ``(partials (2*M,) float64 device tensor [sum, count per metric], per_impression (B,M) or None)``."""
    dev = _need_cuda(scores, labels, offsets)
    lib = L.load()
    B = offsets.numel() - 1
    s = _f32(scores).reshape(-1)
    y = labels.to(torch.int8).contiguous().reshape(-1)
    o = offsets.to(torch.int64).contiguous()
    n_k = len(ks)
    M = 2 + 2 * n_k
    karr = (C.c_int * max(n_k, 1))(*[int(k) for k in ks])
    partials = torch.empty(2 * M, dtype=torch.float64, device=dev)
    per = torch.empty(B, M, dtype=torch.float64, device=dev) if per_impression else None
    ws_bytes = lib.miner_rank_metrics_workspace_bytes(B, n_k)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_rank_metrics(_ptr(s), _ptr(y), _ptr(o), B, TRANSFORMS[transform], karr, n_k, _ptr(partials), _ptr(per),
                                       _ptr(ws), ws_bytes, _stream()))
    return partials, per


def rank_metrics(scores, labels, offsets, transform: str = 'sigmoid', ks: Sequence[int] = (5, 10)) -> Dict[str, float]:
    """np.nanmean of every per-impression metric (reference evaluation.py:56-80)."""
    partials, _ = rank_metrics_raw(scores, labels, offsets, transform, ks)
    p = partials.cpu().view(-1, 2)
    return {n: (float(p[i, 0] / p[i, 1]) if p[i, 1] > 0 else float('nan')) for i, n in enumerate(metric_names(ks))}


# ------------------------------------------------------------------------------------------------ global auc (evaluation.py:53-55)
def auc_split(scores: torch.Tensor, labels: torch.Tensor, offsets: Optional[torch.Tensor], transform: str = 'sigmoid'):
    """Order-preserving uint32 keys of the candidates' probabilities, positives and negatives compacted apart.
    Returns ``(pos_keys (P,), neg_keys (N,))`` as int32-typed device tensors (the bits are unsigned keys); one device sync (P, N)."""
    dev = _need_cuda(scores, labels, offsets)
    lib = L.load()
    s = _f32(scores).reshape(-1)
    y = labels.to(torch.int8).contiguous().reshape(-1)
    T = s.numel()
    o = offsets.to(torch.int64).contiguous() if offsets is not None else None
    B = (o.numel() - 1) if o is not None else 0
    pos = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    neg = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_auc_split(_ptr(s), _ptr(y), _ptr(o), B, T, TRANSFORMS[transform], _ptr(pos), _ptr(neg), _ptr(counts), _stream()))
    P, N = (int(v) for v in counts.tolist())
    return pos[:P], neg[:N]


def sort_u32(keys: torch.Tensor) -> torch.Tensor:
    """LSD radix sort (ascending, as unsigned 32-bit keys) in place; returns ``keys``."""
    dev = _need_cuda(keys)
    lib = L.load()
    assert keys.dtype == torch.int32 and keys.is_contiguous()
    n = keys.numel()
    ws_bytes = int(lib.miner_sort_u32_workspace_bytes(n))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_sort_u32(_ptr(keys), n, _ptr(ws), ws.numel(), _stream()))
    return keys


def auc_count(pos_sorted: torch.Tensor, neg_keys: torch.Tensor) -> int:
    """``sum over negatives of 2 #{pos > neg} + #{pos == neg}`` (= 2 U of the Mann-Whitney statistic), an exact python int."""
    dev = _need_cuda(pos_sorted, neg_keys)
    out = torch.empty(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().miner_auc_count(_ptr(pos_sorted.contiguous()), pos_sorted.numel(), _ptr(neg_keys.contiguous()), neg_keys.numel(),
                                         _ptr(out), _stream()))
    return int(out.item())


# ------------------------------------------------------------------------------------------------ (a13, a14)
def loss_forward(interests: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor, eval_mode: bool = False) -> torch.Tensor:
    """Returns a (3,) device tensor ``[total, disagreement, rank_loss]`` (reference loss.py:27-44 / 68-85)."""
    dev = _need_cuda(interests, logits, labels)
    lib = L.load()
    B, K, D = interests.shape
    Cn = logits.shape[1]
    I, lg, lb = _f32(interests), _f32(logits), _f32(labels)
    out = torch.empty(3, dtype=torch.float32, device=dev)
    ws_bytes = lib.miner_loss_workspace_bytes(B, K)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_loss_fwd(_ptr(I), _ptr(lg), _ptr(lb), B, Cn, K, D, 1 if eval_mode else 0, _ptr(out), _ptr(ws), ws_bytes,
                                   _stream()))
    return out


# ------------------------------------------------------------------------------------------------ train variant (section 8 f1)
class TrainSaved:
    """Intermediates of :func:`train_forward` that :func:`train_backward` reads."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def train_forward(table: torch.Tensor, his_ids: torch.Tensor, his_mask: torch.Tensor, cand_ids: torch.Tensor,
                  w_proj: torch.Tensor, codes: torch.Tensor, w_target: Optional[torch.Tensor], math: int = L.MATH_FP32,
                  score_type: str = 'weighted', bias_mean: Optional[torch.Tensor] = None):
    """Miner.forward (reference model.py:61-138) keeping what the backward needs; any ``score_type``, optional category-bias
    scalar ``bias_mean`` (B,H) (model.py:176-177).

    Dense layout, ``cand_ids`` (B,C).  Returns ``(interests (B,K,D), scores (B,C), saved)``.  ``math=MATH_TENSOR`` runs the
    projection-sized GEMMs of the step on tcgen05 with bf16 operands (bf16 table, D % 64 == 0).
    """
    dev = _need_cuda(table, his_ids, his_mask, cand_ids, w_proj, codes, w_target, bias_mean)
    lib = L.load()
    st = L.SCORE_TYPES.get(score_type, -1)
    if st < 0:
        raise ValueError('Invalid method of aggregating matching score')
    weighted = st == L.SCORE_WEIGHTED
    table = table.detach().contiguous()
    B, H = his_ids.shape
    Cn = cand_ids.shape[1]
    D = table.shape[1]
    K, Dc = codes.shape
    hid, it = _ids(his_ids)
    cid, it2 = _ids(cand_ids)
    if it != it2:
        cid, it2 = _ids(cand_ids.to(his_ids.dtype))
    m = _mask_u8(his_mask)
    wp, cd = _f32(w_proj), _f32(codes)
    wt = _f32(w_target) if weighted else None
    bm = _f32(bias_mean) if bias_mean is not None else None
    f = dict(dtype=torch.float32, device=dev)
    interests, scores = torch.empty(B, K, D, **f), torch.empty(B, Cn, **f)
    t, w = torch.empty(B * H, Dc, **f), torch.empty(B, K, H, **f)
    z = torch.empty(B * K, D, **f) if weighted else None
    wp16 = cast_bf16(wp) if math == L.MATH_TENSOR else None
    wt16 = cast_bf16(wt) if (math == L.MATH_TENSOR and weighted) else None
    ws_bytes = lib.miner_train_workspace_bytes(B, H, K, Dc, D, math)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_train_fwd(_ptr(table), table.shape[0], _table_dtype(table), _ptr(hid), _ptr(m), _ptr(cid), it, _ptr(wp), _ptr(cd),
                                    _ptr(wt), B, H, Cn, K, Dc, D, _ptr(interests), _ptr(scores), _ptr(t), _ptr(w), _ptr(z), math,
                                    _ptr(wp16), _ptr(wt16), st, _ptr(bm), _ptr(ws), ws.numel(), _stream()))
    saved = TrainSaved(table=table, his_ids=hid, his_mask=m, cand_ids=cid, id_dtype=it, w_proj=wp, codes=cd, w_target=wt, t=t, w=w, z=z,
                       interests=interests, ws=ws, dims=(B, H, Cn, K, Dc, D), math=math, w_proj_bf16=wp16, w_target_bf16=wt16,
                       score_type=st, has_bias=bm is not None)
    return interests, scores, saved


def train_backward(saved: TrainSaved, d_scores: Optional[torch.Tensor], d_interests: Optional[torch.Tensor], want_table_grad: bool = False):
    """Gradients ``(grad_w_proj (Dc,D), grad_codes (K,Dc), grad_w_target (D,D) or None, d_bias_mean (B,H) or None,
    grad_table (N,D) fp32 or None)``."""
    lib = L.load()
    B, H, Cn, K, Dc, D = saved.dims
    dev = saved.table.device
    f = dict(dtype=torch.float32, device=dev)
    weighted = saved.score_type == L.SCORE_WEIGHTED
    ds = _f32(d_scores) if d_scores is not None else torch.zeros(B, Cn, **f)
    di = _f32(d_interests) if d_interests is not None else None
    gwp, gc = torch.empty(Dc, D, **f), torch.empty(K, Dc, **f)
    gwt = torch.empty(D, D, **f) if weighted else None
    dbias = torch.empty(B, H, **f) if saved.has_bias else None
    gtab = torch.zeros(saved.table.shape[0], D, **f) if want_table_grad else None
    tg_bytes = int(lib.miner_train_table_grad_workspace_bytes(B, H, Dc, D)) if want_table_grad else 0
    tg_ws = torch.empty(tg_bytes, dtype=torch.uint8, device=dev) if want_table_grad else None
    with torch.cuda.device(dev):
        L.check(lib.miner_train_bwd(_ptr(saved.table), saved.table.shape[0], _table_dtype(saved.table), _ptr(saved.his_ids),
                                    _ptr(saved.his_mask), _ptr(saved.cand_ids), saved.id_dtype, _ptr(saved.w_proj), _ptr(saved.codes),
                                    _ptr(saved.w_target), _ptr(saved.t), _ptr(saved.w), _ptr(saved.interests), _ptr(saved.z), _ptr(ds),
                                    _ptr(di), B, H, Cn, K, Dc, D, _ptr(gwp), _ptr(gc), _ptr(gwt), saved.math, _ptr(saved.w_proj_bf16),
                                    _ptr(saved.w_target_bf16), saved.score_type, _ptr(dbias), _ptr(gtab), _ptr(tg_ws), tg_bytes, _ptr(saved.ws),
                                    saved.ws.numel(), _stream()))
    return gwp, gc, gwt, dbias, gtab


def loss_backward(interests: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor, grad_out: Optional[torch.Tensor] = None):
    """Backward of Loss.compute (reference loss.py:27-44): ``(d_interests (B,K,D), d_logits (B,C))``."""
    dev = _need_cuda(interests, logits, labels, grad_out)
    lib = L.load()
    B, K, D = interests.shape
    Cn = logits.shape[1]
    I, lg, lb = _f32(interests), _f32(logits), _f32(labels)
    go = _f32(grad_out).reshape(1) if grad_out is not None else None
    di = torch.empty(B, K, D, dtype=torch.float32, device=dev)
    dl = torch.empty(B, Cn, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.miner_loss_bwd(_ptr(I), _ptr(lg), _ptr(lb), _ptr(go), B, Cn, K, D, _ptr(di), _ptr(dl), _stream()))
    return di, dl
