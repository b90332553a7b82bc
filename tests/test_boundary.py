"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the Python mirror keeps the reference's names, the product never touches the oracle, and the host-side sharding /
reduction logic (world_size 2 over gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'miner_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(miner_[a-z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_every_declared_symbol():
    from miner_b200 import _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/miner_b200.h but not exported'
    assert sorted(_lib.EXPORTS) == declared
    assert lib.miner_abi_version() == _lib.ABI_VERSION


def test_built_for_sm100a_with_blackwell_instructions():
    from miner_b200 import _lib
    _lib.load()
    out = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out.stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM', 'LDGSTS'):
        assert mnemonic in out.stdout, f'{mnemonic} missing from SASS'


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, 'miner_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'miner_oracle' not in src.replace(
                    'oracle/miner_oracle.py', ''), f


def test_reference_api_names_and_state_dict_keys():
    import inspect
    import miner_b200 as mb
    enc = mb.TableNewsEncoder(torch.zeros(5, 64))
    m = mb.Miner(enc, True, 8, 24, 'weighted', 0.2, num_category=7, category_embed_dim=10, category_pad_token_id=0)
    keys = set(m.state_dict().keys())
    # parameter names a reference checkpoint carries (checkpoint row)
    assert {'poly_attn.linear.weight', 'poly_attn.context_codes', 'target_aware_attn.linear.weight',
            'category_embedding.weight'} <= keys
    assert m.poly_attn.linear.weight.shape == (24, 64) and m.poly_attn.context_codes.shape == (8, 24)
    assert m.target_aware_attn.linear.weight.shape == (64, 64)
    sig = inspect.signature(mb.Miner.forward)
    assert list(sig.parameters)[1:] == ['title', 'title_mask', 'his_title', 'his_title_mask', 'his_mask', 'sapo', 'sapo_mask',
                                        'his_sapo', 'his_sapo_mask', 'category', 'his_category']
    assert list(inspect.signature(mb.Miner.__init__).parameters)[1:] == [
        'news_encoder', 'use_category_bias', 'num_context_codes', 'context_code_dim', 'score_type', 'dropout', 'num_category',
        'category_embed_dim', 'category_pad_token_id', 'category_embed']
    assert list(inspect.signature(mb.PolyAttention.forward).parameters)[1:] == ['embeddings', 'attn_mask', 'bias']
    assert list(inspect.signature(mb.TargetAwareAttention.forward).parameters)[1:] == ['query', 'key', 'value']
    assert enc.embed_dim == 64
    with pytest.raises(AssertionError):
        mb.Miner(enc, True, 8, 24, 'weighted', 0.2)            # assert num_category is not None (model.py:49)


def test_cpu_tensors_fail_loudly():
    import miner_b200 as mb
    from miner_b200 import ops, _lib
    with pytest.raises(_lib.MinerError, match='no CPU fallback'):
        ops.gather(torch.zeros(4, 8), torch.zeros(3, dtype=torch.int64))
    enc = mb.TableNewsEncoder(torch.zeros(5, 64))
    m = mb.Miner(enc, False, 8, 24, 'weighted', 0.2).eval()
    z = torch.zeros(2, 3, 1, dtype=torch.long)
    zh = torch.zeros(2, 4, 1, dtype=torch.long)
    with pytest.raises(_lib.MinerError):
        m(z, z, zh, zh, torch.ones(2, 4, dtype=torch.bool), z, z, zh, zh)
    m.score_type = 'median'
    with pytest.raises(ValueError, match='Invalid method of aggregating matching score'):
        m(z, z, zh, zh, torch.ones(2, 4, dtype=torch.bool), z, z, zh, zh)


def test_shard_bounds_balance_and_cover():
    from miner_b200.parallel import shard_bounds, local_shard
    g = torch.Generator().manual_seed(0)
    counts = torch.randint(2, 300, (1000,), generator=g)
    offs = torch.zeros(1001, dtype=torch.int64)
    offs[1:] = torch.cumsum(counts, 0)
    for ws in (1, 2, 4, 8):
        b = shard_bounds(offs, ws, his_len=50)
        assert b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
        cost = [int(offs[e] - offs[s]) + 50 * (e - s) for s, e in b]
        assert max(cost) - min(cost) <= 4 * (350 + 50), cost
        assert all(s % 4 == 0 for s, _ in b)                  # shards start on a tile boundary of the table-level kernel
        assert shard_bounds(offs, ws, his_len=50, align=1)[0][0] == 0
    s, e, loc = local_shard(offs, 1, 4, 50)
    assert loc[0] == 0 and loc.numel() == e - s + 1 and int(loc[-1]) == int(offs[e] - offs[s])
    assert shard_bounds(torch.zeros(1, dtype=torch.int64), 4) == [(0, 0)] * 4


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from miner_b200.parallel import shard_bounds, allreduce_partials, finalize_metrics
from oracle import miner_oracle as O
import numpy as np
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
g = np.load(os.path.join(sys.argv[1], 'tests', 'golden', 'metrics.npz'))
offs = torch.from_numpy(g['offsets'])
s, e = shard_bounds(offs, world, 50)[rank]
names = ['group_auc', 'mrr', 'ndcg@5', 'ndcg@10', 'hit@5', 'hit@10']
per = O.per_impression_metrics(g['labels'], g['probs'], g['offsets'][s:e + 1])   # CPU stand-in for the per-rank kernel output
part = torch.zeros(2 * len(names), dtype=torch.float64)
for i, n in enumerate(names):
    v = per[n]
    part[2 * i] = np.nansum(v); part[2 * i + 1] = np.count_nonzero(~np.isnan(v))
allreduce_partials(part)
res = finalize_metrics(part, names)
if rank == 0:
    for n in names:
        assert abs(res[n] - float(g['agg_' + n])) < 1e-12, (n, res[n], float(g['agg_' + n]))
    print('OK')
dist.destroy_process_group()
'''


def test_two_rank_metric_reduction_gloo(tmp_path):
    """world_size 2 over gloo: sharded per-impression metrics + all-reduce of [sum,count] equal the single-process means."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29611', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'OK' in outs[0]


GRAD_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from miner_b200.parallel import allreduce_gradients, FlatGradients
from miner_b200 import synth
from oracle import miner_oracle as O
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
B, H, N, D, K, Dc = 8, 6, 40, 16, 4, 8
table = synth.make_table(N, D, 3)
w = synth.make_weights(D, K, Dc, 3)
his, mask, _, cand, _, labels = synth.make_train_batch(B, H, N, 4, 3)

def grads(sl):
    ps = [t.clone().requires_grad_(True) for t in (w.w_proj, w.context_codes, w.w_target)]
    I, S = O.miner_forward(table, his[sl], mask[sl], cand[sl], ps[0], ps[1], ps[2], 'weighted')
    O.loss_compute(I, S, labels[sl].float()).backward()
    return ps

half = B // world
local = grads(slice(rank * half, (rank + 1) * half))        # CPU stand-in for the per-rank CUDA backward
allreduce_gradients(local)
full = grads(slice(0, B))
for a, b in zip(local, full):
    assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-6), float((a.grad - b.grad).abs().max())
# the same through FlatGradients: the three gradients are views of one buffer, one in-place all-reduce
ps = [torch.nn.Parameter(t.clone()) for t in (w.w_proj, w.context_codes, w.w_target)]
fg = FlatGradients(ps)
for _ in range(2):                                           # second round: zero() keeps the views, autograd accumulates in place
    fg.zero()
    sl = slice(rank * half, (rank + 1) * half)
    I, S = O.miner_forward(table, his[sl], mask[sl], cand[sl], ps[0], ps[1], ps[2], 'weighted')
    O.loss_compute(I, S, labels[sl].float()).backward()
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() and p.grad.data_ptr() < fg.flat.data_ptr() + 4 * fg.flat.numel() for p in ps)
    fg.allreduce()
    for a, b in zip(ps, full):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-6), float((a.grad - b.grad).abs().max())
if rank == 0:
    print('OK')
dist.destroy_process_group()
'''


def test_two_rank_gradient_averaging_gloo(tmp_path):
    """world_size 2 over gloo: per-rank gradients of equal local batches, averaged with one flat all-reduce, equal the gradient
    of the loss over the global batch (SURVEY.md section 8e caveat)."""
    script = tmp_path / 'grad_worker.py'
    script.write_text(GRAD_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29613', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'OK' in outs[0]


AUC_WORKER = r'''
import os, sys, torch, torch.distributed as dist, numpy as np
sys.path.insert(0, sys.argv[1])
from miner_b200.evaluation import global_auc
from miner_b200.parallel import shard_bounds
from oracle import miner_oracle as O
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
g = np.load(os.path.join(sys.argv[1], 'tests', 'golden', 'metrics.npz'))
offs = torch.from_numpy(g['offsets'])

# CPU stand-ins for the three device calls (same contracts: uint32 order keys in int32 storage, ascending sort, 2U count)
def key(p):
    b = p.astype(np.float32).view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
def split(scores, labels, offsets, transform):
    k = key(scores.numpy()); y = labels.numpy()
    return torch.from_numpy(k[y > 0].view(np.int32).copy()), torch.from_numpy(k[y <= 0].view(np.int32).copy())
def sort(t):
    return torch.from_numpy(np.sort(t.numpy().view(np.uint32)).view(np.int32).copy())
def count(pos, neg):
    p, n = pos.numpy().view(np.uint32), neg.numpy().view(np.uint32)
    lb, ub = np.searchsorted(p, n, 'left'), np.searchsorted(p, n, 'right')
    return int((2 * (len(p) - ub) + (ub - lb)).sum())

s, e = shard_bounds(offs, world, 50)[rank]
c0, c1 = int(offs[s]), int(offs[e])
probs, labels = torch.from_numpy(g['probs'].astype(np.float32)), torch.from_numpy(g['labels'].astype(np.int8))
auc = global_auc(probs[c0:c1], labels[c0:c1], None, 'none', _kernels=(split, sort, count))
full = O.auc_score(g['labels'], g['probs'].astype(np.float32).astype(np.float64))
assert abs(auc - full) < 1e-12, (auc, full)
assert abs(auc - float(g['agg_auc'])) < 1e-7
if rank == 0:
    print('OK')
dist.destroy_process_group()
'''


def test_two_rank_global_auc_gloo(tmp_path):
    """world_size 2 over gloo: every rank holds a shard of the candidates; all-gather of the positive keys + int64 all-reduce of
    [2U, N] give the single-process auc (the device calls replaced by numpy stand-ins with the same contracts)."""
    script = tmp_path / 'auc_worker.py'
    script.write_text(AUC_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29617', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'OK' in outs[0]


def test_table_level_shape_query_is_host_only():
    """`miner_score_table_supported` is a host-side shape query (no GPU): the table-level kernel covers H <= 256, K <= 64,
    D a multiple of 64; everything else falls back to the reference-order families."""
    from miner_b200 import _lib
    lib = _lib.load()
    ok = lambda h, k, d: bool(lib.miner_score_table_supported(h, k, d))
    assert ok(50, 32, 768) and ok(1, 1, 64) and ok(128, 64, 256) and ok(200, 32, 768) and ok(256, 64, 64)
    assert not ok(257, 32, 768) and not ok(50, 65, 768) and not ok(50, 32, 100) and not ok(0, 32, 768)
    # tiling of a shape: impressions per tile (2 K TMEM lanes each) and 128-slot halves; workspace = 256 B of counters + headers + records
    from miner_b200 import ops
    assert ops.score_table_geometry(50, 32) == (2, 1) and ops.score_table_geometry(50, 8) == (2, 1) and ops.score_table_geometry(32, 16) == (4, 1)
    assert ops.score_table_geometry(100, 32) == (2, 2) and ops.score_table_geometry(200, 32) == (1, 2) and ops.score_table_geometry(50, 64) == (1, 1)
    assert ops.score_table_geometry(128, 64) == (1, 1) and ops.score_table_geometry(129, 64) == (1, 2) and ops.score_table_geometry(64, 32) == (2, 1)
    assert ops.score_table_geometry(65, 32) == (2, 2)
    ws = lib.miner_score_table_workspace_bytes
    assert ws(1000, 50, 32) == 256 + 4096 + 500 * 128 * 8 and ws(3, 200, 32) == 256 + 256 + 3 * 256 * 8 and ws(0, 50, 32) == 256


def test_default_math_selection_is_host_logic():
    """Which kernel family grouped evaluation picks depends only on dtype and shape (no GPU needed to decide)."""
    from miner_b200 import ops, _lib
    bf = torch.zeros(4, 768, dtype=torch.bfloat16)
    assert ops.default_eval_math(bf, 50, 32) == _lib.MATH_TABLE
    assert ops.default_eval_math(bf, 200, 64) == _lib.MATH_TABLE
    assert ops.default_eval_math(bf, 300, 32) == _lib.MATH_TENSOR                     # history beyond the 256-slot tile: reference order on tcgen05
    assert ops.default_eval_math(torch.zeros(4, 768), 50, 32) == _lib.MATH_FP32       # fp32 table: reference arithmetic
    assert ops.default_eval_math(torch.zeros(4, 100, dtype=torch.bfloat16), 50, 32) == _lib.MATH_FP32   # D % 64 != 0
    assert ops.default_math(bf, 768) == _lib.MATH_TENSOR                              # Miner.forward keeps the reference operation order
    class _Shape:                                                                     # a 3M-row table (4.6 GB) without allocating it
        dtype, shape = torch.bfloat16, (3_000_000, 768)
    assert ops.default_eval_math(_Shape, 50, 32) == _lib.MATH_TENSOR                  # beyond the kernel's 32-bit row offsets: reference order


def test_host_evaluator_wave_schedule_is_host_logic():
    """The wave schedule of the host pipeline (small first wave, doubling up to a cap) covers every impression exactly once and cuts on
    multiples of 4 impressions (tile boundaries of the table-level kernel), for any batch size."""
    from miner_b200.pipeline import wave_bounds
    assert wave_bounds(0, 16384, 2.0, 262144) == [0]
    b = wave_bounds(1_000_000, 16384, 2.0, 262144)
    assert b[0] == 0 and b[-1] == 1_000_000 and b[1] == 16384 and b[2] == 16384 + 32768
    sizes = [y - x for x, y in zip(b[:-1], b[1:])]
    assert max(sizes) == 262144 and len(sizes) == 7
    for B, first, growth, cap in ((1501, 128, 2.0, 2048), (333, 83, 1.5, 400), (7, 4, 1.0, 4), (100, 1000, 3.0, 10)):
        b = wave_bounds(B, first, growth, cap)
        assert b[0] == 0 and b[-1] == B and all(y > x for x, y in zip(b[:-1], b[1:]))
        assert all(x % 4 == 0 for x in b[:-1])
    assert wave_bounds(4096, 1024, 1.0, 1024) == [0, 1024, 2048, 3072, 4096]          # equal waves


def test_bind_to_gpu_cpus_without_nvml_changes_nothing():
    import os
    from miner_b200 import parallel
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_gpu_cpus(0) is None or os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
