"""Generate golden vectors by running the UNMODIFIED reference (MrRobot2211/miner).

Run in the authoring container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

It imports the reference's own ``Miner``, ``PolyAttention``, ``TargetAwareAttention``
(src/model/model.py), ``pairwise_cosine_similarity`` (src/utils.py), ``Loss`` (src/loss.py)
and the evaluators / metric functions (src/evaluation.py), puts a table-lookup stub behind
the ``NewsEncoder`` call contract (``Miner`` only touches
``news_encoder.embed_dim`` and ``news_encoder(title_encoding=..)``), runs them on seeded
synthetic inputs and writes ``tests/golden/*.npz``.  The oracle (oracle/miner_oracle.py) and
the CUDA path are both checked against these files.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

from src.model.model import Miner, PolyAttention, TargetAwareAttention  # noqa: E402
from src.utils import pairwise_cosine_similarity                         # noqa: E402
from src.loss import Loss                                                # noqa: E402
from src import evaluation as ref_eval                                   # noqa: E402
from sklearn.metrics import roc_auc_score                                # noqa: E402

from miner_b200 import synth                                             # noqa: E402


class TableStub(nn.Module):
    """Stands in for NewsEncoder: same ``embed_dim`` property and forward keywords."""

    def __init__(self, table):
        super().__init__()
        self.table = table

    @property
    def embed_dim(self):
        return self.table.shape[1]

    def forward(self, title_encoding, title_attn_mask, sapo_encoding=None, sapo_attn_mask=None):
        return self.table[title_encoding[:, 0]]


def run_miner(model, his_ids, his_mask, cand_ids, category=None, his_category=None):
    B, C = cand_ids.shape
    H = his_ids.shape[1]
    z = torch.zeros(B, C, 1, dtype=torch.long)
    zh = torch.zeros(B, H, 1, dtype=torch.long)
    with torch.no_grad():
        return model(title=cand_ids[..., None], title_mask=z, his_title=his_ids[..., None], his_title_mask=zh,
                     his_mask=his_mask, sapo=z, sapo_mask=z, his_sapo=zh, his_sapo_mask=zh,
                     category=category, his_category=his_category)


def build_model(table, w: synth.Weights, score_type, use_bias, K, Dc):
    kw = {}
    if use_bias:
        kw = dict(num_category=w.cat_emb.shape[0], category_embed_dim=w.cat_emb.shape[1], category_pad_token_id=0)
    m = Miner(TableStub(table), use_bias, K, Dc, score_type, 0.2, **kw).eval()
    with torch.no_grad():
        m.poly_attn.linear.weight.copy_(w.w_proj)
        m.poly_attn.context_codes.copy_(w.context_codes)
        if score_type == 'weighted':
            m.target_aware_attn.linear.weight.copy_(w.w_target)
        if use_bias:
            m.category_embedding.weight.copy_(w.cat_emb)
    return m


def checksum(t: torch.Tensor) -> float:
    return float(t.double().sum().item())


def golden_model(name, N, D, H, K, Dc, C, B, NC, Ec, seed, store_inputs):
    table = synth.make_table(N, D, seed)
    w = synth.make_weights(D, K, Dc, seed, NC, Ec)
    g = torch.Generator().manual_seed(seed + 7)
    his_ids, his_mask, his_cat = synth.make_history(B, H, N, g, NC)
    cand = torch.randint(1, N + 1, (B, C), generator=g)
    cand_cat = torch.randint(1, NC, (B, C), generator=g)
    out = dict(N=N, D=D, H=H, K=K, Dc=Dc, C=C, B=B, NC=NC, Ec=Ec, seed=seed,
               his_ids=his_ids.numpy(), his_mask=his_mask.numpy(), his_cat=his_cat.numpy(),
               cand=cand.numpy(), cand_cat=cand_cat.numpy(),
               ck_table=checksum(table), ck_wp=checksum(w.w_proj), ck_codes=checksum(w.context_codes),
               ck_wt=checksum(w.w_target), ck_cat=checksum(w.cat_emb))
    if store_inputs:
        out.update(table=table.numpy(), w_proj=w.w_proj.numpy(), codes=w.context_codes.numpy(),
                   w_target=w.w_target.numpy(), cat_emb=w.cat_emb.numpy())
    for st in ('weighted', 'max', 'mean'):
        m = build_model(table, w, st, False, K, Dc)
        I, S = run_miner(m, his_ids, his_mask, cand)
        out[f'scores_{st}'] = S.numpy()
        if st == 'weighted':
            out['interests'] = I.numpy()
            # reference-faithful eval layout: one row per candidate (src/reader.py:376-379)
            hi = his_ids.repeat_interleave(C, 0)
            hm = his_mask.repeat_interleave(C, 0)
            _, S1 = run_miner(m, hi, hm, cand.reshape(-1, 1))
            out['scores_weighted_c1'] = S1.reshape(B, C).numpy()
            # bf16-valued table (the tensor-core path stores the table in bf16)
            tb = table.to(torch.bfloat16).float()
            mb = build_model(tb, w, st, False, K, Dc)
            Ib, Sb = run_miner(mb, his_ids, his_mask, cand)
            out['scores_weighted_bf16table'] = Sb.numpy()
            out['interests_bf16table'] = Ib.numpy()
            # op-level: PolyAttention / TargetAwareAttention called directly
            E = table[his_ids]
            with torch.no_grad():
                out['poly_direct'] = m.poly_attn(embeddings=E, attn_mask=his_mask, bias=None).numpy()
                cr = table[cand]
                match = torch.matmul(cr, I.permute(0, 2, 1))
                out['target_direct'] = m.target_aware_attn(query=I, key=cr, value=match).numpy()
    # category bias on (never set in a shipped config, but part of the API)
    mb = build_model(table, w, 'weighted', True, K, Dc)
    I, S = run_miner(mb, his_ids, his_mask, cand, category=cand_cat, his_category=his_cat)
    out['interests_bias'] = I.numpy()
    out['scores_bias'] = S.numpy()
    with torch.no_grad():
        out['category_bias'] = pairwise_cosine_similarity(mb.category_embedding(his_cat), mb.category_embedding(cand_cat)).numpy()
    # a candidate with the pad category NaNs its whole row
    cc = cand_cat.clone()
    cc[0, 0] = 0
    I, S = run_miner(mb, his_ids, his_mask, cand, category=cc, his_category=his_cat)
    out['scores_bias_padcand'] = S.numpy()
    out['interests_bias_padcand'] = I.numpy()
    # losses on the (B, C) train layout
    labels = torch.zeros(B, C, dtype=torch.long)
    labels[torch.arange(B), torch.randint(0, C, (B,), generator=g)] = 1
    m = build_model(table, w, 'weighted', False, K, Dc)
    m.train()
    z = torch.zeros(B, C, 1, dtype=torch.long)
    zh = torch.zeros(B, H, 1, dtype=torch.long)
    I, S = m(title=cand[..., None], title_mask=z, his_title=his_ids[..., None], his_title_mask=zh, his_mask=his_mask,
             sapo=z, sapo_mask=z, his_sapo=zh, his_sapo_mask=zh)
    loss = Loss(nn.CrossEntropyLoss(reduction='mean')).compute(I, S, labels)
    loss.backward()
    out['labels'] = labels.numpy()
    out['loss'] = loss.item()
    out['grad_w_proj'] = m.poly_attn.linear.weight.grad.numpy()
    out['grad_codes'] = m.poly_attn.context_codes.grad.numpy()
    out['grad_w_target'] = m.target_aware_attn.linear.weight.grad.numpy()
    with torch.no_grad():
        bl = (torch.rand(B, C, generator=g) < 0.3).float()
        out['eval_labels'] = bl.numpy()
        out['eval_loss'] = Loss.compute_eval_loss(I.detach(), S.detach(), bl)
        out['cosine_zero_diag'] = pairwise_cosine_similarity(I.detach(), I.detach(), zero_diagonal=True).numpy()
    if not store_inputs:
        for k in ('grad_w_proj', 'grad_w_target', 'interests_bias_padcand', 'interests_bf16table', 'poly_direct',
                  'interests_bias'):
            out.pop(k)
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items() if k.startswith('scores')})


def golden_metrics():
    rng = np.random.default_rng(36)
    n_imp = 200
    counts = np.clip(np.round(rng.lognormal(np.log(20), 0.5, n_imp)), 2, 300).astype(np.int64)
    counts[:4] = [2, 3, 16, 300]
    offsets = np.concatenate([[0], np.cumsum(counts)])
    T = int(offsets[-1])
    logits = (rng.standard_normal(T) * 0.35).astype(np.float32)
    labels = (rng.random(T) < 0.15).astype(np.int64)
    labels[offsets[:-1]] = 1
    labels[offsets[:-1] + 1] = 0
    imp_ids = np.repeat(np.arange(n_imp) * 3 + 5, counts)      # sparse, increasing ids
    # arrival order: shuffled batches, as a DataLoader(shuffle=False) over per-candidate samples would keep order;
    # we keep file order (reader.py:41-56) but feed in batches of 32 (eval_batch_size)
    samples = [SimpleNamespace(impression=SimpleNamespace(impression_id=int(i), label=[int(l)]))
               for i, l in zip(imp_ids, labels)]
    ev = ref_eval.SlowEvaluator(SimpleNamespace(samples=samples))
    lt = torch.from_numpy(logits)
    it = torch.from_numpy(imp_ids)
    for s in range(0, T, 32):
        ev.eval_batch(lt[s:s + 32].reshape(-1, 1), it[s:s + 32])
    metrics = ['auc', 'group_auc', 'mrr', 'ndcg@5', 'ndcg@10', 'hit@5', 'hit@10']
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        scores = ev.compute_scores(metrics, save_result=False)
    probs = torch.sigmoid(lt).double().numpy()
    per = {m: np.empty(n_imp) for m in metrics[1:]}
    for i in range(n_imp):
        t = labels[offsets[i]:offsets[i + 1]]
        p = probs[offsets[i]:offsets[i + 1]]
        per['group_auc'][i] = roc_auc_score(t, p)
        per['mrr'][i] = ref_eval.compute_mrr_score(np.array(t), np.array(p))
        for k in (5, 10):
            per[f'ndcg@{k}'][i] = ref_eval.compute_ndcg_score(np.array(t), np.array(p), k)
            per[f'hit@{k}'][i] = ref_eval.is_hit(np.array(t), np.array(p), k)
    # known answers quoted in SURVEY.md section 8c
    y = np.array([0, 1, 0, 0, 1, 0]); s = np.array([.1, .9, .3, .8, .2, .05])
    ka = dict(ka_y=y, ka_s=s, ka_mrr=ref_eval.compute_mrr_score(y, s), ka_ndcg5=ref_eval.compute_ndcg_score(y, s, 5),
              ka_ndcg10=ref_eval.compute_ndcg_score(y, s, 10), ka_hit5=ref_eval.is_hit(y, s, 5),
              ka_auc=roc_auc_score(y, s))
    # ties (n <= 16: numpy's argsort is an insertion sort, i.e. stable)
    ty = np.array([0, 1, 0, 1, 0, 0, 1, 0]); ts = np.array([.5, .5, .5, .2, .9, .2, .9, .1])
    tie = dict(tie_y=ty, tie_s=ts, tie_mrr=ref_eval.compute_mrr_score(ty, ts), tie_ndcg5=ref_eval.compute_ndcg_score(ty, ts, 5),
               tie_hit1=ref_eval.is_hit(ty, ts, 1), tie_hit2=ref_eval.is_hit(ty, ts, 2), tie_auc=roc_auc_score(ty, ts),
               eq_mrr=ref_eval.compute_mrr_score(np.array([0, 1, 0, 0, 1, 0]), np.full(6, .3)),
               eq_auc=roc_auc_score(np.array([0, 1, 0, 0, 1, 0]), np.full(6, .3)))
    with np.errstate(invalid='ignore'):
        nopos_mrr = ref_eval.compute_mrr_score(np.zeros(4, dtype=np.int64), np.array([.1, .2, .3, .4]))
    # FastEvaluator (softmax over npratio+1) on a (B,5) block
    fl = (rng.standard_normal((64, 5)) * 0.5).astype(np.float32)
    fy = np.zeros((64, 5), dtype=np.int64); fy[np.arange(64), rng.integers(0, 5, 64)] = 1
    fsamples = [SimpleNamespace(impression=SimpleNamespace(impression_id=i, label=[int(v) for v in fy[i]])) for i in range(64)]
    fev = ref_eval.FastEvaluator(SimpleNamespace(samples=fsamples))
    fev.eval_batch(torch.from_numpy(fl), torch.arange(64))
    with contextlib.redirect_stdout(io.StringIO()):
        fscores = fev.compute_scores(metrics, save_result=False)
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), logits=logits, labels=labels, offsets=offsets, imp_ids=imp_ids,
                        probs=probs, nopos_mrr=nopos_mrr, fast_logits=fl, fast_labels=fy,
                        **{f'agg_{k}': v for k, v in scores.items()}, **{f'per_{k}': v for k, v in per.items()},
                        **{f'fast_{k}': v for k, v in fscores.items()}, **ka, **tie)
    print('metrics', scores, fscores, ka['ka_mrr'], ka['ka_ndcg5'], tie['eq_mrr'])


def golden_fastformer():
    """BASELINE configs[4]: the reference's FastFormer (src/model/model.py:223-341, hidden size 256) behind the table stub, with
    deterministic weights (synth.deterministic_state) so the fixture holds only inputs and outputs."""
    from src.model.model import FastFormer
    N, D, H, C, B, seed = 300, 256, 50, 5, 6, 36
    table = synth.make_table(N, D, seed)
    m = FastFormer(TableStub(table), 'weighted', 0.2).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth.deterministic_state(shapes, seed))
    g = torch.Generator().manual_seed(seed + 5)
    his_ids, his_mask, _ = synth.make_history(B, H, N, g)
    his_mask[0] = True                                             # one full history
    his_ids[0] = torch.randint(1, N + 1, (H,), generator=g)
    cand = torch.randint(1, N + 1, (B, C), generator=g)
    scores = run_miner(m, his_ids, his_mask, cand)
    with torch.no_grad():
        user = m.fast_attn(input_embs=table[his_ids], attention_mask=his_mask)
    np.savez_compressed(os.path.join(HERE, 'fastformer.npz'), his_ids=his_ids.numpy(), his_mask=his_mask.numpy(), cand=cand.numpy(),
                        scores=scores.numpy(), user=user.numpy(), keys=np.array(sorted(shapes)),
                        shapes=np.array([str(shapes[k]) for k in sorted(shapes)]), dims=np.array([N, D, H, C, B, seed]))
    print('fastformer', scores.shape, float(scores.abs().max()), len(shapes), 'tensors')


if __name__ == '__main__':
    torch.manual_seed(36)
    torch.set_num_threads(1)     # single-thread MKL: deterministic summation order in the fixtures
    golden_model('model_small', N=64, D=64, H=12, K=8, Dc=24, C=5, B=6, NC=7, Ec=10, seed=36, store_inputs=True)
    golden_model('model_odd', N=50, D=40, H=7, K=5, Dc=9, C=3, B=5, NC=5, Ec=6, seed=11, store_inputs=True)
    golden_model('model_full', N=400, D=768, H=50, K=32, Dc=200, C=20, B=8, NC=20, Ec=100, seed=36, store_inputs=False)
    golden_metrics()
    golden_fastformer()
