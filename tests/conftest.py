import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


@pytest.fixture(scope='session')
def golden():
    return load_golden


def golden_model_inputs(g):
    """Tensors for a model golden; regenerates table/weights from the seed when they were not stored."""
    from miner_b200 import synth
    t = lambda k: torch.from_numpy(g[k])
    if 'table' in g:
        table, wp, codes, wt, cat = t('table'), t('w_proj'), t('codes'), t('w_target'), t('cat_emb')
    else:
        table = synth.make_table(int(g['N']), int(g['D']), int(g['seed']))
        w = synth.make_weights(int(g['D']), int(g['K']), int(g['Dc']), int(g['seed']), int(g['NC']), int(g['Ec']))
        wp, codes, wt, cat = w.w_proj, w.context_codes, w.w_target, w.cat_emb
        # the fixture pins the regenerated inputs through their checksums
        for key, ten in (('ck_table', table), ('ck_wp', wp), ('ck_codes', codes), ('ck_wt', wt), ('ck_cat', cat)):
            assert abs(float(ten.double().sum()) - float(g[key])) < 1e-9, key
    return dict(table=table, w_proj=wp, codes=codes, w_target=wt, cat_emb=cat, his_ids=t('his_ids'),
                his_mask=t('his_mask'), his_cat=t('his_cat'), cand=t('cand'), cand_cat=t('cand_cat'))
