import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_layer(name):
    """Golden layer fixture -> `parts` dict in the oracle's numpy convention."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, f'layer_{name}.npz'))
    parts = dict(bits=int(z['bits']), qfn=str(z['qfn']), codes=z['codes'], scales=z['scales'],
                 zeros=z['zeros'], bias=z['bias'] if 'bias' in z else None,
                 scaleWH=z['scaleWH'] if 'scaleWH' in z else None, U=None, V=None)
    for side in 'UV':
        if f'{side}_B0' in z:
            parts[side] = ([z[f'{side}_B0'], z[f'{side}_B1']], z[f'{side}_p_in'], z[f'{side}_p_out'])
    return parts, z


LAYER_NAMES = ['l2b_incoh', 'l2b_incoh_rg', 'l3b_incoh', 'l4b_plain', 'l2b_kron', 'l4b_noperm',
               'l3b_rescale', 'l2b_qfna_proj']
