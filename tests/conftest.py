import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files

    def keys(self, prefix=''):
        return [k for k in self.z.files if k.startswith(prefix)]

    def sub(self, prefix):
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}


@pytest.fixture(scope='session')
def g_ops():
    return Golden('ops.npz')


@pytest.fixture(scope='session')
def g_modules():
    return Golden('modules.npz')


@pytest.fixture(scope='session')
def g_model():
    return Golden('model.npz')


@pytest.fixture(scope='session')
def g_pl():
    return Golden('pl.npz')


@pytest.fixture(scope='session')
def g_sg3d():
    return Golden('sg3d.npz')


@pytest.fixture(scope='session')
def g_resample():
    return Golden('resample.npz')


def up_cases(g):
    """[(name, shape, filter taps or None, kwargs, wrapper)] as recorded by make_golden.py."""
    return [ast.literal_eval(str(s)) for s in g['up.cases']]


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative to the tensor's scale' error used for every fp tolerance."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
