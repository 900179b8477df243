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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: d[k] for k in d.files}


def sub(d, prefix, device=None):
    """{'0.weight': tensor, ...} for keys 'prefix/0.weight' of a golden dict."""
    out = {k[len(prefix) + 1:]: torch.from_numpy(d[k]) for k in d if k.startswith(prefix + '/')}
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


@pytest.fixture(scope='session')
def golden_decoder():
    return load_golden('decoder_solve.npz')


@pytest.fixture(scope='session')
def golden_encoder():
    return load_golden('encoder_loop.npz')


@pytest.fixture(scope='session')
def golden_schedule():
    return load_golden('schedule.npz')


@pytest.fixture(scope='session')
def golden_stage():
    return load_golden('decoder_stage.npz')


@pytest.fixture(scope='session')
def golden_encoder_stage():
    return load_golden('encoder_stage.npz')
