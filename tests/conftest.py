import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        data = np.load(GOLDEN / f'{name}.npz')
        return {
            k: torch.from_numpy(data[k]) if data[k].dtype.kind in 'fiu' else data[k]
            for k in data.files}
    return load


def relative_error(actual, expected):
    """max|a - b| / max|b|: the parity metric of BASELINE.json (<= 1e-4)"""
    actual = actual.detach().double().cpu()
    expected = expected.detach().double().cpu()
    return float((actual - expected).abs().max() / expected.abs().max().clamp_min(1e-30))
