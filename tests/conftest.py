import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / 'tests' / 'golden'
for p in (ROOT / 'nerf-attention_b200', ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def golden():
    """Outputs of the real reference, produced by tests/golden/make_golden.py."""
    meta = json.loads((GOLDEN / 'golden.json').read_text())
    return {
        'meta': meta['cases'],
        'torch': meta['torch'],
        'forward': np.load(GOLDEN / 'forward.npz'),
        'fits': np.load(GOLDEN / 'fits.npz'),
        'synthetic': np.load(GOLDEN / 'synthetic.npz'),
    }


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return 'cuda'
