import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """GPU tests must never pass silently on a machine without a GPU: skip them there, and on
    a GPU box make sure the native library is the one under test."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def refex_cases():
    with open(os.path.join(GOLDEN, 'refex_cases.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def refex_random():
    return np.load(os.path.join(GOLDEN, 'refex_random.npz'))


@pytest.fixture(scope='session')
def nmf_cases():
    return np.load(os.path.join(GOLDEN, 'nmf_cases.npz'))


@pytest.fixture(scope='session')
def prune_cases():
    with open(os.path.join(GOLDEN, 'prune_cases.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def roles_cases():
    with open(os.path.join(GOLDEN, 'roles_cases.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def prune_reference_tables():
    """The reference's own known-answer tables for the pruner (tests/test_features/
    test_prune.py), re-run through the unmodified reference by tests/golden/make_golden.py."""
    with open(os.path.join(GOLDEN, 'prune_reference_tables.json')) as f:
        return json.load(f)
