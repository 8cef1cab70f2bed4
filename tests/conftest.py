import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


# the tests stay inside the repo: no kernel cache in the user's home (the disk-cache test names its own)
os.environ.setdefault('BSI_CACHE_DIR', 'off')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        # a kernel that never returns must not eat the GPU box's time limit: the thread method
        # ends the process even while the main thread sits in a driver call
        if config.pluginmanager.hasplugin('timeout'):
            for item in items:
                if 'gpu' in item.keywords and item.get_closest_marker('timeout') is None:
                    item.add_marker(pytest.mark.timeout(240, method='thread'))
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
