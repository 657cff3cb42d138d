import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    # the multi-process tests (torchrun, two ranks; on a 1-GPU box both ranks time-share cuda:0) run last: every
    # single-process parity test is through before the slowest, most environment-dependent ones start
    last = [it for it in items if 'test_gpu_dist' in it.nodeid]
    if last:
        items[:] = [it for it in items if 'test_gpu_dist' not in it.nodeid] + last


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    from femo_b200 import engine
    if not torch.cuda.is_available() or engine.device_count() == 0:
        pytest.fail('gpu-marked test selected but no CUDA device is visible (no CPU fallback exists)')
    return 0
