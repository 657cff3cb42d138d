"""Multi-GPU path on real GPUs: runs tests/dist_check.py under torchrun when the box
has at least two devices (the driver's 1-GPU test box skips it; gpurun --gpus 2 runs it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('famid', [2, 1])
def test_slab_partition_matches_single_gpu(cuda_device, famid):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    R = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(R),
           '--master-addr', '127.0.0.1', '--master-port', '29611', os.path.join(ROOT, 'tests', 'dist_check.py'),
           str(famid), '64', str(64 * R)]
    env = dict(os.environ, FEMO_DIST_MIN_ROWS='16')     # several distributed levels even on this small mesh
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'OK' in out.stdout


def test_hex_slab_partition_matches_single_gpu(cuda_device):
    """z-slab partition of the hexahedral SIMP family vs the unpartitioned box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    R = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(R),
           '--master-addr', '127.0.0.1', '--master-port', '29613', os.path.join(ROOT, 'tests', 'dist_check_hex.py'),
           '16', '8', str(16 * R)]
    env = dict(os.environ, FEMO_DIST_MIN_ROWS='16')
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'OK' in out.stdout
