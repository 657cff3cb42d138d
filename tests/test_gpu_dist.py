"""Multi-rank path on real GPUs: tests/dist_check*.py under torchrun compare the slab-partitioned engine with the
unpartitioned problem (assembly 1e-12, SpMV with poisoned ghosts 1e-13, GMG-PCG, SNES state, total derivative).

  * on a 1-GPU box (the driver's test box) the ranks share cuda:0: the engine's own peer-memory transport
    (csrc/link.cuh) maps its windows between processes on the same device, so the whole distributed code path --
    slab layouts, halo exchange, fused all-reduces, partitioned multigrid, level gathers -- is exercised;
  * with >= 2 GPUs both transports run: link over NVLink peer memory and NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _modes():
    import torch
    if torch.cuda.device_count() < 2:
        return [dict(FEMO_DIST_SAME_DEVICE='1', FEMO_COMM='link')]
    return [dict(FEMO_COMM='link'), dict(FEMO_COMM='nccl')]


def _run(script, args, port, extra):
    R = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(R),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', script)] + [str(a) for a in args]
    # several distributed levels and the halo / interior overlap path even on these small meshes
    env = dict(os.environ, FEMO_DIST_MIN_ROWS='16', FEMO_OVERLAP_MIN_ROWS='256', **extra)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'OK' in out.stdout


@pytest.mark.parametrize('famid', [2, 1])
def test_slab_partition_matches_single_gpu(cuda_device, famid):
    for k, mode in enumerate(_modes()):
        _run('dist_check.py', [famid, 64, 128], 29611 + 10 * k + famid, mode)


def test_hex_slab_partition_matches_single_gpu(cuda_device):
    """z-slab partition of the hexahedral SIMP family vs the unpartitioned box."""
    for k, mode in enumerate(_modes()):
        _run('dist_check_hex.py', [16, 8, 32], 29641 + 10 * k, mode)


def test_femo_api_on_a_partitioned_mesh(cuda_device):
    """FEA + FEAModel + Simulator with one slab per rank (createUnitSquareMesh under femo_b200.dist): owned-dof arrays at
    the API, halo exchanges / all-reduces inside the engine; state, functional and adjoint totals vs one GPU."""
    for k, mode in enumerate(_modes()):
        _run('dist_check_api.py', [64], 29671 + 10 * k, mode)


@pytest.mark.parametrize('famid', [1, 2, 'motor'])
def test_unstructured_partition_matches_single_gpu(cuda_device, famid):
    """RCB partition of a perturbed, cell-shuffled triangle mesh (femo_b200/partition.py + femo_problem_set_partition): assembly,
    SpMV with poisoned ghosts, functional, AMG-preconditioned CG, Newton / SNES state and adjoint gradient vs the unpartitioned
    mesh; family 1 = Poisson with Dirichlet rows, family 2 = nonlinear Poisson with Nitsche terms on the true boundary facets,
    'motor' = nonlinear magnetostatics on the annulus (cell tags, boundary facets, GMRES with the distributed AMG)."""
    for k, mode in enumerate(_modes()):
        _run('dist_check_part.py', [48, famid], 29701 + 10 * k + (7 if famid == 'motor' else famid), mode)
