"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm's line (run here at a small n) and
the `config` both arms print for the same flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_line(*flags, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--n', '96'] + list(flags),
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().split('\n')[-1])


def test_reference_arm_line():
    d = _ref_line('--steps', '2', '--warmup', '2', env={'OMP_NUM_THREADS': '1'})      # what torchrun exports to its workers
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['steps'] == 2 and d['warmup'] == 2 and d['higher_is_better'] is True
    assert d['metric'] == 'state+adjoint solves/s' and d['unit'] == 'solves/s' and d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert abs(d['value'] * d['ms_per_step'] / 1e3 - 1.0) < 1e-9
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and cb['unit'] == d['unit'] and 'sample' in cb
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count()
    assert cb['cores'] == cores                                   # not the single thread OMP_NUM_THREADS=1 asked for
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_both_arms_share_one_config():
    sys.path.insert(0, ROOT)
    import bench as B
    for world in (1, 2, 8):
        d = _ref_line('--steps', '1', '--warmup', '1', '--gpus', str(world))
        assert d['n_gpus'] == world and d['config'] == B.arm_config(96, 'p1', world)
    c1, c8 = B.arm_config(4000, 'p1', 1), B.arm_config(4000, 'p1', 8)
    assert c1['dofs'] == 16008001 and c1['parallelism'] == '1 GPU' and 'model' not in c1
    assert c8['dofs'] == 4097 * (8 * 4096 + 1) and '8 GPUs' in c8['workload']
    assert B.arm_config(256, 'hex', 8)['dofs'] == 51022467        # C4-3D
