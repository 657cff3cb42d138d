"""Target for ncu: MODE=step runs one warm-up and one timed-size state+adjoint
step of the bench workload; MODE=spmv runs only a few fine-level SpMVs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
n = int(os.environ.get('N', '4000'))
es = bench.EngineStep(n, 0)
if os.environ.get('MODE', 'step') == 'spmv':
    es.p.assemble_jacobian(plain=True, bc=False, out=es.vals)
    x = es.p.new_vector(es.p.N, 1.0); y = es.p.new_vector(es.p.N)
    for _ in range(8):
        es.p.spmv(0, es.vals, x, out=y)
else:
    es.step()
    torch.cuda.synchronize()
    print('LAUNCHES_BEFORE_TIMED_STEP', es.p.launch_count())
    es.step()
    print('LAUNCHES_AFTER', es.p.launch_count(), es.info)
torch.cuda.synchronize()
