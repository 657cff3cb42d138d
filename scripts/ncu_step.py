"""Target for ncu.  MODE=step: one warm-up and one timed-size state+adjoint step of the bench workload;
MODE=spmv: a few fine-level CSR SpMVs; MODE=probe: after one step (hierarchy set up), three fine-level launches of each
V-cycle operator instantiation (modes 0-3) and of the fp64 DIA SpMV of the CG recurrence (modes 4-5) between cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
n = int(os.environ.get('N', '4000'))
es = bench.EngineStep(n, 0)
mode = os.environ.get('MODE', 'step')
if mode == 'spmv':
    es.p.assemble_jacobian(plain=True, bc=False, out=es.vals)
    x = es.p.new_vector(es.p.N, 1.0); y = es.p.new_vector(es.p.N)
    for _ in range(8):
        es.p.spmv(0, es.vals, x, out=y)
elif mode == 'probe':
    es.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for m in range(6):
        for _ in range(3):
            es.p.vcycle_op_probe(m)
    x = es.p.new_vector(es.p.N, 1.0); y = es.p.new_vector(es.p.N)
    for _ in range(3):
        es.p.spmv(0, es.vals, x, out=y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    es.step()
    torch.cuda.synchronize()
    print('LAUNCHES_BEFORE_TIMED_STEP', es.p.launch_count())
    es.step()
    print('LAUNCHES_AFTER', es.p.launch_count(), es.info)
torch.cuda.synchronize()
