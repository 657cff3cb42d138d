"""Per-kernel timing probe on the bench workload (not the bench itself)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
t0 = time.time()
mesh = E.EngineMesh.unit_square(n)
p = E.EngineProblem(mesh, 2)
t1 = time.time()
p.upload(0)
torch.cuda.synchronize()
t2 = time.time()
print('layout build %.1fs upload %.1fs static %.2f GB work %.2f GB' % (t1 - t0, t2 - t1, p.static_bytes / 1e9, p.work_bytes / 1e9))
N, M = p.N, p.M[0]
nnz = p.pattern_info(0)['nnz']
u = p.new_vector(N, 0.0); f = p.new_vector(M, 0.1)
p.set_coefficient(0, u); p.set_coefficient(1, f)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

vals = p.new_vector(nnz); R = p.new_vector(N)
print('residual  %.3f ms' % timeit(lambda: p.assemble_residual(R)))
print('jacobian  %.3f ms' % timeit(lambda: p.assemble_jacobian(out=vals)))
dv = p.new_vector(p.pattern_info(1)['nnz'])
print('dRdm      %.3f ms' % timeit(lambda: p.assemble_dRdm(0, dv)))
print('output    %.3f ms' % timeit(lambda: p.assemble_output(0)))
g = p.new_vector(N)
print('output_du %.3f ms' % timeit(lambda: p.assemble_output_grad(0, 0, g)))
x = p.new_vector(N, 1.0); y = p.new_vector(N)
ms = timeit(lambda: p.spmv(0, vals, x, out=y), reps=20)
byt = 12 * nnz + 20 * N
print('spmv      %.3f ms  %.0f GB/s (algorithmic %.3f GB)' % (ms, byt / ms / 1e6, byt / 1e9))
ms = timeit(lambda: p.spmv(0, vals, x, transpose=True, out=y), reps=20)
print('spmv^T    %.3f ms' % ms)
b = p.assemble_residual()
for ce in (50,):
    xx = p.new_vector(N, 0.0)
    torch.cuda.synchronize(); t = time.time()
    xx, info = p.linear_solve(vals, b, xx, rtol=1e-30, max_it=500, check_every=ce)
    torch.cuda.synchronize(); dt = time.time() - t
    print('CG 500 its check_every=%d: %.3f ms/it' % (ce, dt * 1e3 / 500), info)
