"""Timing probe of the 3-D hexahedral SIMP family (SURVEY.md section 8d, C4-3D per-GPU size)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 64, 64))]
t0 = time.time()
mesh = E.EngineMesh.box_hex((0, 0, 0), (2.0 * nx, 2.0 * ny, 2.0 * nz), nx, ny, nz)
fc, fl = mesh.exterior_facets()
tag = np.nonzero(fl == 3)[0].astype(np.int32)              # traction on the whole x = hi face
p = E.EngineProblem(mesh, E.FAMILY_SIMP_HEX8, [0.3, 0.0, -0.25, 0.0, 3.0], tagged=tag)
nodes = np.arange((ny + 1) * (nz + 1)) * (nx + 1)          # x = 0 face
p.set_bc([np.stack([3 * nodes, 3 * nodes + 1, 3 * nodes + 2], axis=1).ravel().astype(np.int32)])
t1 = time.time()
lv = p.enable_multigrid()
t2 = time.time()
p.upload(0)
torch.cuda.synchronize()
t3 = time.time()
N, M = p.N, p.M[0]
nnz = p.pattern_info(0)['nnz']
print('cells %d dofs %d nnz %d levels %d | layout %.1fs mg %.1fs upload %.1fs | static %.2f GB work %.2f GB'
      % (M, N, nnz, lv, t1 - t0, t2 - t1, t3 - t2, p.static_bytes / 1e9, p.work_bytes / 1e9))
rng = np.random.default_rng(0)
u = p.new_vector(N, 0.0)
rho = p.to_device(np.clip(0.86 * rng.random(M), 1e-3, 1.0) if os.environ.get('RHO', 'random') == 'random' else np.full(M, 0.5))
p.set_coefficient(0, u); p.set_coefficient(1, rho)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

vals = p.new_vector(nnz); vals_bc = p.new_vector(nnz); R = p.new_vector(N)
print('residual  %.3f ms' % timeit(lambda: p.assemble_residual(R)))
ms = timeit(lambda: p.assemble_jacobian(out=vals, bc=True, out_bc=vals_bc))
print('jacobian (plain + bc) %.3f ms   scratch %.2f GB' % (ms, 576 * M * 8 / 1e9))
dv = p.new_vector(p.pattern_info(1)['nnz'])
print('dRdm      %.3f ms' % timeit(lambda: p.assemble_dRdm(0, dv)))
x = p.new_vector(N, 1.0); y = p.new_vector(N)
ms = timeit(lambda: p.spmv(0, vals, x, out=y), reps=20)
byt = 12 * nnz + 20 * N
print('spmv      %.3f ms  %.0f GB/s (algorithmic %.3f GB)' % (ms, byt / ms / 1e6, byt / 1e9))
b = p.assemble_residual()
b = -b
for pre in (2,):
    xx = p.new_vector(N, 0.0)
    torch.cuda.synchronize(); t = time.time()
    xx, info = p.linear_solve(vals_bc, b, xx, rtol=1e-8, max_it=2000, precond=pre)
    torch.cuda.synchronize(); dt = time.time() - t
    print('precond %d: %.1f ms' % (pre, dt * 1e3), info)
u.zero_()
torch.cuda.synchronize(); t = time.time()
info = p.newton_solve(kind='Newton', precond=2, krylov_rtol=1e-8)
torch.cuda.synchronize()
print('newton (3 fixed its) %.1f ms' % ((time.time() - t) * 1e3), info, 'compliance', p.assemble_output(1))
