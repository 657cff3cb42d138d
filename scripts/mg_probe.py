"""GMG-PCG convergence / timing probe."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
fam = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.time()
p = E.EngineProblem(E.EngineMesh.unit_square(n), fam)
lv = p.enable_multigrid()
t1 = time.time()
if fam == 1:
    x = p.mesh.coords()
    lists = [np.nonzero(np.isclose(x[:, a], b, atol=1e-6))[0] for a, b in ((0, 0.), (0, 1.), (1, 0.), (1, 1.))]
    p.set_bc(lists)
p.upload(0)
torch.cuda.synchronize()
print('n=%d levels=%d layout %.1fs upload %.1fs static %.2f GB work %.2f GB' % (n, lv, t1 - t0, time.time() - t1, p.static_bytes / 1e9, p.work_bytes / 1e9))
N, M = p.N, p.M[0]
u = p.new_vector(N, 0.0); f = p.new_vector(M, 0.1)
p.set_coefficient(0, u); p.set_coefficient(1, f)
if fam == 1:
    p.set_coefficient(2, p.new_vector(N, 0.0))
vals, vals_bc = p.assemble_jacobian(plain=True, bc=(fam == 1))
v = vals_bc if fam == 1 else vals
b = p.assemble_residual()
import itertools
for deg, ratio in ((1, 2.0), (1, 4.0), (2, 2.0), (2, 3.0), (2, 4.0), (2, 6.0), (3, 6.0)):
    for rep in range(2):
        x = p.new_vector(N, 0.0)
        torch.cuda.synchronize(); t = time.time()
        x, info = p.linear_solve(v, b, x, rtol=1e-10, precond=2, cheb_degree=deg, cheb_ratio=ratio, max_it=200)
        torch.cuda.synchronize(); dt = time.time() - t
    print('deg %d ratio %4.0f: %.1f ms  its %d' % (deg, ratio, dt * 1e3, info['iterations']))
u.zero_()
torch.cuda.synchronize(); t = time.time()
info = p.newton_solve(kind='SNES', krylov_rtol=1e-10, precond=2)
torch.cuda.synchronize()
print('SNES solve %.1f ms' % ((time.time() - t) * 1e3), info)
