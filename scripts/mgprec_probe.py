"""fp32-valued vs fp64-valued V-cycle: iterations and time of the GMG-PCG solve (bench workload and hex)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E


def run(p, vals, b, tag):
    for rtol in ((1e-10,) if os.environ.get('SKIP_HEX') else (1e-10, 1e-13)):
        for prec in (1, 0):
            best = 1e9
            for _ in range(3):
                x = p.new_vector(p.N, 0.0)
                torch.cuda.synchronize(); t = time.perf_counter()
                x, info = p.linear_solve(vals, b, x, rtol=rtol, max_it=500, precond=2, mg_precision=prec)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
            r = b - p.spmv(0, vals, x)
            print('%s rtol %.0e %s: %2d its %.2f ms  true relres %.2e' % (tag, rtol, 'fp64' if prec else 'fp32', info['iterations'],
                                                                      best * 1e3, float(r.norm() / b.norm())))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
p = E.EngineProblem(E.EngineMesh.unit_square(n), 2)
p.enable_multigrid(); p.upload(0)
u = p.new_vector(p.N, 0.0); f = p.new_vector(p.M[0], 0.1)
p.set_coefficient(0, u); p.set_coefficient(1, f)
vals, _ = p.assemble_jacobian()
b = p.assemble_residual()
run(p, vals, b, 'nlpoisson n=%d' % n)
del p, vals, b
if os.environ.get('SKIP_HEX'):
    sys.exit(0)
torch.cuda.empty_cache()
nx, ny, nz = 128, 64, 64
mesh = E.EngineMesh.box_hex((0, 0, 0), (2.0 * nx, 2.0 * ny, 2.0 * nz), nx, ny, nz)
fc, fl = mesh.exterior_facets()
p = E.EngineProblem(mesh, E.FAMILY_SIMP_HEX8, [0.3, 0.0, -0.25, 0.0, 3.0], tagged=np.nonzero(fl == 3)[0].astype(np.int32))
nodes = np.arange((ny + 1) * (nz + 1)) * (nx + 1)
p.set_bc([np.stack([3 * nodes, 3 * nodes + 1, 3 * nodes + 2], axis=1).ravel().astype(np.int32)])
p.enable_multigrid(); p.upload(0)
u = p.new_vector(p.N, 0.0)
rho = p.to_device(np.clip(0.86 * np.random.default_rng(0).random(p.M[0]), 1e-3, 1.0))
p.set_coefficient(0, u); p.set_coefficient(1, rho)
_, vals_bc = p.assemble_jacobian(plain=False, bc=True)
b = -p.assemble_residual()
run(p, vals_bc, b, 'hex 128x64x64')
