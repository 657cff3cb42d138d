import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, scipy.sparse.linalg as spla
from _cases import relerr
from _cases_motor import MotorCase
for nr, nth in ((6, 24), (12, 48)):
    c = MotorCase(nr, nth, seed=2)
    vals, _ = c.p.assemble_jacobian()
    A = c.csr(0, vals); b = np.random.default_rng(0).standard_normal(c.F.N)
    xo = spla.spsolve(A.tocsc(), b)
    d = A.diagonal()
    print('N', c.F.N, 'cond est', np.linalg.cond(A.toarray()), 'cond jacobi', np.linalg.cond((A.toarray().T / d).T))
    import time, torch
    for pre, deg, ratio in ((0, 0, 0), (1, 8, 60), (1, 12, 150), (1, 16, 300), (1, 24, 600)):
        torch.cuda.synchronize(); t = time.time()
        x, info = c.p.linear_solve(vals, c.p.to_device(b), rtol=1e-12, method=1, precond=pre, max_it=3000, cheb_degree=deg, cheb_ratio=ratio)
        torch.cuda.synchronize()
        print('pre', pre, deg, ratio, info['iterations'], info['converged'], '%.3g' % info['rnorm'], '%.2e' % relerr(x.cpu().numpy(), xo), '%.0f ms' % ((time.time() - t) * 1e3))
