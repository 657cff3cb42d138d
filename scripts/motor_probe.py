"""Motor magnetostatics (config 5b) at growing sizes on the synthetic annulus: 5-step load ramp (SNES + GMRES) + adjoint
of the flux-density functional.  PRECOND=cheb: degree-24 Chebyshev-Jacobi polynomial (round 1); PRECOND=amg (default):
smoothed-aggregation V-cycle (csrc/amg.cuh).  Prints times and iteration counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E
from femo_b200.forms import motor as pde
from femo_b200.fea.fem import Mesh

sizes = [(int(a), int(b)) for a, b in (s.split('x') for s in (sys.argv[1:] or ['128x512', '256x1024']))]
amg = os.environ.get('PRECOND', 'amg') == 'amg'
kw = dict(method=1, precond=4, cheb_degree=2, cheb_ratio=4.0) if amg else dict(method=1, precond=1, cheb_degree=24, cheb_ratio=600.0)
for nr, nth in sizes:
    em = E.EngineMesh.annulus(nr, nth)
    tags = pde.synthetic_motor_tags(Mesh(em, 'triangle'))
    p = E.EngineProblem(em, E.FAMILY_MOTOR_EM, pde.em_params(838.e3, 12, 36, 4e-7 * np.pi, 0.0, 282.2 / 0.00016231), cell_tags=tags)
    p.upload(0)
    u, uh = p.new_vector(p.N, 0.0), p.new_vector(p.M[0], 0.0)
    p.set_coefficient(0, u); p.set_coefficient(1, uh)
    if amg:
        t0 = time.perf_counter()
        p.set_param(6, 0.2)
        vbc, _ = p.assemble_jacobian()          # no Dirichlet rows in this family (penalty / Nitsche terms)
        p.enable_amg(vbc)
        torch.cuda.synchronize()
        print('AMG pattern phase %.2f s: %r' % (time.perf_counter() - t0, p.amg), flush=True)
    for rep in range(2):
        u.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = p.launch_count()
        kit = nit = 0
        for st in range(1, 6):
            p.set_param(6, st / 5)
            info = p.newton_solve(kind='SNES', krylov_rtol=1e-10, krylov_max_it=40000, **kw)
            kit += info['krylov_iterations']; nit += info['iterations']
        torch.cuda.synchronize(); t1 = time.perf_counter()
        vals, _ = p.assemble_jacobian()
        lam, li = p.linear_solve(vals, p.assemble_output_grad(0, 0), transpose=True, rtol=1e-10, max_it=40000, **kw)
        g = p.assemble_output_grad(0, 1)
        p.axpy(-1.0, p.spmv(1, p.assemble_dRdm(0), lam, transpose=True), g)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print('annulus %dx%d (%d dofs): ramp %.1f ms (%d Newton its, %d GMRES its), adjoint %.1f ms (%d its, converged %s), %d launches'
          % (nr, nth, p.N, (t1 - t0) * 1e3, nit, kit, (t2 - t1) * 1e3, li['iterations'], li['converged'], p.launch_count() - l0), flush=True)
    del p
    torch.cuda.empty_cache()
