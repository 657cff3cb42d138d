"""Launch-bound regime: GMG-PCG solves on small lattices with the PCG iteration replayed from a CUDA graph
(default) against plain launches (FEMO_NO_GRAPH=1).  Prints ms per solve, iterations, graph replays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E
for n in (128, 256, 512):
    p = E.EngineProblem(E.EngineMesh.unit_square(n), E.FAMILY_NLPOISSON_P1)
    p.enable_multigrid(); p.upload(0)
    u, f = p.new_vector(p.N, 0.0), p.new_vector(p.M[0], 0.1)
    p.set_coefficient(0, u); p.set_coefficient(1, f)
    vals, _ = p.assemble_jacobian()
    b = p.to_device(np.random.default_rng(0).standard_normal(p.N))
    for mode in ('graph', 'plain'):
        if mode == 'plain':
            os.environ['FEMO_NO_GRAPH'] = '1'
        for _ in range(3):
            x, info = p.linear_solve(vals, b, rtol=1e-10, precond=2, cheb_degree=2)
        torch.cuda.synchronize(); g0 = p.graph_replays(); l0 = p.launch_count(); t = time.perf_counter()
        for _ in range(20):
            x, info = p.linear_solve(vals, b, rtol=1e-10, precond=2, cheb_degree=2)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
        print('n=%d (%d dofs) %s: %.3f ms per solve, %d iterations, %d graph replays, %d kernel launches per solve'
              % (n, p.N, mode, dt * 1e3, info['iterations'], (p.graph_replays() - g0) // 20, (p.launch_count() - l0) // 20), flush=True)
        os.environ.pop('FEMO_NO_GRAPH', None)
