"""Target for ncu: one GMG-PCG solve of the hexahedral SIMP operator (per-GPU size of C4-3D)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
es = bench.EngineStep(int(os.environ.get('N', '256')), 0, kind='hex')
p = es.p
p.assemble_jacobian(plain=True, bc=True, out=es.vals, out_bc=es.vals_bc)
b = p.assemble_residual()
x = p.new_vector(p.N, 0.0)
x, info = p.linear_solve(es.vals_bc, b, x, rtol=1e-10, precond=2)
torch.cuda.synchronize()
print('SOLVE1', info, p.launch_count())
x.zero_()
x, info = p.linear_solve(es.vals_bc, b, x, rtol=1e-10, precond=2)
torch.cuda.synchronize()
print('SOLVE2', info, p.launch_count())
