"""Assembly timings of the register-heavy element kernels (CUDA events, 10 launches each after 3 warm-ups): the motor
families on the synthetic annulus (dual-number Jacobians: k_motor_mm<JAC> spills, femo_b200/csrc/ptxas.log) and the general
hexahedron cell kernel k_simp_hex_cell (238 registers), which uniform boxes bypass (k_segreduce_jac_k0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from femo_b200 import engine as E
from femo_b200.forms import motor as pde
from femo_b200.fea.fem import Mesh


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nr, nth = (int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else '512x2048').split('x'))
em = E.EngineMesh.annulus(nr, nth)
mesh = Mesh(em, 'triangle')
tags = pde.synthetic_motor_tags(mesh)
sides = [pde.annulus_circle_sides(mesh, k) for k in (0, nr // 2, nr)]
fc, fl = np.concatenate([q[0] for q in sides]), np.concatenate([q[1] for q in sides])
o = np.lexsort((fl, fc))
rng = np.random.default_rng(0)
for name, p in (('magnetostatics (k_motor_em)', E.EngineProblem(em, E.FAMILY_MOTOR_EM, pde.em_params(838.e3, 12, 36, 4e-7 * np.pi, 0.0, 282.2 / 0.00016231), cell_tags=tags)),
                ('mesh motion (k_motor_mm)', E.EngineProblem(em, E.FAMILY_MOTOR_MM, [5e3], facets=(fc[o], fl[o]), cell_tags=tags))):
    p.upload(0)
    u = p.to_device(1e-4 * rng.standard_normal(p.N))
    m = p.to_device(1e-5 * rng.standard_normal(p.M[0]))
    p.set_coefficient(0, u); p.set_coefficient(1, m)
    vals = p.new_vector(p.pattern_info(0)['nnz']); dv = p.new_vector(p.pattern_info(1)['nnz']); r = p.new_vector(p.N)
    print('%s, annulus %dx%d, %d dofs, %d cells: residual %.3f ms, Jacobian %.3f ms, dR/dm %.3f ms, dJ/du %.3f ms'
          % (name, nr, nth, p.N, 2 * nr * nth, t(lambda: p.assemble_residual(r)), t(lambda: p.assemble_jacobian(out=vals)),
             t(lambda: p.assemble_dRdm(0, dv)), t(lambda: p.assemble_output_grad(0, 0))), flush=True)
    del p, vals, dv
    torch.cuda.empty_cache()
# general hexahedra: the same box as arrays (no lattice shortcut), vertices jittered
nx, ny, nz = 96, 48, 24
box = E.EngineMesh.box_hex((0.0, 0.0, 0.0), (2.0 * nx, 2.0 * ny, 2.0 * nz), nx, ny, nz)
x = box.coords() + 0.2 * (rng.random((box.nverts, 3)) - 0.5)
hm = E.EngineMesh.from_arrays('hexahedron', x, box.cells())
fcs, fls = hm.exterior_facets()
p = E.EngineProblem(hm, E.FAMILY_SIMP_HEX8, [0.3, 0.0, -0.25, 0.0, 3.0], tagged=np.arange(min(64, fcs.size), dtype=np.int32))
p.upload(0)
u = p.to_device(1e-3 * rng.standard_normal(p.N)); rho = p.to_device(0.2 + 0.8 * rng.random(p.M[0]))
p.set_coefficient(0, u); p.set_coefficient(1, rho)
vals = p.new_vector(p.pattern_info(0)['nnz']); r = p.new_vector(p.N)
print('general hexahedra (k_simp_hex_cell), %dx%dx%d jittered, %d dofs, %d cells: residual %.3f ms, Jacobian %.3f ms (%.1f GB of element blocks)'
      % (nx, ny, nz, p.N, nx * ny * nz, t(lambda: p.assemble_residual(r)), t(lambda: p.assemble_jacobian(out=vals)), 8 * 576 * nx * ny * nz / 1e9), flush=True)
