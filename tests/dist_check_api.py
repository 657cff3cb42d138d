"""Multi-rank run THROUGH THE femo API (torchrun, one rank per GPU or all ranks on cuda:0): FEA + FEAModel + Simulator on
this rank's slab of the unit square (createUnitSquareMesh under an initialised femo_b200.dist), state solve and adjoint
totals, compared with the unpartitioned engine problem solved on the same GPU.  Exit code 0 = every rank agrees."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femo_b200 import engine as E  # noqa: E402
from femo_b200 import dist as fd  # noqa: E402


def relerr(a, b):
    den = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (den if den > 0 else 1.0)


def main():
    same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'
    lr = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    if same:
        dist.init_process_group('gloo')
    else:
        dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    rank, R = fd.init(lr)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    from femo_b200.fea.fea_b200 import FEA, createUnitSquareMesh, FunctionSpace, Function, TestFunction
    from femo_b200.forms.nonlinear_poisson import pdeRes, outputForm
    from femo_b200.csdl_opt import FEAModel, Simulator
    from femo_b200.fea import utils_b200
    utils_b200.KRYLOV['rtol'] = 1e-12
    mesh = createUnitSquareMesh(n)
    assert mesh.slab is not None and mesh.slab['rank'] == rank
    fea = FEA(mesh)
    f = Function(FunctionSpace(mesh, ('DG', 0)))
    Vu = FunctionSpace(mesh, ('CG', 1))
    u = Function(Vu)
    fea.add_input('f', f)
    fea.add_state(name='u', function=u, residual_form=pdeRes(u, TestFunction(Vu), f), arguments=['f'])
    fea.add_output(name='l2_functional', type='scalar', form=outputForm(u, f), arguments=['f', 'u'])
    fea.PDE_SOLVER = 'SNES'
    fea.REPORT = False
    model = FEAModel(fea=[fea], debug_mode=False)
    nf = fea.inputs_dict['f']['shape']
    sl = mesh.slab
    rows = n // R
    assert nf == 2 * n * rows and fea.states_dict['u']['shape'] == (n + 1) * (rows + (1 if rank == R - 1 else 0))
    # a global, rank-independent input field: f(cell) from the global cell index
    gcell = np.arange(2 * n * n, dtype=np.float64).reshape(n, 2 * n)[rank * rows:(rank + 1) * rows].ravel()
    f_own = 0.1 + 0.05 * np.sin(gcell * 0.37)
    model.create_input('f', shape=nf, val=f_own)
    sim = Simulator(model)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
        tot = sim.compute_totals('l2_functional', 'f')[('l2_functional', 'f')]
    J = float(np.ravel(sim['l2_functional'])[0])
    # the same problem unpartitioned, engine level, on this rank's GPU
    pg = E.EngineProblem(E.EngineMesh.unit_square(n), E.FAMILY_NLPOISSON_P1)
    pg.enable_multigrid()
    pg.upload(lr)
    fg = 0.1 + 0.05 * np.sin(np.arange(2 * n * n, dtype=np.float64) * 0.37)
    gu, gf = pg.new_vector(pg.N, 0.0), pg.to_device(fg)
    pg.set_coefficient(0, gu); pg.set_coefficient(1, gf)
    pg.newton_solve(kind='SNES', krylov_rtol=1e-12, precond=2, cheb_degree=2)
    Jg = pg.assemble_output(0)
    vals, _ = pg.assemble_jacobian()
    lam, _ = pg.linear_solve(vals, pg.assemble_output_grad(0, 0), transpose=True, rtol=1e-12, precond=2, cheb_degree=2)
    gg = pg.assemble_output_grad(0, 1)
    pg.axpy(-1.0, pg.spmv(1, pg.assemble_dRdm(0), lam, transpose=True), gg)
    a = rank * rows
    b = a + rows + (1 if rank == R - 1 else 0)
    fails = []
    e = relerr(np.asarray(sim['u']), gu.cpu().numpy().reshape(n + 1, n + 1)[a:b].ravel())
    if not e < 1e-8:
        fails.append('state %.3e' % e)
    if not abs(J - Jg) <= 1e-10 * abs(Jg):
        fails.append('J %r vs %r' % (J, Jg))
    e = relerr(np.asarray(tot), gg.cpu().numpy().reshape(n, 2 * n)[rank * rows:(rank + 1) * rows].ravel())
    if not e < 1e-7:
        fails.append('total derivative %.3e' % e)
    if fd.stats()['link_error']:
        fails.append('link transport timed out')
    torch.cuda.synchronize()
    flag = torch.tensor([len(fails)], device='cpu' if same else 'cuda')
    dist.all_reduce(flag)
    for m in fails:
        print('[rank %d] FAIL %s' % (rank, m), flush=True)
    if rank == 0:
        print('dist_check_api n=%d ranks=%d: %s (J=%.12g)' % (n, R, 'OK' if flag.item() == 0 else 'FAILED', J), flush=True)
    fd.finalize()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == '__main__':
    sys.exit(main())
