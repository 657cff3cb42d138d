"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
the slab-partitioned engine vs the same problem solved unpartitioned on each
rank's own GPU.  Exit code 0 = all comparisons passed on every rank."""
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
    same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'      # all ranks on cuda:0 (link transport, 1-GPU boxes)
    lr = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    if same:
        dist.init_process_group('gloo')
    else:
        dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    rank, R = fd.init(lr)
    famid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    gny = int(sys.argv[3]) if len(sys.argv) > 3 else 64 * R
    fails = []

    def check(name, a, b, tol):
        e = relerr(np.asarray(a), np.asarray(b))
        if not e < tol:
            fails.append('%s: rel err %.3e > %.1e' % (name, e, tol))

    p = fd.SlabProblem(famid, nx, gny, rank, R)
    levels = p.enable_multigrid()
    pg = E.EngineProblem(E.EngineMesh.unit_square(nx, gny), famid)
    pg.enable_multigrid()
    s = p.slab
    rows = slice(s['crow0'], s['crow0'] + s['ncrows'] + 1)
    crows = slice(s['crow0'], s['crow0'] + s['ncrows'])
    orow = slice(s['crow0'] + s['own0'], s['crow0'] + s['own1'])
    ocrow = slice(s['crow0'] + s['cown0'], s['crow0'] + s['cown1'])
    xg = pg.mesh.coords()
    if famid == 1:
        lists = [np.nonzero(np.isclose(xg[:, a], b, atol=1e-6))[0] for a, b in ((0, 0.), (0, 1.), (1, 0.), (1, 1.))]
        pg.set_bc(lists)
        xl = p.local_coords()
        p.set_bc([np.nonzero(np.isclose(xl[:, a], b, atol=1e-6))[0] for a, b in ((0, 0.), (0, 1.), (1, 0.), (1, 1.))])
    check('coords', p.local_coords(), xg.reshape(gny + 1, nx + 1, 2)[rows].reshape(-1, 2), 1e-300)
    p.upload(lr)
    pg.upload(lr)
    rng = np.random.default_rng(7)
    ug = 0.3 * rng.standard_normal(pg.N)
    fg = rng.standard_normal(pg.M[0])

    def loc_nodes(v):
        return np.ascontiguousarray(v.reshape(gny + 1, nx + 1)[rows]).ravel()

    def loc_cells(v):
        return np.ascontiguousarray(v.reshape(gny, 2 * nx)[crows]).ravel()

    def own_nodes_g(v):
        return v.reshape(gny + 1, nx + 1)[orow].ravel()

    def own_cells_g(v):
        return v.reshape(gny, 2 * nx)[ocrow].ravel()

    def own_nodes_l(t):
        return p.owned(t).cpu().numpy()

    def own_cells_l(t):
        return t[s['cown_off']:s['cown_off'] + s['cown_n']].cpu().numpy()

    d_u, d_f = p.to_device(loc_nodes(ug)), p.to_device(loc_cells(fg))
    g_u, g_f = pg.to_device(ug), pg.to_device(fg)
    p.set_coefficient(0, d_u); p.set_coefficient(1, d_f)
    pg.set_coefficient(0, g_u); pg.set_coefficient(1, g_f)
    if famid == 1:
        uex = np.sin(xg[:, 0]) * np.cos(xg[:, 1])
        p.set_coefficient(2, p.to_device(loc_nodes(uex)))
        pg.set_coefficient(2, pg.to_device(uex))
    TOL = 1e-12
    # assembly: owned rows need no communication
    check('residual', own_nodes_l(p.assemble_residual()), own_nodes_g(pg.assemble_residual().cpu().numpy()), TOL)
    bc = famid == 1
    v, vbc = p.assemble_jacobian(plain=True, bc=bc)
    vg, vgbc = pg.assemble_jacobian(plain=True, bc=bc)
    xv = rng.standard_normal(pg.N)
    # SpMV with halo exchange: scramble the ghost rows first to prove they are refreshed
    xl = loc_nodes(xv).reshape(-1, nx + 1).copy()
    if s['own0'] > 0:
        xl[0] = 1e30
    if rank < R - 1:
        xl[-1] = -1e30
    y = p.spmv(0, vbc if bc else v, p.to_device(xl.ravel()))
    check('spmv+halo', own_nodes_l(y), own_nodes_g(pg.spmv(0, vgbc if bc else vg, pg.to_device(xv)).cpu().numpy()), 1e-13)
    dv, dvg = p.assemble_dRdm(0), pg.assemble_dRdm(0)
    check('dRdm^T x', own_cells_l(p.spmv(1, dv, p.to_device(loc_nodes(xv)), transpose=True)),
          own_cells_g(pg.spmv(1, dvg, pg.to_device(xv), transpose=True).cpu().numpy()), 1e-13)
    Jl, Jg = p.assemble_output(0), pg.assemble_output(0)
    if not abs(Jl - Jg) <= 1e-12 * abs(Jg):
        fails.append('output %r vs %r' % (Jl, Jg))
    check('dJdu', own_nodes_l(p.assemble_output_grad(0, 0)), own_nodes_g(pg.assemble_output_grad(0, 0).cpu().numpy()), TOL)
    check('dJdm', own_cells_l(p.assemble_output_grad(0, 1)), own_cells_g(pg.assemble_output_grad(0, 1).cpu().numpy()), TOL)
    # distributed GMG-PCG vs single GPU
    b = rng.standard_normal(pg.N)
    x, info = p.linear_solve(vbc if bc else v, p.to_device(loc_nodes(b)), rtol=1e-12, precond=2)
    xg_, infog = pg.linear_solve(vgbc if bc else vg, pg.to_device(b), rtol=1e-12, precond=2)
    check('gmg-pcg solve', own_nodes_l(x), own_nodes_g(xg_.cpu().numpy()), 1e-8)
    if not info['converged'] or info['iterations'] > infog['iterations'] + 3:
        fails.append('distributed PCG iterations %r vs single %r' % (info, infog))
    # nonlinear state solve + adjoint gradient
    f0 = 0.1 * np.ones(pg.M[0]) if famid == 2 else fg
    d_f.copy_(p.to_device(loc_cells(f0))); g_f.copy_(pg.to_device(f0))
    d_u.zero_(); g_u.zero_()
    kind = 'SNES' if famid == 2 else 'Newton'
    ni = p.newton_solve(kind=kind, krylov_rtol=1e-12, precond=2)
    nig = pg.newton_solve(kind=kind, krylov_rtol=1e-12, precond=2)
    check('state', own_nodes_l(d_u), own_nodes_g(g_u.cpu().numpy()), 1e-8)
    if ni['iterations'] != nig['iterations']:
        fails.append('newton iterations %r vs %r' % (ni, nig))
    v, vbc = p.assemble_jacobian(plain=True, bc=bc)
    vg, vgbc = pg.assemble_jacobian(plain=True, bc=bc)
    lam, li = p.linear_solve(vbc if bc else v, p.assemble_output_grad(0, 0), transpose=True, rtol=1e-12, precond=2)
    lamg, _ = pg.linear_solve(vgbc if bc else vg, pg.assemble_output_grad(0, 0), transpose=True, rtol=1e-12, precond=2)
    g = p.assemble_output_grad(0, 1)
    p.axpy(-1.0, p.spmv(1, p.assemble_dRdm(0), lam, transpose=True), g)
    gg = pg.assemble_output_grad(0, 1)
    pg.axpy(-1.0, pg.spmv(1, pg.assemble_dRdm(0), lamg, transpose=True), gg)
    check('total derivative', own_cells_l(g), own_cells_g(gg.cpu().numpy()), 1e-7)
    torch.cuda.synchronize()
    if fd.stats()['link_error']:
        fails.append('link transport timed out')
    flag = torch.tensor([len(fails)], device='cpu' if same else 'cuda')
    dist.all_reduce(flag)
    for f in fails:
        print('[rank %d] FAIL %s' % (rank, f), flush=True)
    if rank == 0:
        print('dist_check family %d nx=%d gny=%d ranks=%d levels=%d: %s  (%s)' %
              (famid, nx, gny, R, levels, 'OK' if flag.item() == 0 else 'FAILED', fd.stats()), flush=True)
    fd.finalize()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == '__main__':
    sys.exit(main())
