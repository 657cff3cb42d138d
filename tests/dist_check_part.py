"""Multi-rank run on an UNSTRUCTURED partition (torchrun, one rank per GPU or all ranks on cuda:0): a perturbed, cell-shuffled
triangle mesh cut by femo_b200.partition (RCB, owned-first numbering, one-cell ghost layer), the Poisson family of
examples/poisson_opt with Dirichlet rows, against the unpartitioned problem on the same GPU: owned rows of the residual
and of an SpMV with poisoned ghosts, the all-reduced functional, the AMG-preconditioned CG solve, the Newton state and the
adjoint total derivative dJ/df on the owned cells.  Exit code 0 = every rank agrees."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from femo_b200 import engine as E  # noqa: E402
from femo_b200 import dist as fd  # noqa: E402
from femo_b200 import partition as P  # noqa: E402


def relerr(a, b):
    den = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (den if den > 0 else 1.0)


def boundary_nodes(x):
    return np.nonzero((np.abs(x[:, 0]) < 1e-9) | (np.abs(x[:, 0] - 1) < 1e-9) | (np.abs(x[:, 1]) < 1e-9) | (np.abs(x[:, 1] - 1) < 1e-9))[0]


def main():
    same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'
    lr = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    dist.init_process_group('gloo' if same else 'nccl', **({} if same else dict(device_id=torch.device('cuda', lr))))
    rank, R = fd.init(lr)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    famid = int(sys.argv[2]) if len(sys.argv) > 2 else E.FAMILY_POISSON_P1
    from oracle import mesh as om                           # test infrastructure: only the lattice generator
    m0 = om.unit_square_tri(n, n + 5)
    rng = np.random.default_rng(3)
    x = m0.coords.copy()
    inner = np.ones(x.shape[0], dtype=bool)
    inner[boundary_nodes(x)] = False
    x[inner] += 0.3 / n * (rng.random((inner.sum(), 2)) - 0.5)
    cells = m0.cells[rng.permutation(m0.cells.shape[0])].astype(np.int64)
    _, _, views = P.partition_mesh(x, cells, R)
    v = views[rank]
    # global fields (replicated), restricted to the local numbering
    fg = 1.0 + 0.5 * np.sin(np.arange(cells.shape[0]) * 0.37)
    ug = rng.standard_normal(x.shape[0])
    uex = np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) / (2 * np.pi ** 2)
    bg = boundary_nodes(x)

    linear = famid == E.FAMILY_POISSON_P1

    def build(prob, lx, lverts, lcells):
        if linear:
            prob.set_bc([boundary_nodes(lx).astype(np.int32)])
        prob.upload(lr)
        u, f = prob.to_device((ug if linear else 0.2 * ug)[lverts]), prob.to_device(fg[lcells])
        prob.set_coefficient(0, u); prob.set_coefficient(1, f)
        if linear:
            prob.set_coefficient(2, prob.to_device(uex[lverts]))
        return u, f

    gmesh = E.EngineMesh.from_arrays('triangle', x, cells)
    # the nonlinear Poisson form carries Nitsche terms on the TRUE boundary: the global mesh's exterior facets, restricted per rank
    gfac = None if linear else gmesh.exterior_facets()
    p = fd.PartProblem(famid, views, rank, facets=gfac)
    u, f = build(p, v.coords, v.verts_global, v.cells_global)
    pg = E.EngineProblem(gmesh, famid, facets=gfac)
    gu, gf = build(pg, x, np.arange(x.shape[0]), np.arange(cells.shape[0]))
    no, nc = v.n_owned_verts, v.n_owned_cells
    own_v, own_c = v.verts_global[:no], v.cells_global[:nc]
    fails = []

    def chk(name, a, b, tol):
        e = relerr(a, b)
        if not e < tol:
            fails.append('%s %.3e' % (name, e))

    # assembly: owned rows complete without communication
    chk('residual', p.assemble_residual().cpu().numpy()[:no], pg.assemble_residual().cpu().numpy()[own_v], 1e-12)
    vals, vbc = p.assemble_jacobian(plain=True, bc=linear)
    gvals, gvbc = pg.assemble_jacobian(plain=True, bc=linear)
    if not linear:
        vbc, gvbc = vals, gvals
    # SpMV with poisoned ghosts: the engine refreshes them
    xs = rng.standard_normal(x.shape[0])
    xl = p.to_device(xs[v.verts_global])
    xl[no:] = float('nan')
    chk('spmv', p.spmv(0, vals, xl).cpu().numpy()[:no], pg.spmv(0, gvals, pg.to_device(xs)).cpu().numpy()[own_v], 1e-13)
    J, Jg = p.assemble_output(0), pg.assemble_output(0)
    if not abs(J - Jg) <= 1e-12 * abs(Jg):
        fails.append('functional %r vs %r' % (J, Jg))
    # AMG-preconditioned CG (per-rank block preconditioner) against the unpartitioned AMG-CG solve
    p.enable_amg(vbc)
    pg.enable_amg(gvbc)
    b = rng.standard_normal(x.shape[0])
    if linear:
        b[bg] = 0.0
    xs1, i1 = p.linear_solve(vbc, p.to_device(b[v.verts_global]), rtol=1e-11, precond=4, cheb_degree=2, cheb_ratio=4.0)
    xs2, i2 = pg.linear_solve(gvbc, pg.to_device(b), rtol=1e-11, precond=4, cheb_degree=2, cheb_ratio=4.0)
    if not (i1['converged'] and i2['converged']):
        fails.append('CG did not converge %r %r' % (i1, i2))
    chk('linear solve', xs1.cpu().numpy()[:no], xs2.cpu().numpy()[own_v], 1e-8)
    chk('linear solve ghosts', xs1.cpu().numpy()[no:], xs2.cpu().numpy()[v.verts_global[no:]], 1e-8)
    if i1['iterations'] > 2 * i2['iterations'] + 10:
        fails.append('block preconditioner too weak: %d vs %d iterations' % (i1['iterations'], i2['iterations']))
    # state (reference's NewtonSolver, 3 fixed iterations) and the adjoint total derivative dJ/df
    kw = dict(kind='Newton' if linear else 'SNES', krylov_rtol=1e-11, precond=4, cheb_degree=2, cheb_ratio=4.0)
    p.newton_solve(**kw)
    pg.newton_solve(**kw)
    chk('state', u.cpu().numpy()[:no], gu.cpu().numpy()[own_v], 1e-8)

    def total(prob):
        pl, a = prob.assemble_jacobian(plain=True, bc=linear)
        a = a if linear else pl
        lam, li = prob.linear_solve(a, prob.assemble_output_grad(0, 0), transpose=True, rtol=1e-11, precond=4, cheb_degree=2,
                                    cheb_ratio=4.0)
        g = prob.assemble_output_grad(0, 1)
        prob.axpy(-1.0, prob.spmv(1, prob.assemble_dRdm(0), lam, transpose=True), g)
        return g.cpu().numpy(), li
    gl, li = total(p)
    gg, _ = total(pg)
    chk('dJ/df', gl[:nc], gg[own_c], 1e-7)
    if fd.stats()['link_error']:
        fails.append('link transport timed out')
    torch.cuda.synchronize()
    flag = torch.tensor([len(fails)], device='cpu' if same else 'cuda')
    dist.all_reduce(flag)
    for msg in fails:
        print('[rank %d] FAIL %s' % (rank, msg), flush=True)
    if rank == 0:
        print('dist_check_part n=%d family=%d ranks=%d: %s (CG iterations %d partitioned / %d single, adjoint %d)'
              % (n, famid, R, 'OK' if flag.item() == 0 else 'FAILED', i1['iterations'], i2['iterations'], li['iterations']), flush=True)
    fd.finalize()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


def main_motor():
    """Config 5b on a partition: nonlinear magnetostatics on the synthetic annulus (216 tagged subdomains, Nitsche terms on
    the two boundary circles, non-symmetric Jacobian), RCB-partitioned as a plain array mesh; SNES + GMRES with the
    distributed AMG against the same problem on one GPU."""
    same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'
    lr = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    dist.init_process_group('gloo' if same else 'nccl', **({} if same else dict(device_id=torch.device('cuda', lr))))
    rank, R = fd.init(lr)
    from femo_b200.forms import motor as pde
    from femo_b200.fea.fem import Mesh
    nr, nth = 24, 96
    am = E.EngineMesh.annulus(nr, nth)
    x, cells = am.coords(), am.cells().astype(np.int64)
    tags = pde.synthetic_motor_tags(Mesh(am, 'triangle'))
    gmesh = E.EngineMesh.from_arrays('triangle', x, cells)
    gfac = gmesh.exterior_facets()
    params = pde.em_params(838.e3, 12, 36, 4e-7 * np.pi, 0.0, 282.2 / 0.00016231)
    _, _, views = P.partition_mesh(x, cells, R)
    v = views[rank]
    rng = np.random.default_rng(5)
    uh = 2e-4 * rng.standard_normal(2 * x.shape[0])                       # mesh displacement (input), 2 dofs per vertex

    def build(prob, lverts):
        prob.upload(lr)
        prob.set_param(6, 0.2)                                             # first step of the load ramp
        idx = (2 * lverts[:, None] + np.arange(2)[None, :]).ravel()
        u, m = prob.new_vector(prob.N, 0.0), prob.to_device(uh[idx])
        prob.set_coefficient(0, u); prob.set_coefficient(1, m)
        return u, m

    p = fd.PartProblem(E.FAMILY_MOTOR_EM, views, rank, params=params, cell_tags=tags, facets=gfac)
    pg = E.EngineProblem(gmesh, E.FAMILY_MOTOR_EM, params, facets=gfac, cell_tags=tags)
    u, m = build(p, v.verts_global)
    gu, gm = build(pg, np.arange(x.shape[0]))
    no = v.n_owned_verts
    own_v = v.verts_global[:no]
    own_d = (2 * own_v[:, None] + np.arange(2)[None, :]).ravel()
    fails = []

    def chk(name, a, b, tol):
        e = relerr(a, b)
        if not e < tol:
            fails.append('%s %.3e' % (name, e))
    chk('residual', p.assemble_residual().cpu().numpy()[:no], pg.assemble_residual().cpu().numpy()[own_v], 1e-11)
    vals, _ = p.assemble_jacobian()
    gvals, _ = pg.assemble_jacobian()
    p.enable_amg(vals)
    pg.enable_amg(gvals)
    kw = dict(kind='SNES', krylov_rtol=1e-11, krylov_max_it=3000, method=1, precond=4, cheb_degree=2, cheb_ratio=4.0)
    ni, nig = p.newton_solve(**kw), pg.newton_solve(**kw)
    chk('state', u.cpu().numpy()[:no], gu.cpu().numpy()[own_v], 1e-7)
    J, Jg = p.assemble_output(0), pg.assemble_output(0)
    if not abs(J - Jg) <= 1e-8 * abs(Jg):
        fails.append('functional %r vs %r' % (J, Jg))

    def total(prob):
        a, _ = prob.assemble_jacobian()
        lam, li = prob.linear_solve(a, prob.assemble_output_grad(0, 0), transpose=True, rtol=1e-11, max_it=3000, method=1,
                                    precond=4, cheb_degree=2, cheb_ratio=4.0)
        g = prob.assemble_output_grad(0, 1)
        prob.axpy(-1.0, prob.spmv(1, prob.assemble_dRdm(0), lam, transpose=True), g)
        return g.cpu().numpy(), li
    gl, li = total(p)
    gg, lig = total(pg)
    if not (li['converged'] and lig['converged']):
        fails.append('adjoint GMRES did not converge')
    chk('dJ/duhat', gl[:2 * no], gg[own_d], 1e-6)
    if fd.stats()['link_error']:
        fails.append('link transport timed out')
    torch.cuda.synchronize()
    flag = torch.tensor([len(fails)], device='cpu' if same else 'cuda')
    dist.all_reduce(flag)
    for msg in fails:
        print('[rank %d] FAIL %s' % (rank, msg), flush=True)
    if rank == 0:
        print('dist_check_part motor %dx%d ranks=%d: %s (GMRES iterations %d partitioned / %d single over %d Newton steps, adjoint %d / %d)'
              % (nr, nth, R, 'OK' if flag.item() == 0 else 'FAILED', ni['krylov_iterations'], nig['krylov_iterations'], ni['iterations'],
                 li['iterations'], lig['iterations']), flush=True)
    fd.finalize()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == '__main__':
    sys.exit(main_motor() if len(sys.argv) > 2 and sys.argv[2] == 'motor' else main())
