"""Multi-GPU parity check of the hexahedral SIMP family (torchrun, one rank per GPU): the z-slab partitioned
engine vs the same box solved unpartitioned on each rank's own GPU.  Exit code 0 = all comparisons passed."""
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
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    gnz = int(sys.argv[3]) if len(sys.argv) > 3 else 16 * R
    lo, hi = (0.0, 0.0, 0.0), (2.0 * nx, 2.0 * ny, 2.0 * gnz)
    params = [0.3, 0.0, -0.25, 0.0, 3.0]
    fails = []

    def check(name, a, b, tol):
        e = relerr(np.asarray(a), np.asarray(b))
        if not e < tol:
            fails.append('%s: rel err %.3e > %.1e' % (name, e, tol))

    FACE = 1 << 3                                                  # traction on x = hi, clamp x = lo
    p = fd.SlabProblem(E.FAMILY_SIMP_HEX8, nx, gnz, rank, R, lo=lo, hi=hi, params=params, ny=ny, face_mask=FACE)
    mesh = E.EngineMesh.box_hex(lo, hi, nx, ny, gnz)
    fc, fl = mesh.exterior_facets()
    pg = E.EngineProblem(mesh, E.FAMILY_SIMP_HEX8, params, tagged=np.nonzero(fl == 3)[0].astype(np.int32))
    levels = p.enable_multigrid()
    pg.enable_multigrid()
    s = p.slab
    pl, cl = (nx + 1) * (ny + 1), nx * ny                          # nodes per plane, cells per layer
    rows = slice(s['crow0'], s['crow0'] + s['ncrows'] + 1)
    crows = slice(s['crow0'], s['crow0'] + s['ncrows'])
    orow = slice(s['crow0'] + s['own0'], s['crow0'] + s['own1'])
    ocrow = slice(s['crow0'] + s['cown0'], s['crow0'] + s['cown1'])

    def clamp(coords):
        nodes = np.nonzero(np.isclose(coords[:, 0], 0.0, atol=1e-9))[0]
        return [np.stack([3 * nodes, 3 * nodes + 1, 3 * nodes + 2], axis=1).ravel().astype(np.int32)]
    xg = mesh.coords()
    pg.set_bc(clamp(xg))
    p.set_bc(clamp(p.local_coords()))
    check('coords', p.local_coords(), xg.reshape(gnz + 1, pl, 3)[rows].reshape(-1, 3), 1e-300)
    p.upload(lr)
    pg.upload(lr)
    rng = np.random.default_rng(7)
    ug = 0.3 * rng.standard_normal(pg.N)
    rg = np.clip(0.86 * rng.random(pg.M[0]), 1e-2, 1.0)

    def loc_nodes(v):
        return np.ascontiguousarray(v.reshape(gnz + 1, pl * 3)[rows]).ravel()

    def loc_cells(v):
        return np.ascontiguousarray(v.reshape(gnz, cl)[crows]).ravel()

    def own_nodes_g(v):
        return v.reshape(gnz + 1, pl * 3)[orow].ravel()

    def own_cells_g(v):
        return v.reshape(gnz, cl)[ocrow].ravel()

    def own_nodes_l(t):
        return p.owned(t).cpu().numpy()

    def own_cells_l(t):
        return t[s['cown_off']:s['cown_off'] + s['cown_n']].cpu().numpy()

    d_u, d_r = p.to_device(loc_nodes(ug)), p.to_device(loc_cells(rg))
    g_u, g_r = pg.to_device(ug), pg.to_device(rg)
    p.set_coefficient(0, d_u); p.set_coefficient(1, d_r)
    pg.set_coefficient(0, g_u); pg.set_coefficient(1, g_r)
    TOL = 1e-12
    check('residual', own_nodes_l(p.assemble_residual()), own_nodes_g(pg.assemble_residual().cpu().numpy()), TOL)
    v, vbc = p.assemble_jacobian(plain=True, bc=True)
    vg, vgbc = pg.assemble_jacobian(plain=True, bc=True)
    xv = rng.standard_normal(pg.N)
    xl = loc_nodes(xv).reshape(-1, pl * 3).copy()
    if s['own0'] > 0:
        xl[0] = 1e30                                               # ghost planes must be refreshed by the halo exchange
    if rank < R - 1:
        xl[-1] = -1e30
    y = p.spmv(0, vbc, p.to_device(xl.ravel()))
    check('spmv+halo', own_nodes_l(y), own_nodes_g(pg.spmv(0, vgbc, pg.to_device(xv)).cpu().numpy()), 1e-13)
    dv, dvg = p.assemble_dRdm(0), pg.assemble_dRdm(0)
    check('dRdm^T x', own_cells_l(p.spmv(1, dv, p.to_device(loc_nodes(xv)), transpose=True)),
          own_cells_g(pg.spmv(1, dvg, pg.to_device(xv), transpose=True).cpu().numpy()), 1e-13)
    for k in (0, 1):
        Jl, Jg = p.assemble_output(k), pg.assemble_output(k)
        if not abs(Jl - Jg) <= 1e-12 * abs(Jg):
            fails.append('output %d: %r vs %r' % (k, Jl, Jg))
        check('dJdu %d' % k, own_nodes_l(p.assemble_output_grad(k, 0)), own_nodes_g(pg.assemble_output_grad(k, 0).cpu().numpy()), TOL)
        check('dJdm %d' % k, own_cells_l(p.assemble_output_grad(k, 1)), own_cells_g(pg.assemble_output_grad(k, 1).cpu().numpy()), TOL)
    # distributed GMG-PCG vs single GPU
    b = rng.standard_normal(pg.N)
    b[clamp(xg)[0]] = 0.0
    x, info = p.linear_solve(vbc, p.to_device(loc_nodes(b)), rtol=1e-12, precond=2, max_it=400)
    xg_, infog = pg.linear_solve(vgbc, pg.to_device(b), rtol=1e-12, precond=2, max_it=400)
    check('gmg-pcg solve', own_nodes_l(x), own_nodes_g(xg_.cpu().numpy()), 1e-7)
    if not info['converged'] or info['iterations'] > infog['iterations'] + 4:
        fails.append('distributed PCG iterations %r vs single %r' % (info, infog))
    # state solve + compliance adjoint
    d_u.zero_(); g_u.zero_()
    ni = p.newton_solve(kind='Newton', krylov_rtol=1e-12, precond=2)
    nig = pg.newton_solve(kind='Newton', krylov_rtol=1e-12, precond=2)
    check('state', own_nodes_l(d_u), own_nodes_g(g_u.cpu().numpy()), 1e-7)
    v, vbc = p.assemble_jacobian(plain=True, bc=True)
    vg, vgbc = pg.assemble_jacobian(plain=True, bc=True)
    lam, li = p.linear_solve(vbc, p.assemble_output_grad(1, 0), transpose=True, rtol=1e-12, precond=2, max_it=400)
    lamg, _ = pg.linear_solve(vgbc, pg.assemble_output_grad(1, 0), transpose=True, rtol=1e-12, precond=2, max_it=400)
    g = p.assemble_output_grad(1, 1)
    p.axpy(-1.0, p.spmv(1, p.assemble_dRdm(0), lam, transpose=True), g)
    gg = pg.assemble_output_grad(1, 1)
    pg.axpy(-1.0, pg.spmv(1, pg.assemble_dRdm(0), lamg, transpose=True), gg)
    check('total derivative', own_cells_l(g), own_cells_g(gg.cpu().numpy()), 1e-6)
    torch.cuda.synchronize()
    if fd.stats()['link_error']:
        fails.append('link transport timed out')
    flag = torch.tensor([len(fails)], device='cpu' if same else 'cuda')
    dist.all_reduce(flag)
    for f in fails:
        print('[rank %d] FAIL %s' % (rank, f), flush=True)
    if rank == 0:
        print('dist_check_hex %dx%dx%d ranks=%d levels=%d: %s  (%s; PCG its %d vs %d single)' %
              (nx, ny, gnz, R, levels, 'OK' if flag.item() == 0 else 'FAILED', fd.stats(), info['iterations'],
               infog['iterations']), flush=True)
    fd.finalize()
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == '__main__':
    sys.exit(main())
