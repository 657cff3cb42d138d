"""Correctness of the BENCHED solver configuration (bench.py: fp32 DIA V-cycle, full-multigrid start,
inexact Newton, Krylov rtol 1e-10) at sizes well above the unit-test meshes:

  * n = 512 against the oracle's SuperLU path (state, adjoint, dJ/df)            -- the oracle finishes in seconds
  * n = 1024 / 2048 through size-independent properties: fp64 residual norms of the state and adjoint
    systems recomputed with the assembled fp64 operator, and a directional finite difference of dJ/df
  * the three V-cycle operator layouts (DIA fp32, CSR fp32, CSR fp64) agree

Everything goes through the C ABI (femo_b200.engine)."""
import os

import numpy as np
import pytest

from _cases import Case, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def bench_step(c, f):
    """bench.py's EngineStep.step for the nonlinear Poisson family, returning everything it computes."""
    p = c.p
    c.set_input(f)
    c.set_state(np.zeros(c.p.N))
    ni = p.newton_solve(kind='SNES', krylov_rtol=RTOL, precond=2, cheb_degree=2)
    vals, _ = p.assemble_jacobian(plain=True, bc=False)
    dv = p.assemble_dRdm(0)
    J = p.assemble_output(0)
    dJdu = p.assemble_output_grad(0, 0)
    grad = p.assemble_output_grad(0, 1)
    lam, li = p.linear_solve(vals, dJdu, transpose=True, rtol=RTOL, precond=2, cheb_degree=2)
    p.axpy(-1.0, p.spmv(1, dv, lam, transpose=True), grad)
    return dict(ni=ni, li=li, vals=vals, J=J, dJdu=dJdu, lam=lam, grad=grad)


def test_bench_settings_match_direct_solve_n512(cuda_device):
    c = Case(2, 512, seed=0, mg=True)
    f = 0.1 * np.ones(c.F.M)
    r = bench_step(c, f)
    assert r['ni']['converged'] and r['li']['converged']
    u = c.d_u.cpu().numpy()
    uo, oinfo = c.sp.solve_snes(np.zeros(c.F.N), [f])
    assert relerr(u, uo) < 1e-8
    (go,), lamo = c.sp.total_derivative(0, uo, [f])
    assert relerr(r['lam'].cpu().numpy(), lamo) < 1e-8
    assert relerr(r['grad'].cpu().numpy(), go) < 1e-8


@pytest.mark.parametrize('n', [1024, 2048])
def test_bench_settings_residuals_and_fd(cuda_device, n):
    c = Case(2, n, seed=0, mg=True, oracle=False)
    p = c.p
    M = p.M[0]
    f = 0.1 * np.ones(M)
    r = bench_step(c, f)
    assert r['ni']['converged'] and r['li']['converged']
    # nonlinear residual at the returned state, fp64 assembly
    R = p.assemble_residual()
    print('n=%d: ||F(u)|| = %.3e (||F(u0)|| = %.3e), newton its %d, krylov its %d + %d' % (
        n, float(R.norm()), r['ni']['fnorm0'], r['ni']['iterations'], r['ni']['krylov_iterations'], r['li']['iterations']))
    assert float(R.norm()) <= 1e-7 * r['ni']['fnorm0']
    # adjoint system: || A^T lam - dJ/du || / || dJ/du || with the assembled fp64 operator
    res = p.spmv(0, r['vals'], r['lam'], transpose=True)
    res -= r['dJdu']
    assert float(res.norm()) <= 10 * RTOL * float(r['dJdu'].norm())
    # directional finite difference of the reduced functional along a smooth direction
    d = 1.0 + 0.5 * np.sin(2 * np.pi * np.arange(M) / M)
    g = r['grad'].cpu().numpy()
    h = 1e-4
    Jp = bench_step(c, f + h * d)['J']
    Jm = bench_step(c, f - h * d)['J']
    fd = (Jp - Jm) / (2 * h)
    assert abs(fd - g @ d) <= 1e-5 * abs(fd), (fd, g @ d)


def test_vcycle_layouts_agree(cuda_device):
    """DIA fp32 planes, the fp32 CSR copy and the assembled fp64 values inside the V-cycle give the same
    solution (to the Krylov tolerance) in the same number of iterations (+-1)."""
    sols, its = [], []
    for mode in ('dia', 'csr32', 'csr64'):
        if mode == 'csr32':
            os.environ['FEMO_NO_DIA'] = '1'
        try:
            c = Case(2, 200, 136, seed=5, mg=True, oracle=False)
            c.set_state(0.2 * np.sin(3 * c.coords[:, 0]) * np.cos(2 * c.coords[:, 1]))
            vals, _ = c.p.assemble_jacobian(plain=True, bc=False)
            b = c.p.to_device(np.random.default_rng(2).standard_normal(c.p.N))
            x, info = c.p.linear_solve(vals, b, rtol=1e-12, precond=2, cheb_degree=2,
                                       mg_precision=1 if mode == 'csr64' else 0)
        finally:
            os.environ.pop('FEMO_NO_DIA', None)
        assert info['converged'], (mode, info)
        sols.append(x.cpu().numpy())
        its.append(info['iterations'])
    assert relerr(sols[0], sols[2]) < 1e-9 and relerr(sols[1], sols[2]) < 1e-9
    assert max(its) - min(its) <= 1, its


@pytest.mark.parametrize('deg', [1, 2, 3, 4])
def test_dia_smoother_degrees(cuda_device, deg):
    """Every Chebyshev degree takes a different kernel path (fused degree-2 pre-smoother, chained steps)."""
    c = Case(1, 96, 80, seed=1, mg=True, oracle=False)
    _, vbc = c.p.assemble_jacobian(plain=False, bc=True)
    b = c.p.to_device(np.random.default_rng(4).standard_normal(c.p.N))
    x, info = c.p.linear_solve(vbc, b, rtol=1e-11, precond=2, cheb_degree=deg)
    x64, info64 = c.p.linear_solve(vbc, b, rtol=1e-11, precond=2, cheb_degree=deg, mg_precision=1)
    assert info['converged'] and info64['converged']
    assert relerr(x.cpu().numpy(), x64.cpu().numpy()) < 1e-8
    assert abs(info['iterations'] - info64['iterations']) <= 1


@pytest.mark.parametrize('n,ny', [(1, 1), (2, 3), (33, 7), (128, 96)])
def test_node_centric_jacobian_matches_general_path(cuda_device, n, ny):
    """lattice_asm.cuh (one thread per node, rows written straight into the CSR values) against the general
    cell-kernel + sorted-segmented-reduction path, plain and BC'd copies."""
    from _cases import square_boundary_lists
    c = Case(2, n, ny, seed=n, oracle=False)
    c.p.set_bc(square_boundary_lists(c.coords), None)
    v1, b1 = c.p.assemble_jacobian(plain=True, bc=True)
    r1 = c.p.assemble_residual()
    os.environ['FEMO_NO_LATTICE_ASM'] = '1'
    try:
        v0, b0 = c.p.assemble_jacobian(plain=True, bc=True)
        r0 = c.p.assemble_residual()
    finally:
        os.environ.pop('FEMO_NO_LATTICE_ASM', None)
    assert relerr(v1.cpu().numpy(), v0.cpu().numpy()) < 1e-13
    assert relerr(b1.cpu().numpy(), b0.cpu().numpy()) < 1e-13
    assert relerr(r1.cpu().numpy(), r0.cpu().numpy()) < 1e-13


def test_fused_output_and_gradient(cuda_device):
    c = Case(2, 40, 24, seed=3, oracle=False)
    J, g = c.p.assemble_output_and_grad(0)
    assert abs(J - c.p.assemble_output(0)) <= 1e-15 * abs(J)
    assert np.array_equal(g.cpu().numpy(), c.p.assemble_output_grad(0, 0).cpu().numpy())
    c1 = Case(1, 12, seed=3, oracle=False)                      # families without a fused kernel take the two passes
    J1, g1 = c1.p.assemble_output_and_grad(0)
    assert J1 == c1.p.assemble_output(0) and np.array_equal(g1.cpu().numpy(), c1.p.assemble_output_grad(0, 0).cpu().numpy())


def test_fused_coarse_vcycle_matches_per_level_kernels(cuda_device):
    """mgfused.cuh: the cooperative coarse V-cycle kernel against the per-level launches (same row arithmetic)."""
    res = []
    os.environ['FEMO_NO_GRAPH'] = '1'          # graph replays record the per-level kernels: compare plain launches
    for fused in (True, False):
        if not fused:
            os.environ['FEMO_NO_MGFUSED'] = '1'
        try:
            c = Case(1, 300, 212, seed=2, mg=True, oracle=False)
            _, vbc = c.p.assemble_jacobian(plain=False, bc=True)
            b = c.p.to_device(np.random.default_rng(7).standard_normal(c.p.N))
            l0 = c.p.launch_count()
            x, info = c.p.linear_solve(vbc, b, rtol=1e-11, precond=2, cheb_degree=2)
            res.append((x.cpu().numpy(), info, c.p.launch_count() - l0))
        finally:
            os.environ.pop('FEMO_NO_MGFUSED', None)
    os.environ.pop('FEMO_NO_GRAPH', None)
    assert res[0][1]['converged'] and res[1][1]['converged']
    assert res[0][1]['iterations'] == res[1][1]['iterations']
    assert relerr(res[0][0], res[1][0]) < 1e-10
    assert res[0][2] < res[1][2] / 2, (res[0][2], res[1][2])          # far fewer launches


def test_cuda_graph_replay_matches_plain_launches(cuda_device):
    """Launch-bound solves replay the PCG iteration from a captured CUDA graph (krylov.cuh)."""
    res = []
    for graph in (True, False):
        if not graph:
            os.environ['FEMO_NO_GRAPH'] = '1'
        try:
            c = Case(2, 160, 120, seed=2, mg=True, oracle=False)
            c.set_state(0.1 * np.sin(2 * c.coords[:, 0]))
            vals, _ = c.p.assemble_jacobian(plain=True, bc=False)
            b = c.p.to_device(np.random.default_rng(9).standard_normal(c.p.N))
            x, info = c.p.linear_solve(vals, b, rtol=1e-12, precond=2, cheb_degree=2)
            res.append((x.cpu().numpy(), info, c.p.graph_replays()))
        finally:
            os.environ.pop('FEMO_NO_GRAPH', None)
    assert res[0][1]['converged'] and res[0][1]['iterations'] == res[1][1]['iterations']
    # the replayed graph records the per-level coarse kernels, plain launches use the cooperative coarse kernel:
    # same row arithmetic, fused multiply-adds may round differently
    assert relerr(res[0][0], res[1][0]) < 1e-11
    assert res[1][2] == 0
    assert res[0][2] == res[0][1]['iterations'] - 1, res[0][2]        # every iteration after the first is a replay


def test_hex_bench_settings_match_direct_solve(cuda_device):
    """bench.py --workload hex at a size the oracle's SuperLU path still factorises (32 x 16 x 8 cells, 15 147 dofs): the
    reference's three fixed Newton iterations with the matrix-free fp32 V-cycle / CG recurrence, density in [1e-4, 1]
    (stiffness contrast 1e-12) smoothed by the engine's 3-D cone filter, compliance and its adjoint total derivative."""
    import ctypes as C
    import torch
    from _cases34 import HexCase, _upload
    from femo_b200._lib import lib, check
    nx, ny, nz = 32, 16, 8
    c = HexCase(nx, ny, nz, seed=3, upload=False)
    assert c.p.enable_multigrid() >= 3
    _upload(c)
    p = c.p
    np.random.seed(0)
    unf = np.clip(0.86 * np.random.random(nx * ny * nz), 1e-4, 1.0)          # run_topo_opt_cantilever_beam.py:175-180
    d_unf, d_rho, d_den = p.to_device(unf), p.new_vector(unf.size), p.new_vector(unf.size)
    check(lib.femo_filter_apply3(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), nx, ny, nz, 2.0, 2.0, 2.0, 4.0,
                                 C.c_void_p(d_unf.data_ptr()), C.c_void_p(d_rho.data_ptr()), C.c_void_p(d_den.data_ptr()), 0))
    rho = d_rho.cpu().numpy()
    assert rho.min() >= 1e-4 and rho.max() <= 1.0
    c.d_m.copy_(d_rho)
    c.d_u.zero_()
    ni = p.newton_solve(kind='Newton', krylov_rtol=RTOL, precond=2, cheb_degree=2)
    assert ni['iterations'] == 3                                             # quirk B1
    vals, vals_bc = p.assemble_jacobian(plain=True, bc=True)
    J, dJdu = p.assemble_output_and_grad(1)
    grad = p.assemble_output_grad(1, 1)
    lam, li = p.linear_solve(vals_bc, dJdu, transpose=True, rtol=RTOL, precond=2, cheb_degree=2)
    assert li['converged']
    p.axpy(-1.0, p.spmv(1, p.assemble_dRdm(0), lam, transpose=True), grad)
    uo, _ = c.sp.solve_newton(np.zeros(c.F.N), [rho])
    assert relerr(c.d_u.cpu().numpy(), uo) < 1e-7
    (go,), lamo = c.sp.total_derivative(1, uo, [rho])
    assert relerr(lam.cpu().numpy(), lamo) < 1e-7
    assert relerr(grad.cpu().numpy(), go) < 1e-7
