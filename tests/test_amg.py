"""Smoothed-aggregation AMG (precond 4; csrc/amg_setup.cpp, csrc/amg.cuh): what replaces MUMPS
(/root/reference/femo/fea/utils_dolfinx.py:405-408,476-512) on meshes without a lattice hierarchy.

CPU part: the pattern phase through the C ABI -- aggregates, the patterns of P / R / A P / P^T A P, the index lists of
the numeric phase evaluated in numpy against scipy's sparse products of the published algorithm (Vanek, Mandel,
Brezina 1996), and the quality of the resulting two-grid / V-cycle iteration.  GPU part: device numeric phase vs host,
AMG-preconditioned CG / GMRES solves vs SuperLU, iteration counts that do not grow with the mesh."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from femo_b200 import engine as E
from _cases import Case, relerr


def _values_on_pattern(p, K):
    rp, col = p.pattern(0)
    rows = np.repeat(np.arange(p.N), np.diff(rp))
    return np.asarray(sp.csr_matrix(K)[rows, col]).ravel()


def _csr(p, lv, vals=None):
    i = p.amg_level_info(lv)
    rp, col = p.amg_level_array(lv, 'rowptr'), p.amg_level_array(lv, 'col')
    v = p.amg_level_values(lv, 'vals') if vals is None else vals
    return sp.csr_matrix((v, col, rp), shape=(i['n'], i['n']))


def _prolongator(p, lv, vals=None):
    i = p.amg_level_info(lv)
    v = p.amg_level_values(lv, 'p_vals') if vals is None else vals
    return sp.csr_matrix((v, p.amg_level_array(lv, 'p_col'), p.amg_level_array(lv, 'p_rowptr')), shape=(i['n'], i['nc']))


def _poisson_case(n, famid=2):
    c = Case(famid, n, bc=True, upload=False)
    K = c.sp.newton_J(0.3 * np.sin(3 * c.coords[:, 0]) + 0.1, [c.f])
    return c, sp.csr_matrix(K), _values_on_pattern(c.p, K)


def test_aggregates_and_patterns():
    c, K, vals = _poisson_case(24)
    p = c.p
    info = p.amg_symbolic(vals, coarse_size=40)
    assert info['levels'] >= 3 and info['coarsest'] <= 40 and info['operator_complexity'] < 1.8
    bc = np.zeros(p.N, dtype=bool)
    bc[c.bc.dofs] = True
    for lv in range(info['levels'] - 1):
        i = p.amg_level_info(lv)
        rp, col, agg = p.amg_level_array(lv, 'rowptr'), p.amg_level_array(lv, 'col'), p.amg_level_array(lv, 'agg')
        if lv == 0:
            assert np.array_equal(agg < 0, bc)                     # Dirichlet rows stay out of the coarse correction
        assert agg.max() == i['nc'] - 1 and np.array_equal(np.unique(agg[agg >= 0]), np.arange(i['nc']))
        assert 3 <= (agg >= 0).sum() / i['nc'] <= 12               # aggregates of a 7-point stencil: about 7 nodes
        # P row r = the distinct aggregates among the columns of A's row r, ascending; sources partition those entries
        prp, pcol = p.amg_level_array(lv, 'p_rowptr'), p.amg_level_array(lv, 'p_col')
        ppp, pps = p.amg_level_array(lv, 'pp_ptr'), p.amg_level_array(lv, 'pp_src')
        for r in range(0, i['n'], 7):
            want = np.unique(agg[col[rp[r]:rp[r + 1]]])
            want = want[want >= 0] if agg[r] >= 0 else want[:0]
            assert np.array_equal(pcol[prp[r]:prp[r + 1]], want)
            for t in range(prp[r], prp[r + 1]):
                src = pps[ppp[t]:ppp[t + 1]]
                assert np.all(np.diff(src) > 0) and np.all((src >= rp[r]) & (src < rp[r + 1])) and np.all(agg[col[src]] == pcol[t])
        # R = P^T through the permutation
        rrp, rcol, rperm = p.amg_level_array(lv, 'r_rowptr'), p.amg_level_array(lv, 'r_col'), p.amg_level_array(lv, 'r_perm')
        P = _prolongator(p, lv)
        R = sp.csr_matrix((p.amg_level_values(lv, 'p_vals')[rperm], rcol, rrp), shape=(i['nc'], i['n']))
        assert abs(R - P.T).max() == 0.0


def test_numeric_lists_reproduce_the_published_algorithm():
    """P = (I - 4/(3 lmax) D^-1 A) T on the non-isolated rows, A P and P^T A P: the index lists evaluated in numpy and the
    library's host numeric phase against scipy sparse products."""
    c, K, vals = _poisson_case(20)
    p = c.p
    info = p.amg_symbolic(vals, coarse_size=30)
    A = K
    for lv in range(info['levels'] - 1):
        i = p.amg_level_info(lv)
        assert abs(_csr(p, lv) - A).max() <= 1e-13 * abs(A).max()
        agg = p.amg_level_array(lv, 'agg')
        keep = agg >= 0
        T = sp.csr_matrix((np.ones(keep.sum()), (np.nonzero(keep)[0], agg[keep])), shape=(i['n'], i['nc']))
        d = A.diagonal()
        lmax = np.max(np.asarray(abs(A).sum(axis=1)).ravel() / np.abs(d))
        assert abs(lmax - i['lmax_host']) <= 1e-14 * lmax
        Pref = sp.diags(keep.astype(float)) @ (T - sp.diags(4.0 / (3.0 * lmax) / d) @ A @ T)
        P = _prolongator(p, lv)
        assert abs(P - Pref).max() <= 1e-14
        # the pair lists, evaluated here
        av, pv = p.amg_level_values(lv, 'vals'), p.amg_level_values(lv, 'p_vals')
        ptr, ia, ib = p.amg_level_array(lv, 'ap_ptr'), p.amg_level_array(lv, 'ap_ia'), p.amg_level_array(lv, 'ap_ib')
        apv = np.add.reduceat(av[ia] * pv[ib], ptr[:-1])
        AP = sp.csr_matrix((apv, p.amg_level_array(lv, 'ap_col'), p.amg_level_array(lv, 'ap_rowptr')), shape=(i['n'], i['nc']))
        assert abs(AP - A @ Pref).max() <= 1e-13 * abs(A).max()
        assert np.max(np.abs(apv - p.amg_level_values(lv, 'ap_vals'))) <= 1e-14 * abs(A).max()
        ptr, ia, ib = p.amg_level_array(lv, 'ac_ptr'), p.amg_level_array(lv, 'ac_ia'), p.amg_level_array(lv, 'ac_ib')
        acv = np.add.reduceat(pv[ia] * apv[ib], ptr[:-1])
        Ac = _csr(p, lv + 1, acv)
        Acref = (Pref.T @ A @ Pref).tocsr()
        assert abs(Ac - Acref).max() <= 1e-13 * abs(A).max()
        assert Acref.nnz <= Ac.nnz                                      # the pattern is structural (keeps numerical zeros)
        A = Acref


def _vcycle_matrix(p, levels, degree=2, ratio=4.0):
    """Error propagation I - M^-1 A of the V-cycle the device runs (Chebyshev-Jacobi smoothing on [lmax/ratio, lmax],
    exact coarsest solve), as a dense matrix built from the host numeric phase."""
    def smoother(A, lmax):
        n = A.shape[0]
        Dinv = sp.diags(1.0 / A.diagonal())
        lmin = lmax / ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        # error polynomial of `degree` Chebyshev steps: e_k = T_k((theta - D^-1 A) / delta) / T_k(sigma) e_0
        Z = (theta * np.eye(n) - (Dinv @ A).toarray()) / delta
        T0, T1, s0, s1 = np.eye(n), Z, 1.0, sigma
        for _ in range(degree - 1):
            T0, T1 = T1, 2 * Z @ T1 - T0
            s0, s1 = s1, 2 * sigma * s1 - s0
        return T1 / s1
    def E(lv):
        A = _csr(p, lv)
        n = A.shape[0]
        if lv == levels - 1:
            return np.zeros((n, n))
        P = _prolongator(p, lv).toarray()
        Ad = A.toarray()
        Ac = _csr(p, lv + 1).toarray()
        Ec = E(lv + 1)
        Mc_inv = (np.eye(Ac.shape[0]) - Ec) @ np.linalg.inv(Ac)
        S = smoother(A, p.amg_level_info(lv)['lmax_host'])
        return S @ (np.eye(n) - P @ Mc_inv @ P.T @ Ad) @ S
    return E(0)


@pytest.mark.parametrize('n', [16, 32])
def test_vcycle_contracts_independently_of_the_mesh(n):
    c, K, vals = _poisson_case(n)
    p = c.p
    info = p.amg_symbolic(vals, coarse_size=30)
    Em = _vcycle_matrix(p, info['levels'])
    A = K.toarray()
    # energy-norm contraction factor of the symmetric V-cycle = spectral radius of the error propagation
    rho = np.max(np.abs(np.linalg.eigvals(Em)))
    assert rho < 0.45, rho


def test_pattern_only_hierarchy_and_vector_blocks():
    """No values: every connection strong, Dirichlet rows from the problem's marks.  Block 2 (Q1 elasticity): node-based
    aggregates, one constant per component."""
    c = Case(2, 20, bc=True, upload=False)
    info = c.p.amg_symbolic(None, coarse_size=40)
    agg = c.p.amg_level_array(0, 'agg')
    assert info['levels'] >= 2 and np.array_equal(np.nonzero(agg < 0)[0], np.unique(c.bc.dofs))
    with pytest.raises(Exception):
        c.p.amg_level_values(0, 'p_vals')
    mesh = E.EngineMesh.rectangle_quad((0.0, 0.0), (2.0, 1.0), 24, 12)
    q = E.EngineProblem(mesh, E.FAMILY_SIMP_Q1)
    x = mesh.coords()
    nodes = np.nonzero(x[:, 0] == 0.0)[0]
    q.set_bc([np.stack([2 * nodes, 2 * nodes + 1], axis=1).ravel()])
    info = q.amg_symbolic(None, coarse_size=60)
    agg = q.amg_level_array(0, 'agg')
    assert info['levels'] >= 2
    free = agg[0::2] >= 0
    assert np.array_equal(agg[0::2] < 0, agg[1::2] < 0)
    assert np.array_equal(agg[1::2][free], agg[0::2][free] + 1) and np.all(agg[0::2][free] % 2 == 0)


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_numeric_phase_matches_host(cuda_device):
    c = Case(2, 48, bc=True)
    p = c.p
    c.set_state(0.3 * np.sin(3 * c.coords[:, 0]) + 0.1)
    _, vals_bc = p.assemble_jacobian(plain=False, bc=True)
    info = p.enable_amg(vals_bc, coarse_size=60).amg
    p.amg_numeric(vals_bc)
    assert info['levels'] >= 3
    for lv in range(info['levels']):
        i = p.amg_level_info(lv)
        assert abs(i['lmax_device'] - i['lmax_host']) <= 1e-13 * i['lmax_host']
        names = ['dinv'] + (['vals'] if lv else []) + (['p_vals', 'ap_vals'] if i['nc'] else [])
        for name in names:
            h, d = p.amg_level_values(lv, name), p.amg_level_values(lv, name, from_device=True)
            assert np.max(np.abs(h - d)) <= 1e-13 * np.max(np.abs(h)), (lv, name)


def _perturbed_square(n, seed=0):
    """Unstructured stand-in: the n x n triangle lattice with interior vertices moved by up to 30 % of the cell size,
    handed over as plain arrays (no lattice hierarchy can be derived from it)."""
    from oracle import mesh as om
    m = om.unit_square_tri(n)
    x = m.coords.copy()
    rng = np.random.default_rng(seed)
    inner = (x[:, 0] > 1e-9) & (x[:, 0] < 1 - 1e-9) & (x[:, 1] > 1e-9) & (x[:, 1] < 1 - 1e-9)
    x[inner] += 0.3 / n * (rng.random((inner.sum(), 2)) - 0.5) * 2
    return x, m.cells


@pytest.mark.gpu
@pytest.mark.parametrize('n', [48, 96, 192])
def test_amg_cg_on_an_unstructured_mesh(cuda_device, n):
    import torch
    from _cases import square_boundary_lists
    x, cells = _perturbed_square(n)
    mesh = E.EngineMesh.from_arrays('triangle', x, cells)
    p = E.EngineProblem(mesh, E.FAMILY_POISSON_P1)
    p.set_bc(square_boundary_lists(x))
    p.upload(0)
    rng = np.random.default_rng(1)
    u, f = p.new_vector(p.N, 0.0), p.to_device(rng.standard_normal(p.M[0]))
    p.set_coefficient(0, u); p.set_coefficient(1, f); p.set_coefficient(2, p.new_vector(p.N, 0.0))
    _, vals_bc = p.assemble_jacobian(plain=False, bc=True)
    p.enable_amg(vals_bc)
    b = p.to_device(rng.standard_normal(p.N))
    b[torch.as_tensor(np.unique(np.concatenate(square_boundary_lists(x))), device=b.device)] = 0.0
    xs, info = p.linear_solve(vals_bc, b, rtol=1e-10, precond=4, cheb_degree=2)
    _, jac = p.linear_solve(vals_bc, b, rtol=1e-10, precond=0)
    rp, col = p.pattern(0)
    K = sp.csr_matrix((vals_bc.cpu().numpy(), col, rp), shape=(p.N, p.N))
    ref = spla.splu(K.tocsc()).solve(b.cpu().numpy())
    assert info['converged'] and relerr(xs.cpu().numpy(), ref) < 1e-8
    # mesh independent: 23 / .. / .. iterations at n = 48 / 96 / 192 on the B200 (Jacobi-CG grows like n)
    assert info['iterations'] <= 30, info
    assert jac['iterations'] > 3 * info['iterations']


@pytest.mark.gpu
def test_amg_gmres_on_the_motor_annulus(cuda_device):
    """Config 5b (nonlinear magnetostatics, 216 tagged subdomains, non-symmetric Jacobian): SNES with AMG-preconditioned
    GMRES against the oracle's SuperLU Newton path, and against the round-1 Chebyshev-polynomial preconditioner."""
    from _cases_motor import MotorCase
    c = MotorCase(nr=24, nth=96, uscale=0.0)               # A_z = 0 start, seeded mesh displacement
    p = c.p
    p.set_param(6, 0.2)                                    # first step of the example's load ramp
    vals, _ = p.assemble_jacobian()
    p.enable_amg(vals)
    ia = p.newton_solve(kind='SNES', krylov_rtol=1e-11, krylov_max_it=2000, method=1, precond=4, cheb_degree=2, cheb_ratio=4.0)
    ua = c.d_u.cpu().numpy().copy()
    c.d_u.zero_()
    ic = p.newton_solve(kind='SNES', krylov_rtol=1e-11, krylov_max_it=20000, method=1, precond=1, cheb_degree=24, cheb_ratio=600.0)
    uc = c.d_u.cpu().numpy().copy()
    assert ia['converged'] and ic['converged'] and ia['iterations'] == ic['iterations']
    assert relerr(ua, uc) < 1e-8
    # 2 400 dofs: 97 against 119 iterations; the gap opens with the mesh (1.05 M dofs: 1 932 against 14 046, profiles/)
    assert ia['krylov_iterations'] <= ic['krylov_iterations'], (ia, ic)
    c.F.js_scale = 0.2
    uo, _ = c.sp.solve_snes(np.zeros(c.F.N), [c.m])
    assert relerr(ua, uo) < 1e-7
    # adjoint (transposed) solve with the same hierarchy
    vals, _ = p.assemble_jacobian()
    b = p.assemble_output_grad(0, 0)
    la, li = p.linear_solve(vals, b, transpose=True, rtol=1e-11, method=1, precond=4, cheb_degree=2, cheb_ratio=4.0)
    K = c.csr(0, vals)
    ref = spla.splu(K.T.tocsc()).solve(b.cpu().numpy())
    assert li['converged'] and relerr(la.cpu().numpy(), ref) < 1e-8


@pytest.mark.gpu
def test_amg_gmres_on_the_mesh_motion_family(cuda_device):
    """Config 5a (2 dofs per node, one-sided Nitsche facets => far from symmetric): plain aggregation with node-based
    aggregates and a degree-4 smoother, SNES state against the oracle's SuperLU path."""
    from _cases_motor import MotorMMCase
    c = MotorMMCase(32, 128, scale=0.0)
    p, F = c.p, c.F
    g = c.radial_bc(0.1 * (0.06 / 32) / 0.09)              # the interior circle grows by a tenth of a cell
    c.d_m.copy_(p.to_device(g))
    c.d_u.zero_()
    vals, _ = p.assemble_jacobian()
    info = p.enable_amg(vals, omega_scale=0.0).amg
    agg = p.amg_level_array(0, 'agg')
    assert info['levels'] >= 3 and np.array_equal(agg[1::2], agg[0::2] + 1)
    ni = p.newton_solve(kind='SNES', krylov_rtol=1e-12, krylov_max_it=3000, method=1, precond=4, cheb_degree=4, cheb_ratio=8.0)
    xo, _ = c.sp.solve_snes(np.zeros(F.N), [g])
    assert ni['converged'] in (1, 2, 3) and relerr(c.d_u.cpu().numpy(), xo) < 1e-7
    assert ni['krylov_iterations'] <= 70 * ni['iterations'], ni
