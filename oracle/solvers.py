"""State solve, adjoint solve and total derivatives with the reference's exact
call order (TEST INFRASTRUCTURE ONLY).

The reference factorises with MUMPS through PETSc KSP preonly + PC lu
(/root/reference/femo/fea/utils_dolfinx.py:405-408,476-512); scipy's SuperLU
`splu` is the direct-solver stand-in here, so oracle states carry direct-solve
accuracy and the engine's Krylov results are compared within solver tolerance.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from . import assembly as asm


def _lu_solve(A, b, transpose=False):
    lu = spla.splu(sp.csc_matrix(A))
    return lu.solve(b, trans='T' if transpose else 'N')


class StatePath:
    """One registered state of a FEA object (fea_dolfinx.py:112-127) plus the
    StateOperation callbacks that act on it (csdl_opt/state_model.py)."""

    def __init__(self, fam, bc=None):
        self.fam, self.bc = fam, bc
        self.N = fam.N

    # -- NonlinearProblem.F / .J   ----------------------
    def newton_F(self, x, m):
        b = asm.assemble_vector(self.fam.residual(x, *m), self.N)
        if self.bc is not None:
            asm.apply_lifting(b, self.fam.jacobian(x, *m), self.bc, x0=x, scale=-1.0)
            asm.set_bc(b, self.bc, x0=x, scale=-1.0)
        return b

    def newton_J(self, x, m):
        return asm.assemble_matrix(self.fam.jacobian(x, *m), (self.N, self.N), self.bc)

    def solve_newton(self, x0, m, max_it=3, rtol=1e-30, atol=1e-50, initialize=False):
        """dolfinx NewtonSolver as configured at utils_dolfinx.py:419-449: with
        atol=1e-50, rtol=1e-30, max_it=3 it always performs 3 solves (quirk B1)."""
        x = np.array(x0, dtype=np.float64, copy=True)
        if initialize:
            x[:] = 0.1                                    # utils_dolfinx.py:433-435
        hist = []
        b = self.newton_F(x, m)
        r0 = np.linalg.norm(b)
        hist.append(r0)
        it = 0
        converged = r0 < atol
        while not converged and it < max_it:
            dx = _lu_solve(self.newton_J(x, m), b)
            x -= dx
            b = self.newton_F(x, m)
            r = np.linalg.norm(b)
            hist.append(r)
            it += 1
            converged = (r < atol) or (r0 > 0 and r / r0 < rtol)
        return x, dict(iterations=it, residuals=hist, converged=bool(converged))

    def solve_snes(self, x0, m, atol=1e-13, rtol=1e-13, stol=1e-8, max_it=100):
        """PETSc SNES newtonls + basic line search, SNESConvergedDefault order
        (utils_dolfinx.py:376-416) [upstream, from memory]."""
        x = np.array(x0, dtype=np.float64, copy=True)
        b = self.newton_F(x, m)
        f0 = fn = np.linalg.norm(b)
        hist = [fn]
        reason = 'ABS' if fn < atol else None
        it = 0
        while reason is None and it < max_it:
            y = _lu_solve(self.newton_J(x, m), b)
            x -= y
            b = self.newton_F(x, m)
            fn = np.linalg.norm(b)
            hist.append(fn)
            it += 1
            if fn < atol:
                reason = 'ABS'
            elif fn <= rtol * f0:
                reason = 'REL'
            elif np.linalg.norm(y) < stol * np.linalg.norm(x):
                reason = 'STOL'
        if reason is None:
            raise RuntimeError('SNES diverged: max_it')   # error_on_nonconvergence, :399
        return x, dict(iterations=it, residuals=hist, reason=reason)

    # -- StateOperation.compute_derivatives (state_model.py:117-158) --------
    def linearise(self, u, m):
        fam = self.fam
        jac = fam.jacobian(u, *m)
        self.dRdu = asm.assemble_matrix(jac, (self.N, self.N), None)          # :132, no BC
        self.dRdm = [asm.assemble_matrix(fam.dRdm(s, u, *m), (self.N, len(m[s])), None)
                     for s in range(len(m))]                                   # :136-146, no BC
        self.A = asm.assemble_matrix(jac, (self.N, self.N), self.bc)           # :149-151, with BC
        return self.A

    # -- FEA.solveLinearBwd / Fwd (fea_dolfinx.py:192-222) -------------------
    def solve_bwd(self, seed):
        return _lu_solve(self.A, seed, transpose=True)

    def solve_fwd(self, rhs):
        """Intended A du = dR (quirk B4: the reference's ksp=None branch is wrong)."""
        return _lu_solve(self.A, rhs)

    # -- total derivative exactly as the CSDL backend chains the callbacks ---
    def total_derivative(self, k, u, m, consistent_bc=False):
        """dJ_k/dm_s = dJ/dm_s - (dR/dm_s)^T A^-T dJ/du   (SURVEY.md section 3.3).

        consistent_bc=False reproduces quirk B2 (un-BC'd dR/dm rows multiply
        lambda_bc); True zeroes those rows, the FD-consistent variant."""
        fam = self.fam
        self.linearise(u, m)
        dJdu = asm.assemble_vector(fam.output_du(k, u, *m), self.N)
        lam = self.solve_bwd(dJdu)
        if consistent_bc and self.bc is not None:
            lam = lam.copy()
            lam[self.bc.dofs] = 0.0
        out = []
        for s in range(len(m)):
            g = asm.assemble_vector(fam.output_dm(k, s, u, *m), len(m[s]))
            out.append(g - self.dRdm[s].T @ lam)
        return out, lam
