"""dolfinx-0.5.1 assembly semantics restated in numpy (TEST INFRASTRUCTURE ONLY).

Follows SURVEY.md Appendix A.1-A.4 [upstream, from memory] for what the
reference reaches through
  assembleVector   /root/reference/femo/fea/utils_dolfinx.py:175-179
  assembleMatrix   utils_dolfinx.py:181-187
  assembleSystem   utils_dolfinx.py:189-202
  NonlinearSNESProblem.F / .J   utils_dolfinx.py:352-373

An "integral block" is a tuple (row_dofs (ne,nr), col_dofs (ne,nc) | None,
tensor (ne,nr[,nc])) -- one per (cell | exterior facet) integral of a form.
"""
import numpy as np
import scipy.sparse as sp


class DirichletBC:
    """Union of dolfinx `dirichletbc` objects on one space.

    `dof_lists` keeps one array per dirichletbc object because dolfinx adds the
    diagonal once per object (fem::set_diagonal through a PETSc ADD_VALUES
    setter [upstream, from memory]): a dof listed by k objects gets diagonal k.
    """

    def __init__(self, n, dof_lists, values):
        self.n = n
        self.dof_lists = [np.asarray(d, dtype=np.int32).ravel() for d in dof_lists]
        self.count = np.zeros(n, dtype=np.int32)
        for d in self.dof_lists:
            np.add.at(self.count, d, 1)
        self.marker = self.count > 0
        self.dofs = np.nonzero(self.marker)[0].astype(np.int32)
        g = np.zeros(n)
        values = np.asarray(values, dtype=np.float64)
        g[self.dofs] = values[self.dofs] if values.ndim and values.size == n else values
        self.g = g


def assemble_scalar(values):
    """Sum of per-entity functional values (utils_dolfinx.py:169-173)."""
    return float(np.sum(np.concatenate([np.ravel(v) for v in values])))


def assemble_vector(blocks, n):
    """A.1: plain sum of element vectors, Dirichlet values NOT applied."""
    b = np.zeros(n)
    for rd, _, Re in blocks:
        b += np.bincount(rd.ravel(), weights=Re.ravel(), minlength=n)
    return b


def pattern(blocks, shape):
    """Column-sorted CSR pattern = union of all (row dof, col dof) pairs (A.2)."""
    rows, cols = [], []
    for rd, cd, _ in blocks:
        rows.append(np.repeat(rd, cd.shape[1], axis=1).ravel())
        cols.append(np.tile(cd, (1, rd.shape[1])).ravel())
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols).astype(np.int64)
    key = np.unique(r * shape[1] + c)
    r, c = key // shape[1], key % shape[1]
    rowptr = np.zeros(shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), c.astype(np.int32)


def assemble_matrix(blocks, shape, bc=None):
    """A.2: element rows AND columns of BC dofs zeroed before insertion, then the
    diagonal is added once per dirichletbc object.  bc=None -> untouched."""
    rows, cols, vals = [], [], []
    for rd, cd, Ae in blocks:
        Ae = np.array(Ae, dtype=np.float64, copy=True)
        if bc is not None:
            Ae[bc.marker[rd]] = 0.0                       # rows
            Ae.transpose(0, 2, 1)[bc.marker[cd]] = 0.0    # columns
        rows.append(np.repeat(rd, cd.shape[1], axis=1).ravel())
        cols.append(np.tile(cd, (1, rd.shape[1])).ravel())
        vals.append(Ae.ravel())
    if bc is not None:
        for d in bc.dof_lists:
            rows.append(d)
            cols.append(d)
            vals.append(np.ones(d.size))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=shape).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def apply_lifting(b, blocks, bc, x0=None, scale=1.0):
    """A.4: b -= scale * A_e[:, bc cols] (g - x0), element matrices unmodified."""
    d = bc.g.copy()
    if x0 is not None:
        d -= x0
    d[~bc.marker] = 0.0
    for rd, cd, Ae in blocks:
        be = np.einsum('eij,ej->ei', Ae, d[cd])
        b -= scale * np.bincount(rd.ravel(), weights=be.ravel(), minlength=b.size)
    return b


def set_bc(b, bc, x0=None, scale=1.0):
    """A.4: b[bc] = scale * (g - x0[bc])."""
    if x0 is None:
        b[bc.dofs] = scale * bc.g[bc.dofs]
    else:
        b[bc.dofs] = scale * (bc.g[bc.dofs] - x0[bc.dofs])
    return b
