"""Quadrature rules used by the oracle (TEST INFRASTRUCTURE ONLY).

Restates the role of basix.make_quadrature that FFCx bakes into every
generated tabulate_tensor (implicit in each `form(...)` call of the reference,
/root/reference/femo/fea/utils_dolfinx.py:173,179,185).  basix itself is not
available offline: for polynomial integrands any rule
exact to the estimated degree reproduces dolfinx to round-off; for the
non-polynomial ones the oracle's rule below IS the contract.

All rules live on the reference cell: interval [0,1], triangle
{(0,0),(1,0),(0,1)} (weights sum to 1/2), square [0,1]^2.
"""
import numpy as np


def gauss_legendre_01(m):
    """m-point Gauss-Legendre on [0,1] (exact to degree 2m-1)."""
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def interval(degree):
    m = (degree + 2) // 2
    return gauss_legendre_01(m)


def triangle(degree):
    """Points (nq,2) and weights (nq,) exact to `degree` on the reference triangle."""
    if degree <= 1:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5])
    if degree == 2:
        p = np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]])
        return p, np.full(3, 1.0 / 6.0)
    if degree <= 4:
        # 6-point Strang-Fix / Dunavant rule, closed form
        s10 = np.sqrt(10.0)
        t = np.sqrt(38.0 - 44.0 * np.sqrt(2.0 / 5.0))
        a1 = (8.0 - s10 + t) / 18.0
        a2 = (8.0 - s10 - t) / 18.0
        sw = np.sqrt(213125.0 - 53320.0 * s10)
        w1 = (620.0 + sw) / 3720.0
        w2 = (620.0 - sw) / 3720.0
        pts, wts = [], []
        for a, w in ((a1, w1), (a2, w2)):
            b = 1.0 - 2.0 * a
            pts += [[a, a], [a, b], [b, a]]
            wts += [0.5 * w / 1.0] * 3
        return np.array(pts), np.array(wts)
    # collapsed (Duffy) Gauss-Legendre x Gauss-Legendre, exact to `degree`
    m = (degree + 3) // 2
    x, wx = gauss_legendre_01(m)
    pts = np.empty((m * m, 2))
    wts = np.empty(m * m)
    k = 0
    for i in range(m):
        for j in range(m):
            pts[k, 0] = x[i]
            pts[k, 1] = x[j] * (1.0 - x[i])
            wts[k] = wx[i] * wx[j] * (1.0 - x[i])
            k += 1
    return pts, wts


def square(degree):
    m = (degree + 2) // 2
    x, w = gauss_legendre_01(m)
    pts = np.array([[xi, xj] for xi in x for xj in x])
    wts = np.array([wi * wj for wi in w for wj in w])
    return pts, wts


def cube(degree):
    m = (degree + 2) // 2
    x, w = gauss_legendre_01(m)
    pts = np.array([[a, b, c] for a in x for b in x for c in x])
    wts = np.array([a * b * c for a in w for b in w for c in w])
    return pts, wts
