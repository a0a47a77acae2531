"""Quadrature rules (oracle; test infrastructure only).

Restates the reference's
  * tensor Gauss-Legendre rule on hypercubes, point order i1 fastest:
    src/Quadrature/quadrature.jl:85-106 (1-D points from Golub-Welsch,
    src/Quadrature/generate_quadrature.jl:93-96; numpy's leggauss gives the same
    nodes to ~1 ulp)
  * Dunavant triangle tables (constants truncated to 14 digits, copied as
    published): src/Quadrature/gaussquad_tri_table.jl:8-31
  * Jinyun/Keast-minimal tetrahedron tables: src/Quadrature/gaussquad_tet_table.jl:2-68
"""
import math

import numpy as np

__all__ = ["QuadratureRule"]


def _dunavant(n):
    if n == 1:
        return [[0.33333333333333, 0.33333333333333, 1.00000000000000 / 2.0]]
    if n == 2:
        return [[0.16666666666667, 0.16666666666667, 0.33333333333333 / 2.0],
                [0.16666666666667, 0.66666666666667, 0.33333333333333 / 2.0],
                [0.66666666666667, 0.16666666666667, 0.33333333333333 / 2.0]]
    if n == 3:
        return [[0.33333333333333, 0.33333333333333, -0.56250000000000 / 2.0],
                [0.20000000000000, 0.20000000000000, 0.52083333333333 / 2.0],
                [0.20000000000000, 0.60000000000000, 0.52083333333333 / 2.0],
                [0.60000000000000, 0.20000000000000, 0.52083333333333 / 2.0]]
    raise ValueError("oracle: dunavant order 1..3 only")


def _keast_minimal(n):
    if n == 1:
        return [[0.25, 0.25, 0.25, 1.0 / 6.0]]
    if n == 2:
        a = (5.0 + 3.0 * math.sqrt(5.0)) / 20.0
        b = (5.0 - math.sqrt(5.0)) / 20.0
        w = 1.0 / 24.0
        return [[a, b, b, w], [b, a, b, w], [b, b, a, w], [b, b, b, w]]
    if n == 3:
        a1, a2, b2 = 1.0 / 4.0, 1.0 / 2.0, 1.0 / 6.0
        w1, w2 = -2.0 / 15.0, 3.0 / 40.0
        return [[a1, a1, a1, w1], [a2, b2, b2, w2], [b2, a2, b2, w2], [b2, b2, a2, w2], [b2, b2, b2, w2]]
    if n == 4:
        a1, w1 = 1.0 / 4.0, -74.0 / 5625.0
        a2, b2, w2 = 5.0 / 70.0, 11.0 / 14.0, 343.0 / 45000.0
        a3 = (1.0 + math.sqrt(5.0 / 14.0)) / 4.0
        b3 = (1.0 - math.sqrt(5.0 / 14.0)) / 4.0
        w3 = 28.0 / 1125.0
        return [[a1, a1, a1, w1],
                [b2, a2, a2, w2], [a2, b2, a2, w2], [a2, a2, b2, w2], [a2, a2, a2, w2],
                [a3, a3, b3, w3], [a3, b3, a3, w3], [a3, b3, b3, w3],
                [b3, a3, a3, w3], [b3, a3, b3, w3], [b3, b3, a3, w3]]
    raise ValueError("oracle: keast_minimal order 1..4 only")


class QuadratureRule:
    """QuadratureRule{shape}(order) with the reference's default rule per shape."""

    def __init__(self, shape, order=None, weights=None, points=None):
        self.shape = shape
        if weights is not None:
            self.weights = np.asarray(weights, dtype=np.float64)
            self.points = np.asarray(points, dtype=np.float64)
            return
        if shape in ("line", "quadrilateral", "hexahedron"):
            dim = {"line": 1, "quadrilateral": 2, "hexahedron": 3}[shape]
            p, w = np.polynomial.legendre.leggauss(order)
            pts, wts = [], []
            # i_1 is the innermost (fastest) loop: src/Quadrature/quadrature.jl:96-104
            for idx in np.ndindex(*([order] * dim)):
                i = idx[::-1]  # i[0] fastest
                pts.append([p[i[d]] for d in range(dim)])
                wt = 1.0
                for d in range(dim):
                    wt *= w[i[d]]
                wts.append(wt)
            self.points = np.array(pts)
            self.weights = np.array(wts)
        elif shape == "triangle":
            d = np.array(_dunavant(order))
            self.points, self.weights = d[:, :2].copy(), d[:, 2].copy()
        elif shape == "tetrahedron":
            d = np.array(_keast_minimal(order))
            self.points, self.weights = d[:, :3].copy(), d[:, 3].copy()
        else:
            raise ValueError(shape)

    @property
    def nq(self):
        return len(self.weights)
